import torch


def _chunked_feed_forward(ff, hidden_states, chunk_dim, chunk_size):
    n = hidden_states.shape[chunk_dim] // chunk_size
    return torch.cat([ff(h) for h in hidden_states.chunk(n, dim=chunk_dim)], dim=chunk_dim)
