"""RMSNorm / AdaLayerNormSingle leaves (diffusers/models/normalization.py, restated)."""
import torch
from torch import nn

from .embeddings import PixArtAlphaCombinedTimestepSizeEmbeddings


class RMSNorm(nn.Module):
    def __init__(self, dim, eps, elementwise_affine=True, bias=False):
        super().__init__()
        self.eps = eps
        self.elementwise_affine = elementwise_affine
        self.dim = (dim,) if isinstance(dim, int) else tuple(dim)
        self.weight = nn.Parameter(torch.ones(dim)) if elementwise_affine else None
        self.bias = None

    def forward(self, hidden_states):
        input_dtype = hidden_states.dtype
        variance = hidden_states.to(torch.float32).pow(2).mean(-1, keepdim=True)
        hidden_states = hidden_states * torch.rsqrt(variance + self.eps)
        if self.weight is not None:
            if self.weight.dtype in (torch.float16, torch.bfloat16):
                hidden_states = hidden_states.to(self.weight.dtype)
            hidden_states = hidden_states * self.weight
        else:
            hidden_states = hidden_states.to(input_dtype)
        return hidden_states


class AdaLayerNormSingle(nn.Module):
    def __init__(self, embedding_dim, use_additional_conditions=False):
        super().__init__()
        self.emb = PixArtAlphaCombinedTimestepSizeEmbeddings(
            embedding_dim, size_emb_dim=embedding_dim // 3,
            use_additional_conditions=use_additional_conditions)
        self.silu = nn.SiLU()
        self.linear = nn.Linear(embedding_dim, 6 * embedding_dim, bias=True)

    def forward(self, timestep, added_cond_kwargs=None, batch_size=None, hidden_dtype=None):
        embedded_timestep = self.emb(timestep, **added_cond_kwargs, batch_size=batch_size,
                                     hidden_dtype=hidden_dtype)
        return self.linear(self.silu(embedded_timestep)), embedded_timestep
