"""Ring attn1 on real GPUs: (1) the hop kernels on one device (key-sharded merge, in test_kernels_gpu),
(2) the sequence-sharded train step on 2 GPUs over NCCL against the un-sharded path and the oracle."""
import socket

import pytest
import torch


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.gpu
def test_sequence_sharded_train_step_2gpu():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    import torch.multiprocessing as mp
    import ring_checks
    mp.spawn(ring_checks._spawned, args=(2, _free_port()), nprocs=2, join=True)
