"""Optimizer-state sharding on 2 real GPUs over NCCL (see tests/zero_checks.py)."""
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
def test_sharded_adamw_matches_replicated_2gpu():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    import torch.multiprocessing as mp
    import zero_checks
    mp.spawn(zero_checks._spawned, args=(2, _free_port()), nprocs=2, join=True)
