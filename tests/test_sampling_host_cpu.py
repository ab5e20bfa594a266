"""Host tables of the sampling loop (sampling._step_tables, sampling._per_step), on CPU.

The captured denoise step reads its timestep row and Euler step from device tables built once per call; they must say
what the reference computes per step: `current_timestep = min(t_i, 1 - conditioning_mask)` (pipeline_ltx_video.py
:1143-1171) and dt = that row of the FIRST batch entry minus the next lower value of the schedule (rf.py:343-360 called on
`current_timestep[:1]`, pipeline :1262)."""
import pytest
import torch

import ref_block as rb

from b200_ltx.sampling import _per_step, _step_tables
from b200_ltx.lib import B200Error


@pytest.mark.parametrize("sampler", ["uniform", "linear_quadratic"])
@pytest.mark.parametrize("steps", [1, 2, 7, 40])
def test_step_tables_match_the_scheduler_step(sampler, steps):
    grid = rb.uniform_timesteps(steps) if sampler == "uniform" else rb.linear_quadratic_timesteps(steps)
    B, N = 3, 57
    g = torch.Generator().manual_seed(steps)
    mask = torch.rand(B, N, generator=g)
    mask[:, :5] = 1.0            # hard conditioning: timestep 0, never denoised
    mask[:, 5:9] = 0.0           # free tokens
    mask[1, 9:20] = 1.0 - grid[min(2, steps - 1)]    # exactly ON a grid value
    for cm in (None, mask):
        rows, dts = _step_tables(grid, cm)
        T = 1 if cm is None else N
        assert rows.shape == (steps, 1 if cm is None else B, T) and dts.shape == (steps, T)
        assert rows.dtype == torch.float32 and dts.dtype == torch.float32
        for i in range(steps):
            cur = grid[i].view(1, 1).expand(B, 1) if cm is None else torch.minimum(grid[i].view(1, 1), 1.0 - cm)
            assert torch.equal(rows[i].expand(B, T), cur.expand(B, T))
            # dt through the scheduler restatement (bit-equal to the reference's, tests/test_oracle.py): with v = 1 and
            # x = 0 the Euler step returns -dt
            first = cur[:1].expand(1, T).contiguous()
            want = -rb.rf_step(grid, torch.ones(1, T, 1), first, torch.zeros(1, T, 1))[0, :, 0]
            assert torch.equal(dts[i], want), (i, (dts[i] - want).abs().max())
        if cm is not None:
            assert bool((dts[:, :5] == 0).all())         # hard-conditioned tokens never move
    # the last step always lands on 0
    rows, dts = _step_tables(grid, None)
    torch.testing.assert_close(rows[-1, 0, 0] - dts[-1, 0], torch.tensor(0.0), rtol=0, atol=1e-7)


def test_per_step_guidance_lists():
    assert _per_step(3.0, 4, "guidance_scale") == [3.0] * 4
    assert _per_step([1, 2, 3], 3, "stg_scale") == [1.0, 2.0, 3.0]
    with pytest.raises(B200Error, match="rescaling_scale has 2 entries for 3 steps"):
        _per_step((1.0, 0.7), 3, "rescaling_scale")
