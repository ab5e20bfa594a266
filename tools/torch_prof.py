"""Kernel-time table of ONE train step at cfg2 through torch.profiler (CUPTI): which launches are left outside
the big three families.  python tools/torch_prof.py > gpurun_out/torch_prof.log"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile
from b200_ltx import api, lora, train

dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = lora.apply_training_strategy(api.build_model(dict(api.LTXV_2B_CONFIG), device=dev), 32, 32)
g = torch.Generator().manual_seed(1)
for n, p in model.named_parameters():
    if "lora_B" in n:
        p.data.copy_(torch.randn(p.shape, generator=g) * 0.02)
model.train()
opt = torch.optim.AdamW([p for p in model.parameters() if p.requires_grad], lr=1e-4, fused=True)
B, F, H, W = 1, 16, 16, 24
batch = {"latents": torch.randn(B, 128, F, H, W).bfloat16().to(dev), "pose_latents": torch.randn(B, 128, F, H, W).bfloat16().to(dev),
         "ref_image_latents": torch.randn(B, 128, 1, H, W).bfloat16().to(dev)}
prompt = torch.randn(1, 256, 4096).bfloat16().to(dev)
mask = torch.ones(1, 256, dtype=torch.long, device=dev)
mask[:, 15:] = 0


class Cfg:
    rf_log_normal_mu, rf_log_normal_sigma, rf_quantile_min, rf_quantile_max, transformer_loss_weight = -0.5, 1.0, 0.005, 0.999, 1.0


def step():
    opt.zero_grad(set_to_none=True)
    loss, *_ = train.train_step(model, batch, api.RectifiedFlowScheduler(), api.SymmetricPatchifier(1), Cfg, prompt, mask, device=dev)
    loss.backward()
    opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step()
    torch.cuda.synchronize()
rows = [(e.key, e.device_time_total if hasattr(e, "device_time_total") else e.cuda_time_total, e.count)
        for e in prof.key_averages() if (getattr(e, "device_time_total", 0) or getattr(e, "cuda_time_total", 0)) > 0]
kern = [r for r in rows if "kernel" in r[0].lower() or "b200::" in r[0] or "Memcpy" in r[0] or "Memset" in r[0]]
tot = sum(r[1] for r in kern)
print(f"total device kernel time {tot / 1e3:.2f} ms over {sum(r[2] for r in kern)} launches")
for k, t, c in sorted(kern, key=lambda r: -r[1])[:40]:
    print(f"{t / 1e3:8.3f} ms {100 * t / tot:5.1f}%  n={c:4d}  avg {t / c:8.1f} us  {k[:110]}")

# which aten ops (with input shapes) launch the torch-side copy / fill / add kernels
if os.environ.get("B200_PROF_OPS"):
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], record_shapes=True) as prof2:
        step()
        torch.cuda.synchronize()
    print("\naten ops with device time, grouped by input shapes:")
    ops_rows = [e for e in prof2.key_averages(group_by_input_shape=True) if e.key.startswith("aten::")
                and (getattr(e, "self_device_time_total", 0) or getattr(e, "self_cuda_time_total", 0)) > 0]
    for e in sorted(ops_rows, key=lambda e: -(getattr(e, "self_device_time_total", 0) or getattr(e, "self_cuda_time_total", 0)))[:40]:
        t = getattr(e, "self_device_time_total", 0) or getattr(e, "self_cuda_time_total", 0)
        print(f"{t / 1e3:8.3f} ms  n={e.count:4d}  {e.key:28s} {str(e.input_shapes)[:150]}")
