"""Debug helper: which part of the train step breaks CUDA-graph capture (fwd | bwd | all), and does any C-ABI
launch arrive with the legacy stream?   python tools/graph_debug.py fwd|bwd|all"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from b200_ltx import api, lora, ops, train

dev = torch.device("cuda", 0)
cfg = dict(api.LTXV_2B_CONFIG, num_layers=2)
torch.manual_seed(0)
model = lora.apply_training_strategy(api.build_model(cfg, device=dev), 32, 32).train()
opt = torch.optim.AdamW([p for p in model.parameters() if p.requires_grad], lr=1e-4, fused=True, capturable=True)
B, F, H, W = 1, 4, 8, 16
batch = {"latents": torch.randn(B, 128, F, H, W).bfloat16().to(dev),
         "pose_latents": torch.randn(B, 128, F, H, W).bfloat16().to(dev),
         "ref_image_latents": torch.randn(B, 128, 1, H, W).bfloat16().to(dev)}
prompt = torch.randn(1, 256, 4096).bfloat16().to(dev)
mask = torch.ones(1, 256, dtype=torch.long, device=dev)


class Cfg:
    rf_log_normal_mu, rf_log_normal_sigma, rf_quantile_min, rf_quantile_max, transformer_loss_weight = -0.5, 1.0, 0.005, 0.999, 1.0


orig = ops._call


def traced(family, work, unit, cfn, *args, launches=1):
    if not args[-1]:
        print("LEGACY STREAM in", family, cfn.__name__, flush=True)
    return orig(family, work, unit, cfn, *args, launches=launches)


ops._call = traced


def fwd():
    return train.train_step(model, batch, api.RectifiedFlowScheduler(), api.SymmetricPatchifier(1), Cfg, prompt, mask,
                            device=dev)[0]


mode = sys.argv[1]
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(3):
        opt.zero_grad(set_to_none=True)
        l = fwd()
        if mode != "fwd":
            l.backward()
        if mode == "all":
            opt.step()
torch.cuda.current_stream().wait_stream(s)
torch.cuda.synchronize()
opt.zero_grad(set_to_none=True)
g = torch.cuda.CUDAGraph()
try:
    with torch.cuda.graph(g):
        l = fwd()
        if mode != "fwd":
            l.backward()
        if mode == "all":
            opt.step()
    g.replay()
    torch.cuda.synchronize()
    print(mode, "capture OK, loss", float(l))
except Exception as e:
    print(mode, "capture FAILED:", str(e).splitlines()[0])
