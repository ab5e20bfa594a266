"""Model-level parity: the b200 path (mirror modules -> C ABI kernels) against the oracle
(oracle/ref_block.py) on identical weights and inputs.

Tolerance: with err = relative Frobenius error against the fp32 oracle,
    E_ours <= max(2 * E_ref, floor)
where E_ref is the error of the *reference's own bf16 dtype flow* (the oracle run on bf16 tensors)
and the floors are 2e-2 for the velocity output / loss and 3e-2 for the gradients.  max-abs errors are
printed beside every relative error."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import ref_block as rb  # noqa: E402

OUT_FLOOR, GRAD_FLOOR = 2e-2, 3e-2


def rel(a, b):
    a, b = a.float(), b.float()
    return float((a - b).norm() / (b.norm() + 1e-20))


def maxabs(a, b):
    return float((a.float() - b.float()).abs().max())


def build_b200_model(cfg, P, lora_rank, device="cuda", train_mode="lora_audio"):
    from b200_ltx import api, lora
    full = dict(api.LTXV_2B_CONFIG)
    full.update(cfg)
    model = api.build_model(full, device=device)
    if train_mode != "lora_audio":
        model = lora.apply_training_strategy(model, 0, 0, train_mode=train_mode)
    elif lora_rank:
        model = lora.apply_training_strategy(model, lora_rank, lora_rank)
    sd = model.state_dict()
    mapped = {}
    for k, v in sd.items():
        plain = k.replace("base_model.model.", "").replace(".base_layer.", ".")
        mapped[k] = P[plain].to(device=v.device, dtype=v.dtype)
    model.load_state_dict(mapped, strict=True)
    return model


def oracle_loss_grads(P, cfg, batch, t, dtype, device, train_mode="lora_audio"):
    Pd = {}
    for k, v in P.items():
        w = v.to(device=device, dtype=torch.float32 if ("lora_" in k or dtype == torch.float32) else dtype)
        Pd[k] = w.clone().requires_grad_(rb.is_trainable(k, train_mode))
    b = {k: v.to(device) for k, v in batch.items()}
    loss, out = rb.train_step_loss(Pd, cfg, b["latents"].to(dtype), b["ref_image_latents"].to(dtype),
                                   b["pose_latents"].to(dtype), b["prompt_embeds"].to(dtype), b["prompt_mask"],
                                   t.to(device), b["noise"].to(dtype))
    loss.backward()
    grads = {k: v.grad.detach() for k, v in Pd.items() if v.grad is not None}
    return loss.detach(), out.detach(), grads


def b200_loss_grads(model, batch, t, device="cuda"):
    from b200_ltx import api, train
    dev = device

    class Cfg:
        transformer_loss_weight = 1.0
    b = {k: v.to(dev) for k, v in batch.items()}
    noise = b["noise"].to(torch.bfloat16)
    model.zero_grad(set_to_none=True)
    loss, rel_mse, nrmse, ld = train.train_step(model, b, api.RectifiedFlowScheduler(), api.SymmetricPatchifier(1),
                                                Cfg(), b["prompt_embeds"], b["prompt_mask"], device=dev,
                                                t=t.to(dev), noise=noise)
    loss.backward()
    grads = {}
    for n, p in model.named_parameters():
        if p.grad is not None:
            grads[n.replace("base_model.model.", "").replace(".base_layer.", ".")] = p.grad.detach()
    return loss.detach(), grads


def run_parity(cfg, case, verbose=True):
    """Returns a dict of (E_ours, E_ref, maxabs) per quantity; asserts the two-sided tolerance."""
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    dev = "cuda"
    P = rb.init_params(cfg, case["lora_rank"], seed=case["seed_w"])
    batch = rb.synthetic_batch(cfg, case["b"], case["f"], case["h"], case["w"], case["n_ctx"], case["seed_x"],
                               case.get("valid_ctx"))
    # the b200 path consumes bf16 inputs; give every implementation the same bf16-rounded data
    for k in ("latents", "pose_latents", "ref_image_latents", "prompt_embeds", "noise"):
        batch[k] = batch[k].to(torch.bfloat16).float()
    P = {k: (v if "lora_" in k else v.to(torch.bfloat16).float()) for k, v in P.items()}
    t = torch.tensor(case["t"])
    mode = case.get("train_mode", "lora_audio")
    l32, o32, g32 = oracle_loss_grads(P, cfg, batch, t, torch.float32, dev, mode)
    l16, o16, g16 = oracle_loss_grads(P, cfg, batch, t, torch.bfloat16, dev, mode)
    model = build_b200_model(cfg, P, case["lora_rank"], dev, mode)
    out_holder = {}
    root = model.base_model.model if hasattr(model, "base_model") else model
    hook = root.register_forward_hook(lambda m, a, o: out_holder.__setitem__("out", o.sample.detach()))
    lb, gb = b200_loss_grads(model, batch, t)
    hook.remove()
    torch.cuda.synchronize()
    res = {}

    def record(name, ours, ref16, ref32, floor):
        e_o, e_r = rel(ours, ref32), rel(ref16, ref32)
        res[name] = (e_o, e_r, maxabs(ours, ref32), maxabs(ref16, ref32))
        ok = e_o <= max(2 * e_r, floor)
        if verbose:
            print(f"  {name:58s} E_ours={e_o:.3e} (maxabs {res[name][2]:.2e})  E_ref_bf16={e_r:.3e} "
                  f"(maxabs {res[name][3]:.2e}) {'ok' if ok else 'FAIL'}", flush=True)
        return ok
    ok = record("velocity output", out_holder["out"], o16, o32, OUT_FLOOR)
    ok &= record("loss", lb, l16, l32, OUT_FLOOR)
    assert set(gb) == set(g32), (set(gb) ^ set(g32))
    for k in sorted(g32):
        ok &= record("grad " + k, gb[k], g16[k], g32[k], GRAD_FLOOR)
    assert ok, "b200 path outside the two-sided bf16 tolerance"
    return res
