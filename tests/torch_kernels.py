"""Plain-torch stand-ins for the raw `ops.*` kernel wrappers, with the same Python signatures.

TEST INFRASTRUCTURE ONLY: lets the CPU suite run the HOST logic of the product (modules.transformer_forward /
block_forward / attention_forward, api.install on the reference's own classes, the LoRA attribute plumbing, the
caches) without a GPU.  Nothing in the package imports this file; on a GPU box the tests call the real C ABI."""
import contextlib

import torch
import torch.nn.functional as F

BF16 = torch.bfloat16


def _mat(t, rows_are_k):
    return t.float().t() if rows_are_k else t.float()


def _contract(ok, what):
    if not ok:
        raise AssertionError("b200_gemm_bf16 would reject this call (csrc/gemm.cu gemm_impl): " + what)


def _aligned(t):
    return t is None or (t.storage_offset() * t.element_size()) % 16 == 0


def _gemm_contract(a, b, a_mn, b_mn, a2, b2, out, out_dtype, bias, gate, res, aux, groups_block=None):
    """The argument contract of the C ABI (gemm.cu: gemm_impl), so that host code which would be refused on the device
    fails here too: contiguous extents in multiples of 8 elements, 16-byte aligned pointers and pitches."""
    M, K = (a.shape[1], a.shape[0]) if a_mn else a.shape
    N = b.shape[1] if b_mn else b.shape[0]
    K2 = 0 if a2 is None else (a2.shape[0] if a_mn else a2.shape[1])
    if M == 0 or N == 0:
        return
    _contract(N % 8 == 0, f"N = {N} is not a multiple of 8")
    if not a_mn or not b_mn:
        _contract(K % 8 == 0 and K2 % 8 == 0, f"K = {K} / K2 = {K2} must be multiples of 8 for a K-major operand")
    if a_mn:
        _contract(M % 8 == 0, f"M = {M} must be a multiple of 8 for a [K, M] A operand")
    f32 = (out.dtype if out is not None else out_dtype) == torch.float32
    for t, nm in ((a, "A"), (b, "B"), (a2, "A2"), (b2, "B2"), (res, "res"), (aux, "aux"), (gate, "gate")):
        if t is not None:
            _contract(t.stride(1) == 1 and t.stride(0) % 8 == 0 and _aligned(t), f"{nm}: pitch {t.stride()} / alignment")
    if out is not None:
        _contract(out.stride(1) == 1 and out.stride(0) % (4 if f32 else 8) == 0 and _aligned(out), "C: pitch / alignment")
    _contract(_aligned(bias), "bias alignment")
    if groups_block is not None:
        bn, Mg, Ng, Kg, K2g = groups_block
        _contract(Mg % 128 == 0 and Ng % bn == 0 and Kg % 64 == 0 and K2g % 64 == 0,
                  f"batched: per-group M, N, K, K2 = {Mg}, {Ng}, {Kg}, {K2g} vs 128, {bn}, 64, 64")


def gemm(a, b, *, a_rows_are_k=False, b_rows_are_k=False, a2=None, b2=None, out=None, out_dtype=BF16, bias=None,
         gate=None, rows_per_gate=0, res=None, aux=None, epilogue=0, block_n=0, split_k=1, _checked=False):
    if not _checked:
        _gemm_contract(a, b, a_rows_are_k, b_rows_are_k, a2, b2, out, out_dtype, bias, gate, res, aux)
    acc = _mat(a, a_rows_are_k) @ _mat(b, b_rows_are_k).t()
    if a2 is not None:
        acc = acc + _mat(a2, a_rows_are_k) @ _mat(b2, b_rows_are_k).t()
    if bias is not None:
        acc = acc + bias.float()
    if epilogue == 1:
        if aux is not None:
            aux.copy_(acc.to(aux.dtype))
        acc = F.gelu(acc, approximate="tanh")
    elif epilogue == 2:
        h = aux.float().requires_grad_(True)
        with torch.enable_grad():
            F.gelu(h, approximate="tanh").sum().backward()
        acc = acc * h.grad
    if gate is not None:
        acc = acc * gate.float().repeat_interleave(rows_per_gate, 0)[:acc.shape[0]]
    if res is not None:
        acc = acc + res.float()
    if out is None:
        return acc.to(out_dtype)
    out.copy_(acc.to(out.dtype))
    return out


def gemm_batched(a, b, out, M, N, K, groups, offs, *, a_rows_are_k=False, b_rows_are_k=False, a2=None, b2=None, K2=0,
                 bias=None, block_n=0):
    z = (0, 0)
    bn = block_n or (256 if N >= 256 and N % 256 == 0 else (128 if N % 128 == 0 else 64))
    _gemm_contract(a, b, a_rows_are_k, b_rows_are_k, a2, b2, out, out.dtype, bias, None, None, None, (bn, M, N, K, K2))

    def sub(t, name, rows, cols, g):
        r0, c0 = offs.get(name, z)
        return t[g * r0:g * r0 + rows, g * c0:g * c0 + cols]
    for g in range(groups):
        ag = sub(a, "a", K if a_rows_are_k else M, M if a_rows_are_k else K, g)
        bg = sub(b, "b", K if b_rows_are_k else N, N if b_rows_are_k else K, g)
        a2g = b2g = None
        if a2 is not None:
            a2g = sub(a2, "a2", K2 if a_rows_are_k else M, M if a_rows_are_k else K2, g)
            b2g = sub(b2, "b2", K2 if b_rows_are_k else N, N if b_rows_are_k else K2, g)
        bs = None
        if bias is not None:
            o = offs.get("bias", 0)
            bs = bias[g * o:g * o + N]
        gemm(ag, bg, a_rows_are_k=a_rows_are_k, b_rows_are_k=b_rows_are_k, a2=a2g, b2=b2g, bias=bs,
             out=sub(out, "c", M, N, g), _checked=True)
    return out


def norm_mod_fwd(x, scale, shift, rows_per_mod, eps, layernorm=False, out=None):
    xf = x.float()
    if layernorm:
        n = F.layer_norm(xf, (xf.shape[-1],), None, None, eps)
    else:
        n = xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)
    rows = x.shape[0]
    if scale is not None:
        n = n * (1 + scale.float().repeat_interleave(rows_per_mod, 0)[:rows])
    if shift is not None:
        n = n + shift.float().repeat_interleave(rows_per_mod, 0)[:rows]
    y = n.to(BF16)
    if out is not None:
        out.copy_(y)
        return out
    return y


def _rope(x, cos, sin):
    xr = x.reshape(x.shape[0], -1, 2)
    rot = torch.stack((-xr[..., 1], xr[..., 0]), dim=-1).reshape(x.shape)
    return x * cos.float() + rot * sin.float()


def qknorm_rope_fwd(xq, xk, wq, wk, cos, sin, oq, ok, eps=1e-5):
    for x, w, o in ((xq, wq, oq), (xk, wk, ok)):
        if x is None:
            continue
        xf = x.float()
        n = (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)).to(BF16).float() * w.float()
        if cos is not None:
            n = _rope(n.to(BF16).float(), cos, sin)
        o.copy_(n.to(o.dtype))


def fa_fwd(q, k, v, B, H, Nq, Nk, key_bias=None, scale=0.125, need_lse=True, attn1=None, batch_keep=None,
           pass_src=None):
    D = H * 64
    qh = q[:, :D].float().reshape(B, Nq, H, 64).transpose(1, 2)
    kh = k[:, :D].float().reshape(B, Nk, H, 64).transpose(1, 2)
    vh = v[:, :D].float().reshape(B, Nk, H, 64).transpose(1, 2)
    s = qh @ kh.transpose(-1, -2) * scale
    if key_bias is not None:
        s = s + key_bias.float()[:, None, None, :]
    lse = torch.logsumexp(s, dim=-1)
    o = (torch.softmax(s, dim=-1) @ vh).transpose(1, 2).reshape(B * Nq, D).to(BF16)
    if batch_keep is not None:
        src = (pass_src if pass_src is not None else v)[:, :D].reshape(B, Nq, D)
        o = torch.where(batch_keep.reshape(B, 1, 1) == 0, src, o.view(B, Nq, D)).reshape(B * Nq, D)
    return o, (lse if need_lse else None)


def lerp_condition_(tokens, ref, pose, w_ref=0.85, w_pose=0.5, token_offset=0):
    """In-place conditioning of transformer3d.py:447-466 on a contiguous shard [token_offset, token_offset + N) of the
    clip's tokens: frame 0 lerps towards the reference image (0.85), later frames towards the pose latents (0.5)."""
    B, N, C = tokens.shape
    Fr, Hh, Ww = pose.shape[2], pose.shape[3], pose.shape[4]
    HW = Hh * Ww
    assert 0 <= token_offset and token_offset + N <= Fr * HW
    # whole-clip conditioning tokens [B, F*HW, C]: frame 0 from `ref`, the rest from `pose`
    cond = pose.reshape(B, C, Fr * HW).transpose(1, 2).clone()
    cond[:, :HW] = ref.reshape(B, C, HW).transpose(1, 2)
    w = torch.full((Fr * HW, 1), w_pose)
    w[:HW] = w_ref
    sl = slice(token_offset, token_offset + N)
    tokens.copy_(torch.lerp(tokens.float(), cond[:, sl].float(), w[sl]).to(tokens.dtype))
    return tokens


# ---- backward kernels: fp32 autograd through the same math, rounded where the kernels round ----
def norm_mod_bwd(dy, x, scale, rows_per_mod, eps, layernorm=False, dres=None, want_prod=False):
    rows = x.shape[0]
    xf = x.float().detach().requires_grad_(True)
    with torch.enable_grad():
        if layernorm:
            n = F.layer_norm(xf, (xf.shape[-1],), None, None, eps)
        else:
            n = xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)
        y = n if scale is None else n * (1 + scale.detach().float().repeat_interleave(rows_per_mod, 0)[:rows])
        y.backward(dy.float())
    dx = xf.grad if dres is None else xf.grad + dres.float()
    dx = dx.to(BF16)
    return (dx, (dy.float() * n.detach()).to(BF16)) if want_prod else dx


def colsum_groups(a, b=None, rows_per_group=0):
    rows, N = a.shape
    rpg = rows_per_group or rows
    if rows == 0:
        return torch.empty((0, N), dtype=torch.float32)
    assert rows % rpg == 0, "colsum_groups: rows must be a multiple of rows_per_group"
    v = a.float() if b is None else a.float() * b.float()
    return v.view(rows // rpg, rpg, N).sum(1)


def colsum(x):
    return x.float().sum(0)


def rowscale(x, g, rows_per_mod):
    return (x.float() * g.float().repeat_interleave(rows_per_mod, 0)[:x.shape[0]]).to(BF16)


def qknorm_rope_bwd(dq, dk, xq, xk, wq, wk, cos, sin, oq, ok, eps=1e-5, prod_q=None, prod_k=None):
    for d, x, w, o, prod in ((dq, xq, wq, oq, prod_q), (dk, xk, wk, ok, prod_k)):
        if x is None:
            continue
        xf = x.float().detach().requires_grad_(True)
        wf = w.float().detach().requires_grad_(True)
        with torch.enable_grad():
            xhat = xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)
            n = xhat * wf
            if cos is not None:
                n = _rope(n, cos, sin)
            n.backward(d.float())
        o.copy_(xf.grad.to(o.dtype))
        if prod is not None:   # (RoPE^T d) * xhat: its column sum is the gradient of the norm weight
            with torch.enable_grad():
                z = torch.zeros_like(xf, requires_grad=True)
                (_rope(z, cos, sin) if cos is not None else z).backward(d.float())
            prod.copy_((z.grad * xhat.detach()).to(prod.dtype))


def fa_bwd(q, k, v, o, do, lse, B, H, Nq, Nk, dk, dv, key_bias=None, scale=0.125, delta=None, dq_accum=None, attn1=None):
    D = H * 64
    leaves = []
    for t, n in ((q, Nq), (k, Nk), (v, Nk)):
        leaves.append(t[:, :D].float().detach().reshape(B, n, H, 64).transpose(1, 2).requires_grad_(True))
    qh, kh, vh = leaves
    with torch.enable_grad():
        s = qh @ kh.transpose(-1, -2) * scale
        if key_bias is not None:
            s = s + key_bias.float()[:, None, None, :]
        out = (torch.softmax(s, dim=-1) @ vh).transpose(1, 2).reshape(B * Nq, D)
        out.backward(do[:, :D].float())
    flat = [t.grad.transpose(1, 2).reshape(-1, D) for t in leaves]
    dk.copy_(flat[1].to(dk.dtype))
    dv.copy_(flat[2].to(dv.dtype))
    if dq_accum is not None:
        dq_accum[:, :D] += flat[0]
        return dq_accum
    return flat[0].contiguous()


def rf_noise(x0, noise, t, want_xt=True, want_v=True):
    """x_t = (1 - t) x0 + t eps and v = eps - x0 in fp32, rounded once (csrc/elementwise.cu rf_noise_kernel)."""
    if x0.dtype != BF16 or noise.dtype != BF16 or not x0.is_contiguous() or not noise.is_contiguous():
        raise AssertionError("rf_noise: contiguous bf16 tensors required")
    tt = t.float().reshape(-1, *([1] * (x0.dim() - 1)))
    xt = ((1 - tt) * x0.float() + tt * noise.float()).to(BF16) if want_xt else None
    v = (noise.float() - x0.float()).to(BF16) if want_v else None
    return xt, v


def rf_loss(out, target, grad_scale=1.0, want_grad=True):
    """mean((out - target)^2) in fp32 and its gradient 2 (out - target) / numel * grad_scale in bf16."""
    d = out.float() - target.float()
    loss = (d * d).mean()
    dout = (d * (2.0 * grad_scale / d.numel())).to(BF16) if want_grad else None
    return loss, dout


def guidance_step_(v, x, x_next, dt, noise_level, scalars, has_cfg, has_stg, cfg_star=False, rescale=False,
                   workspace=None):
    """The element-wise tail of a sampling step (csrc/guidance.cu), in place on the fp32 latents: guidance combine as the
    oracle restates pipeline_ltx_video.py:1217-1260, Euler update, conditioning select (:1346-1379), and the next
    step's bf16 model input, one copy per condition."""
    import ref_sampling as rs
    B, N, C = x.shape
    conds = 1 + int(bool(has_cfg)) + int(bool(has_stg))
    gs, stg, rsc, t = [float(s) for s in scalars[:4]]
    pred = rs.guidance_combine(v.float(), B, conds, bool(has_cfg), bool(has_stg), gs, stg, rsc if rescale else 1.0,
                               bool(cfg_star))
    den = x - dt.reshape(1, -1, 1) * pred
    if noise_level is not None:
        den = torch.where((t - 1e-6 < noise_level).unsqueeze(-1), den, x)
    x.copy_(den)
    if x_next is not None:
        x_next.copy_(x.to(BF16).repeat(x_next.shape[0] // B, 1, 1))
    return x


class AsIfOnDevice(torch.Tensor):
    """A CPU tensor that answers `is_cuda` with True: lets entry points that insist on device tensors (sampling.Denoiser)
    run their host logic over the stand-ins."""
    is_cuda = property(lambda self: True)


@contextlib.contextmanager
def patched():
    """Run the product's host logic over these stand-ins (CPU tensors allowed, no device check)."""
    from b200_ltx import lib, modules, ops
    names = ["gemm", "gemm_batched", "norm_mod_fwd", "qknorm_rope_fwd", "fa_fwd", "lerp_condition_", "rf_noise", "rf_loss",
             "norm_mod_bwd", "colsum_groups", "colsum", "rowscale", "qknorm_rope_bwd", "fa_bwd", "guidance_step_"]
    saved = {n: getattr(ops, n) for n in names}
    saved_req, saved_dev = modules._require_bf16, lib.require_device
    try:
        for n in names:
            setattr(ops, n, globals()[n])

        def req(t, what):
            if t.dtype != BF16:
                raise lib.B200Error(f"{what} must be bfloat16")
        modules._require_bf16 = req
        lib.require_device = lambda: None
        yield
    finally:
        for n, f in saved.items():
            setattr(ops, n, f)
        modules._require_bf16, lib.require_device = saved_req, saved_dev
