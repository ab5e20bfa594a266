"""Per-kernel parity checks (CUDA path through the C ABI vs a plain torch fp32 reference of the
same op).  Used by tests/test_kernels_gpu.py and by tools/gpu_selftest.py (which runs each group
in its own process so that one faulting kernel cannot hide the others)."""
import math

import torch
import torch.nn.functional as F

BF16 = torch.bfloat16


def _rel(a, b):
    a, b = a.float(), b.float()
    return float((a - b).norm() / (b.norm() + 1e-12))


def _maxabs(a, b):
    return float((a.float() - b.float()).abs().max())


def _assert_close(name, got, ref, rel=1e-2, info=""):
    r, m = _rel(got, ref), _maxabs(got, ref)
    ok = r <= rel and math.isfinite(r)
    print(f"  {name:46s} rel={r:.3e} maxabs={m:.3e} {'ok' if ok else 'FAIL'} {info}", flush=True)
    assert ok, f"{name}: rel {r} > {rel} (maxabs {m}) {info}"


def _randn(*shape, seed=0, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to("cuda", BF16)


# --------------------------------------------------------------------------------------------
def check_gemm_layouts():
    from b200_ltx import ops
    cases = [(128, 256, 64), (256, 512, 256), (200, 136, 72), (96, 64, 128), (640, 2048, 512), (1, 2048, 256),
             (6144 // 4, 768, 1024)]
    for bn in (0, 64, 128, 256):
        for (M, N, K) in cases:
            for a_k in (False, True):
                if a_k and M % 8:
                    continue
                for b_k in (False, True):
                    a = _randn(M, K, seed=1)
                    b = _randn(N, K, seed=2)
                    ref = a.float() @ b.float().t()
                    aa = a.t().contiguous() if a_k else a
                    bb = b.t().contiguous() if b_k else b
                    out = ops.gemm(aa, bb, a_rows_are_k=a_k, b_rows_are_k=b_k, block_n=bn)
                    _assert_close(f"gemm {M}x{N}x{K} bn={bn} aK={int(a_k)} bK={int(b_k)}", out, ref, 6e-3)
    # wgrad form with a ragged reduction (token count not a multiple of 8, nor of the 64-row k block)
    for (M, N, K) in [(64, 2048, 210), (2048, 64, 105), (128, 136, 3)]:
        a, b = _randn(K, M, seed=3), _randn(K, N, seed=4)
        out = ops.gemm(a, b, a_rows_are_k=True, b_rows_are_k=True, out_dtype=torch.float32)
        _assert_close(f"gemm [K,M]x[K,N] ragged K {M}x{N}x{K}", out, a.float().t() @ b.float(), 6e-3)
    torch.cuda.synchronize()


def check_gemm_batched():
    """Strided-batch GEMM (one launch, G sub-block problems) vs per-group torch matmuls, in the four operand
    arrangements CtxKVFn uses."""
    from b200_ltx import ops
    G, M, D, Dc, R = 5, 256, 512, 320, 64
    x = _randn(M, Dc, seed=1, scale=0.5)
    W = _randn(G * D, Dc, seed=2, scale=0.1)
    bias = _randn(G * D, seed=3)
    t = _randn(M, G * R, seed=4, scale=0.3)
    sB = _randn(G * D, R, seed=5, scale=0.1)
    kv = torch.empty(M, G * D, device="cuda", dtype=BF16)
    ops.gemm_batched(x, W, kv, M, D, Dc, G, {"b": (D, 0), "a2": (0, R), "b2": (D, 0), "c": (0, D), "bias": D}, a2=t, b2=sB,
                     K2=R, bias=bias)
    ref = torch.cat([x.float() @ W[g * D:(g + 1) * D].float().T + bias[g * D:(g + 1) * D].float()
                     + t[:, g * R:(g + 1) * R].float() @ sB[g * D:(g + 1) * D].float().T for g in range(G)], dim=1)
    _assert_close("batched fwd (bias + second pair)", kv, ref, 6e-3)
    dkv = _randn(M, G * D, seed=6, scale=0.2)
    dt = torch.empty(M, G * R, device="cuda", dtype=BF16)
    ops.gemm_batched(dkv, sB, dt, M, R, D, G, {"a": (0, D), "b": (D, 0), "c": (0, R)}, b_rows_are_k=True, block_n=64)
    ref = torch.cat([dkv[:, g * D:(g + 1) * D].float() @ sB[g * D:(g + 1) * D].float() for g in range(G)], dim=1)
    _assert_close("batched dt (A cols / B rows displaced)", dt, ref, 6e-3)
    dB = torch.empty(G * D, R, device="cuda", dtype=torch.float32)
    ops.gemm_batched(dkv, t, dB, D, R, M, G, {"a": (0, D), "b": (0, R), "c": (D, 0)}, a_rows_are_k=True, b_rows_are_k=True,
                     block_n=64)
    ref = torch.cat([dkv[:, g * D:(g + 1) * D].float().T @ t[:, g * R:(g + 1) * R].float() for g in range(G)], dim=0)
    _assert_close("batched dB ([K,M] x [K,N], fp32 out, rows displaced)", dB, ref, 6e-3)
    torch.cuda.synchronize()


def check_gemm_epilogues():
    from b200_ltx import ops
    M, N, K, K2 = 384, 512, 256, 64
    a, b = _randn(M, K, seed=1), _randn(N, K, seed=2, scale=0.05)
    a2, b2 = _randn(M, K2, seed=3), _randn(N, K2, seed=4, scale=0.05)
    bias = _randn(N, seed=5)
    gate = _randn(3, 6 * N, seed=6)[:, 2 * N:3 * N]          # strided view, like ada[:, 2]
    res = _randn(M, N, seed=7)
    base = a.float() @ b.float().t()
    _assert_close("gemm dual-K", ops.gemm(a, b, a2=a2, b2=b2), base + a2.float() @ b2.float().t(), 6e-3)
    _assert_close("gemm bias", ops.gemm(a, b, bias=bias), base + bias.float(), 6e-3)
    ref = gate.float().repeat_interleave(128, 0) * (base + bias.float()) + res.float()
    _assert_close("gemm bias+gate+res", ops.gemm(a, b, bias=bias, gate=gate, rows_per_gate=128, res=res), ref, 6e-3)
    pre = torch.empty(M, N, device="cuda", dtype=BF16)
    act = ops.gemm(a, b, bias=bias, epilogue=ops.EPI_GELU, aux=pre)
    pre_ref = (base + bias.float())
    _assert_close("gemm gelu pre-activation", pre, pre_ref, 6e-3)
    _assert_close("gemm gelu", act, F.gelu(pre.float(), approximate="tanh"), 6e-3)
    h = pre.float().requires_grad_(True)
    F.gelu(h, approximate="tanh").sum().backward()
    _assert_close("gemm gelu'", ops.gemm(a, b, epilogue=ops.EPI_GELU_GRAD, aux=pre), base * h.grad, 8e-3)
    o32 = ops.gemm(a, b, out_dtype=torch.float32)
    _assert_close("gemm fp32 out", o32, base, 1e-5 + 6e-3)
    # split-K (skinny wgrad shapes): [K,M] x [K,N] with a long reduction, fp32 atomics into zeros
    for (Ms, Ns, Ks, sp) in [(64, 2048, 6144, 0), (2048, 64, 6144, 0), (256, 256, 1000, 3), (128, 128, 64, 4)]:
        at, bt = _randn(Ks, Ms, seed=21, scale=0.1), _randn(Ks, Ns, seed=22, scale=0.1)
        got = ops.gemm(at, bt, a_rows_are_k=True, b_rows_are_k=True, out_dtype=torch.float32, split_k=sp)
        _assert_close(f"gemm split-K {Ms}x{Ns}x{Ks} sp={sp}", got, at.float().t() @ bt.float(), 6e-3)
    acc = res.clone()
    ops.gemm(a, b, out=acc, res=acc)
    _assert_close("gemm accumulate in place", acc, base + res.float(), 6e-3)
    # column-slice output + row-strided input (packed qkv buffers)
    buf = torch.zeros(M, 3 * N, device="cuda", dtype=BF16)
    ops.gemm(a, b, out=buf[:, N:2 * N], bias=bias)
    _assert_close("gemm strided out", buf[:, N:2 * N], base + bias.float(), 6e-3)
    assert float(buf[:, :N].abs().max()) == 0 and float(buf[:, 2 * N:].abs().max()) == 0
    wide = _randn(M, 3 * K, seed=9)
    _assert_close("gemm strided in", ops.gemm(wide[:, K:2 * K], b), wide[:, K:2 * K].float() @ b.float().t(), 6e-3)
    # per-row gate (per-token timesteps)
    g2 = _randn(M, N, seed=10)
    _assert_close("gemm per-row gate", ops.gemm(a, b, gate=g2, rows_per_gate=1), g2.float() * base, 6e-3)
    # every compiled variant of the slab epilogue (plain / bias / general x none, GELU, GELU', stash x bias-gate presence)
    # on a shape ragged against every tile size, written into the middle of a sentinel-filled buffer: rows past M and
    # columns past N must stay untouched (there is no compute-sanitizer on this pool)
    Mr, Nr, Kr = 600, 456, 192
    ar, br = _randn(Mr, Kr, seed=31), _randn(Nr, Kr, seed=32, scale=0.05)
    biasr, gater, resr = _randn(Nr, seed=33), _randn(2, Nr, seed=34), _randn(Mr, Nr, seed=35)
    baser = ar.float() @ br.float().t()
    gfull = gater.float().repeat_interleave(300, 0)
    hr = (baser + biasr.float()).to(BF16).float().requires_grad_(True)
    F.gelu(hr, approximate="tanh").sum().backward()
    variants = [
        ("plain", {}, baser, None),
        ("bias", dict(bias=biasr), baser + biasr.float(), None),
        ("res", dict(res=resr), baser + resr.float(), None),
        ("bias+gate+res", dict(bias=biasr, gate=gater, rows_per_gate=300, res=resr),
         gfull * (baser + biasr.float()) + resr.float(), None),
        ("gelu", dict(bias=biasr, epilogue=ops.EPI_GELU), F.gelu(hr.detach(), approximate="tanh"), hr.detach()),
        ("gelu'", dict(epilogue=ops.EPI_GELU_GRAD, aux_in=True), baser * hr.grad, None),
        ("stash", dict(bias=biasr, gate=gater, rows_per_gate=300, res=resr, epilogue=ops.EPI_STASH),
         gfull * (baser + biasr.float()).to(BF16).float() + resr.float(), (baser + biasr.float())),
    ]
    for bn in (128, 256):
        for name, kw, want, want_aux in variants:
            kw = dict(kw)
            big = torch.full((Mr + 40, Nr + 72), 7.0, device="cuda", dtype=BF16)
            aux_big = torch.full((Mr + 40, Nr + 72), 5.0, device="cuda", dtype=BF16)
            if kw.pop("aux_in", False):
                aux_big[8:8 + Mr, 16:16 + Nr] = hr.detach().to(BF16)
                kw["aux"] = aux_big[8:8 + Mr, 16:16 + Nr]
            elif want_aux is not None:
                kw["aux"] = aux_big[8:8 + Mr, 16:16 + Nr]
            ops.gemm(ar, br, out=big[8:8 + Mr, 16:16 + Nr], block_n=bn, **kw)
            _assert_close(f"gemm epilogue variant {name} bn={bn}", big[8:8 + Mr, 16:16 + Nr], want, 8e-3)
            if want_aux is not None:
                _assert_close(f"gemm epilogue variant {name} bn={bn} aux", aux_big[8:8 + Mr, 16:16 + Nr], want_aux, 8e-3)
            for buf, fill in ((big, 7.0), (aux_big, 5.0)):
                if buf is aux_big and "aux" not in kw:
                    continue
                if buf is aux_big and name == "gelu'":
                    continue   # input there
                guard = buf.clone()
                guard[8:8 + Mr, 16:16 + Nr] = fill
                assert bool((guard == fill).all()), f"gemm epilogue variant {name} bn={bn}: wrote outside its output"
    torch.cuda.synchronize()


def check_linear_fn():
    """LinearFn / FeedForwardFn forward + backward vs autograd on the same math in fp32."""
    from b200_ltx import ops
    # LoRA ranks: the train config's 32, the config default 8 (config.py:22), 12 (not a multiple of 8: padded GEMM
    # extents, sliced gradients) and 96 (more than one 64-wide k block)
    for r in (32, 8, 12, 96):
        M, K, N, s = 320, 256, 512, 0.5
        x = _randn(M, K, seed=1).requires_grad_(True)
        W = _randn(N, K, seed=2, scale=0.06).requires_grad_(True)
        b = _randn(N, seed=3, scale=0.1).requires_grad_(True)
        A = (_randn(r, K, seed=4, scale=0.06).float()).requires_grad_(True)
        B = (_randn(N, r, seed=5, scale=0.06).float()).requires_grad_(True)
        gate = _randn(2, N, seed=6)
        res = _randn(M, N, seed=7).requires_grad_(True)
        dy = _randn(M, N, seed=8)
        y = ops.LinearFn.apply(x, W, b, A, B, s, gate, 160, res)
        y.backward(dy)
        xr, Wr, br, Ar, Br, rr = [t.detach().float().requires_grad_(True) for t in (x, W, b, A, B, res)]
        yr = gate.float().repeat_interleave(160, 0) * (xr @ Wr.t() + br + s * (xr @ Ar.t()) @ Br.t()) + rr
        yr.backward(dy.float())
        _assert_close(f"LinearFn r={r} y", y, yr, 8e-3)
        assert A.grad.shape == A.shape and B.grad.shape == B.shape
        for nm, g, gr in (("dx", x.grad, xr.grad), ("dW", W.grad, Wr.grad), ("db", b.grad, br.grad),
                          ("dA", A.grad, Ar.grad), ("dB", B.grad, Br.grad), ("dres", res.grad, rr.grad)):
            _assert_close(f"LinearFn r={r} " + nm, g, gr, 1.5e-2)
    M, K = 320, 256
    # feed-forward
    Dff = 1024
    W1 = _randn(Dff, K, seed=11, scale=0.06).requires_grad_(True)
    b1 = _randn(Dff, seed=12, scale=0.1).requires_grad_(True)
    W2 = _randn(K, Dff, seed=13, scale=0.03).requires_grad_(True)
    b2 = _randn(K, seed=14, scale=0.1).requires_grad_(True)
    x2 = _randn(M, K, seed=15).requires_grad_(True)
    gate2 = _randn(2, K, seed=16)
    res2 = _randn(M, K, seed=17).requires_grad_(True)
    dy2 = _randn(M, K, seed=18)
    y2 = ops.FeedForwardFn.apply(x2, W1, b1, W2, b2, gate2, 160, res2)
    y2.backward(dy2)
    refs = [t.detach().float().requires_grad_(True) for t in (x2, W1, b1, W2, b2, res2)]
    xr, W1r, b1r, W2r, b2r, rr = refs
    yr = gate2.float().repeat_interleave(160, 0) * (F.gelu(xr @ W1r.t() + b1r, approximate="tanh") @ W2r.t() + b2r) + rr
    yr.backward(dy2.float())
    _assert_close("FeedForwardFn y", y2, yr, 8e-3)
    for nm, t, tr in zip(("dx", "dW1", "db1", "dW2", "db2", "dres"), (x2, W1, b1, W2, b2, res2), refs):
        _assert_close("FeedForwardFn " + nm, t.grad, tr.grad, 1.5e-2)
    torch.cuda.synchronize()


def check_norm_mod():
    from b200_ltx import ops
    for (rows, D, rpm, ln) in [(96, 256, 48, False), (515, 2048, 515, False), (130, 2048, 1, False),
                               (96, 256, 96, True), (257, 2048, 257, True), (64, 1024, 16, False)]:
        nb = (rows + rpm - 1) // rpm
        x = _randn(rows, D, seed=1, scale=2.0)
        ada = _randn(nb, 6 * D, seed=2, scale=0.3)
        shift, scale = ada[:, :D], ada[:, D:2 * D]
        eps = 1e-6
        xf = x.float().requires_grad_(True)
        if ln:
            n = F.layer_norm(xf, (D,), None, None, eps)
        else:
            n = xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)
        sc = scale.float().repeat_interleave(rpm, 0)[:rows]
        sh = shift.float().repeat_interleave(rpm, 0)[:rows]
        yr = n * (1 + sc) + sh
        y = ops.norm_mod_fwd(x, scale, shift, rpm, eps, ln)
        _assert_close(f"norm_mod fwd rows={rows} D={D} rpm={rpm} ln={int(ln)}", y, yr, 5e-3)
        dy = _randn(rows, D, seed=3)
        dres = _randn(rows, D, seed=4)
        yr.backward(dy.float())
        dx = ops.norm_mod_bwd(dy, x, scale, rpm, eps, ln, dres=dres)
        _assert_close(f"norm_mod bwd rows={rows} D={D} rpm={rpm} ln={int(ln)}", dx, xf.grad + dres.float(), 6e-3)
    torch.cuda.synchronize()


def _rope_ref(x, cos, sin):
    p = x.unflatten(-1, (-1, 2))
    a, b = p.unbind(-1)
    rot = torch.stack((-b, a), dim=-1).flatten(-2)
    return x * cos + rot * sin


def check_qknorm_rope():
    from b200_ltx import ops
    for (rq, rk, D, rope) in [(96, 96, 256, True), (300, 300, 2048, True), (200, 48, 2048, False)]:
        packed = _randn(max(rq, rk), 3 * D, seed=1, scale=1.5)
        xq, xk = packed[:rq, :D], packed[:rk, D:2 * D]
        wq, wk = (1 + 0.1 * _randn(D, seed=2).float()).to(BF16), (1 + 0.1 * _randn(D, seed=3).float()).to(BF16)
        ang = torch.rand(rq, D // 2, device="cuda") * 6.28
        cos = ang.cos().repeat_interleave(2, -1).to(BF16) if rope else None
        sin = ang.sin().repeat_interleave(2, -1).to(BF16) if rope else None
        oq = torch.empty(rq, D, device="cuda", dtype=BF16)
        ok = torch.empty(rk, D, device="cuda", dtype=BF16)
        ops.qknorm_rope_fwd(xq, xk, wq, wk, cos, sin, oq, ok)
        outs, leaves = [], []
        for x, w in ((xq, wq), (xk, wk)):
            xf = x.float().requires_grad_(True)
            y = xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-5) * w.float()
            if rope:
                y = _rope_ref(y, cos.float(), sin.float())
            outs.append(y)
            leaves.append(xf)
        _assert_close(f"qknorm_rope fwd q rows={rq} D={D} rope={int(rope)}", oq, outs[0], 5e-3)
        _assert_close(f"qknorm_rope fwd k rows={rk} D={D} rope={int(rope)}", ok, outs[1], 5e-3)
        dq = torch.randn(rq, D, device="cuda")             # fp32, like the attention backward emits
        dk = _randn(rk, D, seed=5)
        outs[0].backward(dq)
        outs[1].backward(dk.float())
        gq = torch.empty(rq, D, device="cuda", dtype=BF16)
        gk = torch.empty(rk, D, device="cuda", dtype=BF16)
        ops.qknorm_rope_bwd(dq, dk, xq, xk, wq, wk, cos, sin, gq, gk)
        _assert_close(f"qknorm_rope bwd q rows={rq} D={D} rope={int(rope)}", gq, leaves[0].grad, 6e-3)
        _assert_close(f"qknorm_rope bwd k rows={rk} D={D} rope={int(rope)}", gk, leaves[1].grad, 6e-3)
    torch.cuda.synchronize()


def _attn_ref(q, k, v, B, H, Nq, Nk, bias, scale):
    qh = q.float().view(B, Nq, H, 64).transpose(1, 2)
    kh = k.float().view(B, Nk, H, 64).transpose(1, 2)
    vh = v.float().view(B, Nk, H, 64).transpose(1, 2)
    s = qh @ kh.transpose(-1, -2) * scale
    if bias is not None:
        s = s + bias.float()[:, None, None, :]
    p = s.softmax(-1)
    o = (p @ vh).transpose(1, 2).reshape(B * Nq, H * 64)
    return o, torch.logsumexp(s, -1)


ATTN_CASES = [(1, 2, 128, 128, False), (2, 4, 96, 96, False), (1, 2, 300, 300, False), (2, 4, 96, 24, True),
              (1, 32, 512, 256, True), (1, 4, 1024, 1024, False), (2, 3, 200, 333, True),
              # few key tiles, many query tiles: the backward splits the query walk over several CTAs
              (1, 4, 1100, 256, True), (1, 32, 1536, 256, True), (2, 2, 1030, 100, False),
              # >= 512 keys: the double-buffered forward (ragged last key step, bias path, several ring wraps)
              (2, 3, 700, 1000, False), (1, 2, 640, 1111, True), (1, 8, 2048, 1536, False), (1, 1, 130, 577, False),
              # 320 work items on 296 CTA slots: the 24 items of the last wave are split along the keys (3 / 2 parts)
              (1, 32, 1280, 1536, False), (1, 32, 1200, 1111, True),
              # backward: 320 key-tile items on 148 SMs -> the 24 items of the last wave split 2-way along the queries
              (1, 32, 2100, 1280, False)]


def _mask_bias(kind, B, Nk):
    """Key-bias patterns the masked-tile skip has to get right (reference mask: 0 / -10000, transformer3d.py:440-445)."""
    bias = torch.full((B, Nk), -10000.0, device="cuda")
    if kind == "prompt":          # a short prompt padded to Nk: everything past the first tile is dead
        bias[:, :15] = 0.0
    elif kind == "ragged":        # a different live length per batch entry
        for b in range(B):
            bias[b, :15 + 150 * b] = 0.0
    elif kind == "tail":          # only a few late keys are live: leading tiles are dead, nothing may be cut off
        bias[:, Nk - 6:Nk - 2] = 0.0
    elif kind == "holes":         # live keys in the first and the last tile, dead tiles in between
        bias[:, 3] = 0.0
        bias[:, Nk - 2] = 0.5
    elif kind == "dead":          # every key masked: the softmax is uniform, as in the reference
        pass
    return bias


MASK_CASES = [(1, 32, 1536, 256, "prompt"), (2, 4, 1100, 256, "ragged"), (2, 3, 300, 256, "tail"),
              (1, 4, 1100, 512, "holes"), (1, 2, 200, 256, "dead"), (2, 2, 130, 384, "prompt")]


def check_attention_masked_tiles():
    from b200_ltx import ops
    for (B, H, Nq, Nk, kind) in MASK_CASES:
        D = H * 64
        q, k, v = _randn(B * Nq, D, seed=1), _randn(B * Nk, D, seed=2), _randn(B * Nk, D, seed=3)
        do = _randn(B * Nq, D, seed=4)
        bias = _mask_bias(kind, B, Nk)
        o, lse = ops.fa_fwd(q, k, v, B, H, Nq, Nk, bias, 0.125)
        dk = torch.full_like(k, float("nan"))
        dv = torch.full_like(v, float("nan"))
        dq = ops.fa_bwd(q, k, v, o, do, lse, B, H, Nq, Nk, dk, dv, bias, 0.125)
        qf, kf, vf = [t.float().requires_grad_(True) for t in (q, k, v)]
        oref, lref = _attn_ref(qf, kf, vf, B, H, Nq, Nk, bias, 0.125)
        oref.backward(do.float())
        tag = f"B={B} H={H} Nq={Nq} Nk={Nk} mask={kind}"
        _assert_close("fa_fwd o " + tag, o, oref, 6e-3)
        # lse sits near -10000 when every key is masked: compare relative to its magnitude there
        lim = 2e-3 * max(1.0, float(lref.abs().max()) / 10.0)
        assert (lse - lref.detach()).abs().max().item() < lim, "fa_fwd lse " + tag
        _assert_close("fa_bwd dq " + tag, dq, qf.grad, 1.2e-2)
        _assert_close("fa_bwd dk " + tag, dk, kf.grad, 1.2e-2)
        _assert_close("fa_bwd dv " + tag, dv, vf.grad, 1.2e-2)
        if kind != "dead":
            dead = (bias <= -9000).reshape(-1)
            assert float(dk[dead].abs().max()) == 0.0 and float(dv[dead].abs().max()) == 0.0, "masked keys: " + tag
    torch.cuda.synchronize()


def check_attention_fwd():
    from b200_ltx import ops
    for (B, H, Nq, Nk, use_bias) in ATTN_CASES:
        D = H * 64
        qkv = _randn(B * max(Nq, Nk), 3 * D, seed=1)
        q, k, v = qkv[:B * Nq, :D], qkv[:B * Nk, D:2 * D], qkv[:B * Nk, 2 * D:]
        bias = None
        if use_bias:
            bias = torch.zeros(B, Nk, device="cuda")
            bias[:, (Nk * 2) // 3:] = -10000.0
            bias[:, 0] = 0.5
        o, lse = ops.fa_fwd(q, k, v, B, H, Nq, Nk, bias, 0.125)
        oref, lref = _attn_ref(q, k, v, B, H, Nq, Nk, bias, 0.125)
        tag = f"B={B} H={H} Nq={Nq} Nk={Nk} bias={int(use_bias)}"
        _assert_close("fa_fwd o " + tag, o, oref, 8e-3)
        _assert_close("fa_fwd lse " + tag, lse, lref, 1e-3)
    # large-magnitude scores exercise the lazy rescale (N = 256: the few-key kernel, 512 / 1536: the double-buffered one)
    for N in (256, 512, 1536):
        B, H = 1, 2
        q, k, v = _randn(N, 128, seed=3, scale=6.0), _randn(N, 128, seed=4, scale=6.0), _randn(N, 128, seed=5)
        o, lse = ops.fa_fwd(q, k, v, B, H, N, N, None, 0.125)
        oref, lref = _attn_ref(q, k, v, B, H, N, N, None, 0.125)
        _assert_close(f"fa_fwd o (peaked softmax, N={N})", o, oref, 1e-2)
        _assert_close(f"fa_fwd lse (peaked softmax, N={N})", lse, lref, 1e-3)
    torch.cuda.synchronize()


def check_attention_bwd():
    from b200_ltx import ops
    for (B, H, Nq, Nk, use_bias) in ATTN_CASES:
        D = H * 64
        q, k, v = _randn(B * Nq, D, seed=1), _randn(B * Nk, D, seed=2), _randn(B * Nk, D, seed=3)
        do = _randn(B * Nq, D, seed=4)
        bias = None
        if use_bias:
            bias = torch.zeros(B, Nk, device="cuda")
            bias[:, (Nk * 2) // 3:] = -10000.0
        o, lse = ops.fa_fwd(q, k, v, B, H, Nq, Nk, bias, 0.125)
        dk = torch.empty_like(k)
        dv = torch.empty_like(v)
        dq = ops.fa_bwd(q, k, v, o, do, lse, B, H, Nq, Nk, dk, dv, bias, 0.125)
        qf, kf, vf = [t.float().requires_grad_(True) for t in (q, k, v)]
        oref, _ = _attn_ref(qf, kf, vf, B, H, Nq, Nk, bias, 0.125)
        oref.backward(do.float())
        tag = f"B={B} H={H} Nq={Nq} Nk={Nk} bias={int(use_bias)}"
        _assert_close("fa_bwd dq " + tag, dq, qf.grad, 1.2e-2)
        _assert_close("fa_bwd dk " + tag, dk, kf.grad, 1.2e-2)
        _assert_close("fa_bwd dv " + tag, dv, vf.grad, 1.2e-2)
    torch.cuda.synchronize()


def check_attn_core_fn():
    from b200_ltx import ops
    B, H, Nq = 2, 4, 96
    D = H * 64
    qkv = _randn(B * Nq, 3 * D, seed=1).requires_grad_(True)
    wq = (1 + 0.1 * _randn(D, seed=2).float()).to(BF16)
    wk = (1 + 0.1 * _randn(D, seed=3).float()).to(BF16)
    ang = torch.rand(B * Nq, D // 2, device="cuda") * 6.28
    cos, sin = ang.cos().repeat_interleave(2, -1).to(BF16), ang.sin().repeat_interleave(2, -1).to(BF16)
    o = ops.AttnCoreFn.apply(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], wq, wk, cos, sin, None, B, H, Nq, Nq, 0.125)
    do = _randn(B * Nq, D, seed=4)
    o.backward(do)
    x = qkv.detach().float().requires_grad_(True)

    def nrm(t, w):
        y = t * torch.rsqrt(t.pow(2).mean(-1, keepdim=True) + 1e-5) * w.float()
        return _rope_ref(y, cos.float(), sin.float())
    oref, _ = _attn_ref(nrm(x[:, :D], wq), nrm(x[:, D:2 * D], wk), x[:, 2 * D:], B, H, Nq, Nq, None, 0.125)
    oref.backward(do.float())
    _assert_close("AttnCoreFn o", o, oref, 8e-3)
    _assert_close("AttnCoreFn dqkv", qkv.grad, x.grad, 1.5e-2)
    torch.cuda.synchronize()


def check_key_sharded_merge():
    """The ring-attention building blocks on ONE device: attend to two disjoint key shards, fold the
    partials with attn_merge, accumulate dq over the shards with fa_bwd(dq_accum=...) -- must equal
    the single-pass kernels' reference (plain torch fp32 over all keys)."""
    from b200_ltx import ops
    for (B, H, Nq, n_sh, shards) in [(1, 2, 200, 136, 2), (2, 4, 96, 96, 3), (1, 32, 384, 264, 2)]:
        D = H * 64
        Nk = n_sh * shards
        q, do = _randn(B * Nq, D, seed=1), _randn(B * Nq, D, seed=4)
        k, v = _randn(B, Nk, D, seed=2), _randn(B, Nk, D, seed=3)
        ks = [k[:, i * n_sh:(i + 1) * n_sh].reshape(B * n_sh, D).contiguous() for i in range(shards)]
        vs = [v[:, i * n_sh:(i + 1) * n_sh].reshape(B * n_sh, D).contiguous() for i in range(shards)]
        o_acc = torch.empty(B * Nq, D, device="cuda")
        lse_acc = torch.empty(B, H, Nq, device="cuda")
        out = torch.empty(B * Nq, D, device="cuda", dtype=BF16)
        for i in range(shards):
            o_i, lse_i = ops.fa_fwd(q, ks[i], vs[i], B, H, Nq, n_sh, None, 0.125)
            ops.attn_merge(o_acc, lse_acc, o_i, lse_i, B, H, Nq, i == 0, out if i == shards - 1 else None)
        qf, kf, vf = q.float().requires_grad_(True), k.reshape(B * Nk, D).float().requires_grad_(True), \
            v.reshape(B * Nk, D).float().requires_grad_(True)
        oref, lref = _attn_ref(qf, kf, vf, B, H, Nq, Nk, None, 0.125)
        oref.backward(do.float())
        tag = f"B={B} H={H} Nq={Nq} shards={shards}x{n_sh}"
        _assert_close("merged o " + tag, out, oref, 8e-3)
        _assert_close("merged o (fp32 acc) " + tag, o_acc, oref, 8e-3)
        _assert_close("merged lse " + tag, lse_acc, lref, 1e-3)
        # the pre-pass clears the dQ accumulator on its way (a pitched buffer: the columns beside it stay untouched)
        wide = torch.full((B * Nq, D + 8), 3.0, device="cuda")
        dq = wide[:, :D]
        delta = ops.attn_delta(out, do, B, H, Nq, zero=dq)
        assert float(dq.abs().max()) == 0.0 and bool((wide[:, D:] == 3.0).all()), "attn_delta zeroing " + tag
        assert torch.equal(delta, ops.attn_delta(out, do, B, H, Nq)), "attn_delta with / without zeroing " + tag
        dks, dvs = [], []
        for i in range(shards):
            dk, dv = torch.empty_like(ks[i]), torch.empty_like(vs[i])
            ops.fa_bwd(q, ks[i], vs[i], out, do, lse_acc, B, H, Nq, n_sh, dk, dv, None, 0.125, delta=delta, dq_accum=dq)
            dks.append(dk.view(B, n_sh, D))
            dvs.append(dv.view(B, n_sh, D))
        _assert_close("sharded dq " + tag, dq, qf.grad, 1.2e-2)
        _assert_close("sharded dk " + tag, torch.cat(dks, 1).reshape(B * Nk, D), kf.grad, 1.2e-2)
        _assert_close("sharded dv " + tag, torch.cat(dvs, 1).reshape(B * Nk, D), vf.grad, 1.2e-2)
    torch.cuda.synchronize()


def check_empty_and_degenerate():
    """Empty / degenerate inputs go through the C ABI without a launch fault: zero rows, zero queries, a single
    token, one key."""
    from b200_ltx import ops
    z = torch.empty(0, 256, device="cuda", dtype=BF16)
    w = _randn(512, 256, seed=1)
    assert ops.gemm(z, w).shape == (0, 512)
    assert ops.norm_mod_fwd(z, None, None, 1, 1e-6).shape == (0, 256)
    q0 = torch.empty(0, 128, device="cuda", dtype=BF16)
    k1, v1 = _randn(5, 128, seed=2), _randn(5, 128, seed=3)
    o, _ = ops.fa_fwd(q0, k1, v1, 1, 2, 0, 5, None, 0.125)
    assert o.shape == (0, 128)
    # one query, one key: softmax of a single score is 1 -> o == v
    q, k, v = _randn(1, 128, seed=4), _randn(1, 128, seed=5), _randn(1, 128, seed=6)
    o, lse = ops.fa_fwd(q, k, v, 1, 2, 1, 1, None, 0.125)
    _assert_close("fa_fwd 1 query x 1 key == v", o, v, 1e-6)
    do = _randn(1, 128, seed=7)
    dk, dv = torch.empty_like(k), torch.empty_like(v)
    dq = ops.fa_bwd(q, k, v, o, do, lse, 1, 2, 1, 1, dk, dv, None, 0.125)
    assert float(dq.abs().max()) < 1e-5 and float(dk.float().abs().max()) < 1e-5  # d softmax of one score is 0
    _assert_close("fa_bwd 1x1 dv == do", dv, do, 1e-6)
    # all keys but one masked with the reference's -10000 bias: attention collapses onto that key
    q, k, v = _randn(130, 128, seed=8), _randn(70, 128, seed=9), _randn(70, 128, seed=10)
    bias = torch.full((1, 70), -10000.0, device="cuda")
    bias[0, 33] = 0.0
    o, _ = ops.fa_fwd(q, k, v, 1, 2, 130, 70, bias, 0.125)
    _assert_close("fa_fwd single unmasked key", o, v[33:34].expand(130, -1), 1e-6)
    torch.cuda.synchronize()


def check_rf_and_misc():
    from b200_ltx import ops
    B, N, C = 3, 96, 128
    x0, eps = _randn(B, N, C, seed=1), _randn(B, N, C, seed=2)
    t = torch.tensor([0.1, 0.5, 0.93], device="cuda")
    xt, v = ops.rf_noise(x0, eps, t)
    tt = t[:, None, None]
    assert torch.equal(xt, ((1 - tt) * x0 + tt * eps).to(BF16)), "rf_noise x_t must be bit-exact"
    assert torch.equal(v, (-1.0 * x0.float() + 1.0 * eps.float()).to(BF16)), "rf_noise v must be bit-exact"
    out = _randn(B, N, C, seed=3)
    loss, dout = ops.rf_loss(out, v, 1.0)
    of = out.float().requires_grad_(True)
    lr = F.mse_loss(of, v.float())
    lr.backward()
    _assert_close("rf_loss value", loss, lr, 1e-5)
    _assert_close("rf_loss grad", dout, of.grad, 4e-3)
    # conditioning lerp: bit-exact against torch.lerp in bf16
    Fr, H, W = 3, 4, 8
    tok = _randn(B, Fr * H * W, C, seed=4)
    ref = _randn(B, C, 1, H, W, seed=5)
    pose = _randn(B, C, Fr, H, W, seed=6)
    want = tok.clone()
    vw = want.view(B, Fr, H, W, C).permute(0, 4, 1, 2, 3)
    vw[:, :, 0:1] = torch.lerp(vw[:, :, 0:1], ref, 0.85)
    vw[:, :, 1:] = torch.lerp(vw[:, :, 1:], pose[:, :, 1:], 0.5)
    got = ops.lerp_condition_(tok.clone(), ref, pose)
    d = _maxabs(got, want)
    print(f"  lerp_condition maxabs diff vs torch.lerp bf16: {d:.3e}", flush=True)
    assert d <= 2 ** -6, d  # at most one bf16 ulp of O(1) values
    assert _rel(got, want) < 2e-3
    x = _randn(200, 256, seed=7)
    g = _randn(4, 6 * 256, seed=8)[:, 512:768]
    _assert_close("rowscale", ops.rowscale(x, g, 50), x.float() * g.float().repeat_interleave(50, 0), 4e-3)
    _assert_close("colsum", ops.colsum(x), x.float().sum(0), 1e-5)
    torch.cuda.synchronize()


def check_gemm_stream_k():
    """Shapes whose last wave of tiles is partial: the k blocks of those tiles are spread over all CTAs and the
    partials meet in the owner's epilogue (b200_gemm_bf16_ws).  Against fp32 matmul, against the same GEMM with
    stream-K switched off, and twice in a row (the flags must come back clean)."""
    import os
    from b200_ltx import ops
    cases = [  # M, N, K, K2, b_rows_are_k, epilogue operands, block_n
        (6144, 2048, 1024, 0, False, "", 0),
        (6144, 2048, 2048, 64, False, "bias gate res", 0),
        (6144, 2048, 2048, 0, True, "res", 0),
        (6000, 2048, 512, 0, False, "bias", 0),
        (6144, 8192, 512, 0, False, "gelu", 0),
        (4096, 2048, 512, 0, False, "bias res", 128),
        # few-wave problems with ragged M (the shards of 8 / 4 sequence-parallel ranks at cfg5)
        (1584, 2048, 2048, 0, False, "bias gate res", 0),
        (1584, 2048, 2048, 64, False, "bias", 0),
        (1584, 6144, 2048, 0, False, "bias", 0),
        (1584, 8192, 2048, 0, False, "gelu", 0),
        (1584, 2048, 8192, 0, True, "res", 0),
        (1584, 2048, 2048, 0, True, "", 256),
        (3168, 2048, 2048, 0, False, "bias gate res", 0),
        (3168, 8192, 2048, 0, False, "gelu", 0),
        (840, 2048, 2048, 0, False, "bias", 128),
        (256, 2048, 2048, 0, False, "bias", 0),
    ]
    for (M, N, K, K2, b_k, epi, bn) in cases:
        a = _randn(M, K, seed=1, scale=0.5)
        b = _randn(K, N, seed=2, scale=0.5) if b_k else _randn(N, K, seed=2, scale=0.5)
        kw = dict(b_rows_are_k=b_k, block_n=bn)
        ref = a.float() @ (b.float() if b_k else b.float().T)
        if K2:
            a2, b2 = _randn(M, K2, seed=3, scale=0.5), _randn(N, K2, seed=4, scale=0.5)
            kw.update(a2=a2, b2=b2)
            ref = ref + a2.float() @ b2.float().T
        if "bias" in epi:
            kw["bias"] = _randn(N, seed=5)
            ref = ref + kw["bias"].float()
        pre = None
        if "gelu" in epi:
            pre = torch.empty(M, N, device="cuda", dtype=BF16)
            kw.update(epilogue=ops.EPI_GELU, aux=pre, bias=_randn(N, seed=5))
            ref = F.gelu((ref + kw["bias"].float()).to(BF16).float(), approximate="tanh")
        if "gate" in epi:
            kw.update(gate=_randn(1, N, seed=6), rows_per_gate=M)
            ref = ref * kw["gate"].float()
        if "res" in epi:
            kw["res"] = _randn(M, N, seed=7)
            ref = ref + kw["res"].float()
        tag = f"M={M} N={N} K={K}+{K2} b_k={int(b_k)} epi='{epi}' bn={bn}"
        os.environ.pop("B200_GEMM_STREAMK", None)
        plain = ops.gemm(a, b, **kw).clone()
        os.environ["B200_GEMM_STREAMK"] = "1"   # off by default (measured neutral on this power-capped part)
        try:
            out1 = ops.gemm(a, b, **kw).clone()
            out2 = ops.gemm(a, b, **kw)
        finally:
            os.environ.pop("B200_GEMM_STREAMK", None)
        _assert_close("stream-K gemm " + tag, out1, ref, 6e-3)
        _assert_close("stream-K vs whole tiles " + tag, out1, plain, 2e-3)
        assert torch.equal(out1, out2), "stream-K not repeatable: " + tag
    torch.cuda.synchronize()
    ws = ops._gemm_workspace(torch.device("cuda", torch.cuda.current_device()))
    assert int(ws[:16384].view(torch.int32).abs().sum()) == 0, "stream-K flags left set"


def check_adamw():
    """b200_ltx.optim.FusedAdamW (one launch over all tensors) against torch.optim.AdamW, 6 steps, fp32 and bf16
    tensors, aligned and ragged sizes, a parameter without gradient, and a state_dict round trip."""
    from b200_ltx import optim
    shapes = [((2048, 32), torch.float32), ((32, 2048), torch.float32), ((1024, 2048), BF16), ((2048,), BF16),
              ((1000, 3), torch.float32), ((777,), BF16), ((5,), torch.float32)]
    g = torch.Generator(device="cpu").manual_seed(0)
    init = [torch.randn(*s, generator=g).to("cuda", dt) for s, dt in shapes]
    mine = [torch.nn.Parameter(t.clone()) for t in init] + [torch.nn.Parameter(torch.ones(7, device="cuda"))]
    ref = [torch.nn.Parameter(t.clone()) for t in init] + [torch.nn.Parameter(torch.ones(7, device="cuda"))]
    kw = dict(lr=3e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2)
    o_mine, o_ref = optim.FusedAdamW(mine, **kw), torch.optim.AdamW(ref, fused=True, **kw)
    for it in range(6):
        if it == 3:   # LR schedule step + checkpoint round trip
            for o in (o_mine, o_ref):
                o.param_groups[0]["lr"] = 1e-3
            sd = o_mine.state_dict()
            o_mine = optim.FusedAdamW(mine, **kw)
            o_mine.load_state_dict(sd)
        for a, b in zip(mine[:-1], ref[:-1]):
            gr = torch.randn(a.shape, generator=g).to("cuda", a.dtype)
            a.grad, b.grad = gr.clone(), gr.clone()
        o_mine.step()
        o_ref.step()
    for i, (a, b) in enumerate(zip(mine, ref)):
        tol = 2e-6 if a.dtype == torch.float32 else 4e-3
        _assert_close(f"adamw param {i} {tuple(a.shape)} {a.dtype}", a.detach(), b.detach(), tol)
    assert torch.equal(mine[-1].detach(), ref[-1].detach())      # no gradient: untouched
    st = o_mine.state[mine[0]]
    assert float(st["step"]) == 6.0
    _assert_close("adamw exp_avg_sq", st["exp_avg_sq"], o_ref.state[ref[0]]["exp_avg_sq"], 2e-6)
    torch.cuda.synchronize()


def check_guidance_step():
    """Fused sampling tail (guidance combine + Euler + conditioning select + next model input) against the oracle's
    restatement of pipeline_ltx_video.py:1217-1260, :1346-1379 in fp32."""
    import ref_block as rb
    import ref_sampling as rs
    from b200_ltx import ops
    B, N, C = 2, 333, 128
    grid_t = rb.uniform_timesteps(8)
    for (do_cfg, do_stg, star, rsc, per_tok, n_next) in [(False, False, False, 1.0, False, 1), (True, False, False, 1.0, False, 2),
                                                         (True, False, True, 1.0, True, 0), (False, True, False, 0.7, False, 2),
                                                         (True, True, True, 0.7, True, 3), (True, True, False, 1.0, False, 1)]:
        conds = 1 + int(do_cfg) + int(do_stg)
        g = torch.Generator(device="cpu").manual_seed(conds * 10 + int(star))
        v = (torch.randn(conds * B, N, C, generator=g) + 0.1).to(BF16)
        x = torch.randn(B, N, C, generator=g)
        cm = None
        if per_tok:
            cm = torch.zeros(B, N)
            cm[:, :40] = 1.0
            cm[0, 40:90] = 0.6
            cm[1, 40:70] = 0.3
        t = grid_t[3]
        gs, stg = 3.0 if do_cfg else 1.0, 1.5 if do_stg else 0.0
        # oracle, fp32
        pred = rs.guidance_combine(v.float(), B, conds, do_cfg, do_stg, gs, stg, rsc, star)
        cur_t = t.view(1, 1).expand(B, 1)
        if cm is not None:
            cur_t = torch.min(cur_t, 1.0 - cm)
        den = rb.rf_step(grid_t, pred, cur_t[:1], x)
        ref = den if cm is None else torch.where((t - 1e-6 < (1.0 - cm)).unsqueeze(-1), den, x)
        # product
        from b200_ltx.sampling import _step_tables
        rows, dts = _step_tables(grid_t.cuda(), None if cm is None else cm.cuda())
        scal = torch.tensor([gs, stg, rsc, float(t)], device="cuda")
        xg = x.cuda().clone()
        xn = torch.full((n_next * B, N, C), float("nan"), device="cuda", dtype=BF16) if n_next else None
        ops.guidance_step_(v.cuda(), xg, xn, dts[3], None if cm is None else (1.0 - cm).cuda().contiguous(), scal,
                           do_cfg, do_stg, cfg_star=star, rescale=bool(do_stg and rsc != 1.0))
        tag = f"cfg={int(do_cfg)} stg={int(do_stg)} star={int(star)} rescale={rsc} per_token={int(per_tok)}"
        _assert_close("guidance_step x " + tag, xg.cpu(), ref, 2e-5)
        if cm is not None:
            assert torch.equal(xg[:, :40].cpu(), x[:, :40]), "hard-conditioned tokens moved: " + tag
        if n_next:
            want = xg.to(BF16).repeat(n_next, 1, 1)
            assert torch.equal(xn, want), "next model input: " + tag
    torch.cuda.synchronize()


def check_full_mode_pieces():
    """Kernels behind train_mode='full' (training.py:75-91): grouped column sums (plain and of a product), the
    dy * xhat products of the norm backward kernels, the GEMM epilogue that stashes the pre-gate output, and the
    autograd Functions' gradients for AdaLN scale / shift / gate and the qk-norm weights."""
    from b200_ltx import ops
    # grouped column sums
    for (rows, N, rpg) in [(6144, 2048, 6144), (4 * 3328, 2048, 3328), (515 * 2, 256, 515), (96, 6144, 32), (40, 64, 8)]:
        a, b = _randn(rows, N, seed=1), _randn(rows, N, seed=2)
        got = ops.colsum_groups(a, None, rpg)
        _assert_close(f"colsum_groups rows={rows} N={N} rpg={rpg}", got, a.float().view(-1, rpg, N).sum(1), 2e-3)
        got = ops.colsum_groups(a, b, rpg)
        _assert_close(f"colsum_groups product rows={rows} N={N} rpg={rpg}", got,
                      (a.float() * b.float()).view(-1, rpg, N).sum(1), 2e-3)
    # GEMM: stash of the pre-gate output
    M, N, K = 384, 512, 256
    a, b = _randn(M, K, seed=1), _randn(N, K, seed=2, scale=0.05)
    bias, gate, res = _randn(N, seed=5), _randn(3, N, seed=6), _randn(M, N, seed=7)
    for bn in (0, 64, 128, 256):
        u = torch.empty(M, N, device="cuda", dtype=BF16)
        y = ops.gemm(a, b, bias=bias, gate=gate, rows_per_gate=128, res=res, aux=u, epilogue=ops.EPI_STASH, block_n=bn)
        uref = a.float() @ b.float().t() + bias.float()
        _assert_close(f"gemm stash u bn={bn}", u, uref, 6e-3)
        _assert_close(f"gemm stash y bn={bn}", y, gate.float().repeat_interleave(128, 0) * u.float() + res.float(), 6e-3)
    # NormModResFn with trainable scale / shift (per-sample and per-token modulation), LinearFn / FeedForwardFn gates
    for (rows, D, rpm, ln) in [(640, 2048, 320, False), (515, 256, 515, True), (130, 2048, 1, False)]:
        nb = rows // rpm
        x = _randn(rows, D, seed=1, scale=2.0).requires_grad_(True)
        ada = _randn(nb, 6 * D, seed=2, scale=0.3).requires_grad_(True)
        y, xres = ops.NormModResFn.apply(x, ada[:, D:2 * D], ada[:, :D], rpm, 1e-6, ln)
        dy, dres = _randn(rows, D, seed=3), _randn(rows, D, seed=4)
        torch.autograd.backward([y, xres], [dy, dres])
        xf, af = x.detach().float().requires_grad_(True), ada.detach().float().requires_grad_(True)
        n = F.layer_norm(xf, (D,), None, None, 1e-6) if ln else xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-6)
        yr = n * (1 + af[:, D:2 * D].repeat_interleave(rpm, 0)) + af[:, :D].repeat_interleave(rpm, 0)
        torch.autograd.backward([yr, xf * 1.0], [dy.float(), dres.float()])
        tag = f"rows={rows} D={D} rpm={rpm} ln={int(ln)}"
        _assert_close("NormModResFn dx " + tag, x.grad, xf.grad, 6e-3)
        _assert_close("NormModResFn d(scale) " + tag, ada.grad[:, D:2 * D], af.grad[:, D:2 * D], 8e-3)
        _assert_close("NormModResFn d(shift) " + tag, ada.grad[:, :D], af.grad[:, :D], 8e-3)
        assert float(ada.grad[:, 2 * D:].abs().max()) == 0.0
    M, K, N = 640, 256, 512
    x = _randn(M, K, seed=1).requires_grad_(True)
    W = _randn(N, K, seed=2, scale=0.06).requires_grad_(True)
    bb = _randn(N, seed=3, scale=0.1).requires_grad_(True)
    gate = _randn(2, 3 * N, seed=6).requires_grad_(True)
    res = _randn(M, N, seed=7).requires_grad_(True)
    dy = _randn(M, N, seed=8)
    y = ops.LinearFn.apply(x, W, bb, None, None, 1.0, gate[:, N:2 * N], 320, res)
    y.backward(dy)
    xr, Wr, br, gr, rr = [t.detach().float().requires_grad_(True) for t in (x, W, bb, gate, res)]
    yr = gr[:, N:2 * N].repeat_interleave(320, 0) * (xr @ Wr.t() + br) + rr
    yr.backward(dy.float())
    _assert_close("LinearFn (trainable gate) y", y, yr, 8e-3)
    for nm, g, g_ in (("dx", x.grad, xr.grad), ("dW", W.grad, Wr.grad), ("db", bb.grad, br.grad),
                      ("d(gate)", gate.grad, gr.grad), ("dres", res.grad, rr.grad)):
        _assert_close("LinearFn (trainable gate) " + nm, g, g_, 1.5e-2)
    # AttnCoreFn with trainable qk-norm weights: attn1 form (RoPE, paired rows) and attn2 form (no RoPE, Nq != Nk)
    for (B, H, Nq, Nk, rope) in [(2, 4, 96, 96, True), (1, 4, 200, 72, False)]:
        D = H * 64
        q_pre, k_pre, v = [_randn(B * n, D, seed=i).requires_grad_(True) for i, n in ((1, Nq), (2, Nk), (3, Nk))]
        wq = (1 + 0.1 * _randn(D, seed=4).float()).to(BF16).requires_grad_(True)
        wk = (1 + 0.1 * _randn(D, seed=5).float()).to(BF16).requires_grad_(True)
        cos = sin = None
        if rope:
            ang = torch.rand(B * Nq, D // 2, device="cuda") * 6.28
            cos, sin = ang.cos().repeat_interleave(2, -1).to(BF16), ang.sin().repeat_interleave(2, -1).to(BF16)
        o = ops.AttnCoreFn.apply(q_pre, k_pre, v, wq, wk, cos, sin, None, B, H, Nq, Nk, 0.125)
        do = _randn(B * Nq, D, seed=6)
        o.backward(do)
        ref = [t.detach().float().requires_grad_(True) for t in (q_pre, k_pre, v, wq, wk)]

        def nrm(t, w):
            y = t * torch.rsqrt(t.pow(2).mean(-1, keepdim=True) + 1e-5) * w
            return _rope_ref(y, cos.float(), sin.float()) if rope else y
        oref, _ = _attn_ref(nrm(ref[0], ref[3]), nrm(ref[1], ref[4]), ref[2], B, H, Nq, Nk, None, 0.125)
        oref.backward(do.float())
        tag = f"B={B} Nq={Nq} Nk={Nk} rope={int(rope)}"
        _assert_close("AttnCoreFn (trainable norms) o " + tag, o, oref, 8e-3)
        for nm, t, tr in zip(("dq_pre", "dk_pre", "dv", "d(q_norm.w)", "d(k_norm.w)"), (q_pre, k_pre, v, wq, wk), ref):
            _assert_close(f"AttnCoreFn (trainable norms) {nm} " + tag, t.grad, tr.grad, 2e-2)
    torch.cuda.synchronize()


GROUPS = {
    "full_mode_pieces": check_full_mode_pieces,
    "gemm_layouts": check_gemm_layouts,
    "gemm_epilogues": check_gemm_epilogues,
    "gemm_batched": check_gemm_batched,
    "gemm_stream_k": check_gemm_stream_k,
    "linear_fn": check_linear_fn,
    "norm_mod": check_norm_mod,
    "qknorm_rope": check_qknorm_rope,
    "attention_fwd": check_attention_fwd,
    "attention_bwd": check_attention_bwd,
    "attention_masked_tiles": check_attention_masked_tiles,
    "attn_core_fn": check_attn_core_fn,
    "key_sharded_merge": check_key_sharded_merge,
    "empty_and_degenerate": check_empty_and_degenerate,
    "rf_and_misc": check_rf_and_misc,
    "guidance_step": check_guidance_step,
    "adamw": check_adamw,
}
