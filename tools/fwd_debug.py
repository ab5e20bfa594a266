"""Forward-kernel debugging aid: peaked-softmax cases (frequent lazy rescales) at several key counts, error per query tile."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from b200_ltx import ops

def ref(q, k, v, H, N):
    qh = q.float().view(N, H, 64).transpose(0, 1); kh = k.float().view(N, H, 64).transpose(0, 1); vh = v.float().view(N, H, 64).transpose(0, 1)
    s = qh @ kh.transpose(-1, -2) * 0.125
    return (s.softmax(-1) @ vh).transpose(0, 1).reshape(N, H * 64), torch.logsumexp(s, -1)

for scale in (1.0, 6.0):
    for N in (512, 640, 1024, 1536, 2048):
        H = 2
        g = torch.Generator().manual_seed(3)
        q, k, v = [(torch.randn(N, 128, generator=g) * s).cuda().bfloat16() for s in (scale, scale, 1.0)]
        o, lse = ops.fa_fwd(q, k, v, 1, H, N, N, None, 0.125)
        oref, lref = ref(q, k, v, H, N)
        err = (o.float() - oref).norm() / oref.norm()
        per_tile = [(float((o[i:i + 128].float() - oref[i:i + 128]).norm() / oref[i:i + 128].norm())) for i in range(0, N, 128)]
        bad_rows = ((o.float() - oref).abs().amax(1) > 0.05).nonzero().flatten()
        print(f"scale {scale} N {N}: rel {float(err):.3e} lse maxabs {float((lse - lref).abs().max()):.3e}  bad rows {bad_rows.numel()} "
              f"first {bad_rows[:8].tolist()}  tiles {['%.1e' % e for e in per_tile[:6]]}", flush=True)

# which key step's contribution is missing (or duplicated) in the bad rows?
if os.environ.get("B200_DBG_STEPS"):
    N, H, scale = 1536, 2, 6.0
    g = torch.Generator().manual_seed(3)
    q, k, v = [(torch.randn(N, 128, generator=g) * s).cuda().bfloat16() for s in (scale, scale, 1.0)]
    o, lse = ops.fa_fwd(q, k, v, 1, H, N, N, None, 0.125)
    oref, lref = ref(q, k, v, H, N)
    bad = ((o.float() - oref).abs().amax(1) > 0.05).nonzero().flatten().tolist()
    for r in bad[:12]:
        for h in range(H):
            d = (oref[r, h * 64:(h + 1) * 64] - o[r, h * 64:(h + 1) * 64].float())
            if d.abs().max() < 0.05:
                continue
            s = (q[r, h * 64:(h + 1) * 64].float() @ k[:, h * 64:(h + 1) * 64].float().T) * 0.125
            p = torch.softmax(s, -1)
            contrib = torch.stack([p[j * 64:(j + 1) * 64] @ v[j * 64:(j + 1) * 64, h * 64:(h + 1) * 64].float() for j in range(N // 64)])
            res = [(float((d - c).norm()), j) for j, c in enumerate(contrib)]
            res2 = [(float((d + c).norm()), j) for j, c in enumerate(contrib)]
            mass = [round(float(p[j * 64:(j + 1) * 64].sum()), 3) for j in range(N // 64)]
            print(f"row {r} (lane {r % 32}, warp rows {r - r % 32}) head {h}: |diff| {float(d.norm()):.3f}  best 'lost step' {min(res)}  best 'doubled step' {min(res2)}  step masses {mass}")
