"""CPU-side checks of the drop-in boundary: libb200ltx.so loads and exports every symbol that
include/b200ltx.h declares, argument contracts are enforced before any launch, and the product
path fails loudly without a GPU (no CPU fallback)."""
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "b200ltx.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from b200_ltx import lib
    if not os.path.exists(lib.LIB_PATH):
        from b200_ltx import build
        build.build()
    h = lib.load()
    names = _header_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(h, n), n
    assert sorted(lib.PROTOTYPES) == names, "lib.PROTOTYPES and include/b200ltx.h must list the same entry points"
    assert h.b200_version() >= 100


def _strip_comments(src):
    return re.sub(r"//[^\n]*", "", re.sub(r"/\*.*?\*/", "", src, flags=re.S))


def _signatures(src, definitions):
    """name -> (return type, [argument kinds]) of every b200_* function declared (or, with `definitions`, defined as
    extern "C") in `src`.  Kinds: P pointer, L int64_t, I int, F float -- what an FFI has to get right."""
    def kind(t):
        t = t.strip()
        if t in ("void", ""):
            return None
        if "*" in t:
            return "P"
        if "int64_t" in t or "size_t" in t:
            return "L"
        if re.search(r"\bfloat\b", t):
            return "F"
        if re.search(r"\bint(32_t)?\b", t):
            return "I"
        return "?" + t
    out = {}
    pat = r'(extern\s+"C"\s+)?((?:const\s+)?[A-Za-z_0-9]+\s*\*?)\s*\b(b200_[a-z0-9_]+)\s*\(([^)]*)\)\s*([;{])'
    for m in re.finditer(pat, _strip_comments(src)):
        ext, ret, name, args, end = m.groups()
        if definitions and not (ext and end == "{"):
            continue
        out[name] = (ret.replace(" ", ""), [k for k in (kind(a) for a in args.split(",")) if k])
    return out


def test_header_definitions_and_ctypes_prototypes_agree():
    """Three statements of the ABI must agree argument by argument: the declarations in include/b200ltx.h (what a
    maintainer binds), the extern "C" definitions in csrc/*.cu (which do not include the public header, so the compiler
    never compares them) and the ctypes prototypes in lib.py (what the product calls through)."""
    from b200_ltx import lib
    header = _signatures(open(os.path.join(ROOT, "include", "b200ltx.h")).read(), definitions=False)
    csrc = os.path.join(ROOT, "video-generation-for-human-avatars_b200", "csrc")
    defs = {}
    for f in sorted(os.listdir(csrc)):
        if f.endswith(".cu"):
            defs.update(_signatures(open(os.path.join(csrc, f)).read(), definitions=True))
    debug_only = {n for n in defs if n.startswith("b200_debug_")}        # -DB200_TRACE builds, not part of the ABI
    assert set(header) == set(defs) - debug_only
    ctype_kind = {lib.P: "P", lib.L: "L", lib.I: "I", lib.F: "F"}
    ret_kind = {"int": lib.I, "int64_t": lib.L, "constchar*": lib.c_char_p}
    for name, (ret, args) in header.items():
        assert not any(a.startswith("?") for a in args), (name, args)
        assert defs[name] == (ret, args), (name, defs[name], (ret, args))
        res, argtypes = lib.PROTOTYPES[name]
        assert res is ret_kind[ret], (name, ret)
        assert [ctype_kind[t] for t in argtypes] == args, (name, [ctype_kind[t] for t in argtypes], args)


def test_integration_doc_matches_the_abi():
    """INTEGRATION.md is what a maintainer copies from: every entry point it names exists in the header (the `fwd/bwd`
    shorthand of its table expanded), and the ctypes stub it prints carries the argument list lib.py binds."""
    from b200_ltx import lib
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    names = set(_header_symbols())
    mentioned = set()
    for m in re.finditer(r"`(b200_[a-z0-9_]+)((?:/[a-z0-9_]+)*)`", doc):
        base, alts = m.group(1), [a for a in m.group(2).split("/") if a]
        mentioned.add(base)
        for a in alts:                       # `b200_norm_mod_fwd/bwd` -> b200_norm_mod_bwd
            mentioned.add(base[:base.rfind("_") + 1] + a)
    mentioned = {n for n in mentioned if not n.startswith("b200_ltx")}     # the Python package name
    assert mentioned and mentioned <= names, sorted(mentioned - names)
    assert len(mentioned) >= 25
    short = {"P": lib.P, "I": lib.I, "L": lib.L, "F": lib.F}
    stubs = re.findall(r"lib\.(b200_[a-z0-9_]+)\.argtypes = \[([^\]]*)\]", doc)
    assert stubs
    for name, args in stubs:
        assert [short[a.strip()] for a in args.split(",")] == lib.PROTOTYPES[name][1], name


def test_argument_contract_rejected_before_launch():
    from b200_ltx import lib
    h = lib.load()
    # null operands / misaligned pitch -> negative code, message set, nothing launched (works without a GPU)
    rc = h.b200_gemm_bf16(None, 8, 0, None, 8, 0, None, 0, None, 0, 0, None, 8, 0, 128, 128, 64, 0,
                          None, None, 0, 0, None, 0, None, 0, 0, 1, None)
    assert rc < 0 and b"null" in h.b200_last_error()
    rc = h.b200_fa_fwd(16, 64, 16, 64, 16, 64, 16, 64, None, None, 1, 1, 128, 128, 128, 0.125, None)
    assert rc < 0 and b"head_dim" in h.b200_last_error()
    rc = h.b200_norm_mod_fwd(16, 2048, 16, 2048, None, None, 0, 4, 4096, 4, 1e-6, 0, None)
    assert rc < 0
    with pytest.raises(lib.B200Error):
        lib.check(rc, "norm_mod_fwd")


A, MIS = 0x10000, 0x10004   # stand-in device addresses: 16-byte aligned / misaligned (never dereferenced: every
#                               case below is refused by the host-side contract check, before any launch)


def _gemm_args(**kw):
    """Positional argument list of b200_gemm_bf16 for a well-formed 128 x 128 x 64 problem, with overrides."""
    d = dict(A=A, lda=64, a_mn=0, B=A, ldb=64, b_mn=0, A2=None, lda2=0, B2=None, ldb2=0, K2=0, C=A, ldc=128, f32=0,
             M=128, N=128, K=64, epi=0, bias=None, gate=None, gate_stride=0, rows_per_gate=0, res=None, ldres=0,
             aux=None, ldaux=0, block_n=0, split_k=1, stream=None)
    assert set(kw) <= set(d), set(kw) - set(d)
    d.update(kw)
    return list(d.values())


CONTRACT_CASES = [
    # (entry point, arguments, fragment of the message)
    ("b200_gemm_bf16", _gemm_args(N=100), "multiples of 8"),
    ("b200_gemm_bf16", _gemm_args(K=60), "multiples of 8"),
    ("b200_gemm_bf16", _gemm_args(a_mn=1, M=100), "multiples of 8"),
    ("b200_gemm_bf16", _gemm_args(M=-1), "negative dimension"),
    ("b200_gemm_bf16", _gemm_args(K=0), "empty reduction"),
    ("b200_gemm_bf16", _gemm_args(A=MIS), "16-byte aligned"),
    ("b200_gemm_bf16", _gemm_args(lda=60), "16-byte aligned"),
    ("b200_gemm_bf16", _gemm_args(ldc=130, f32=1), "16-byte aligned"),
    ("b200_gemm_bf16", _gemm_args(K2=64), "second operand pair"),
    ("b200_gemm_bf16", _gemm_args(epi=7), "unknown epilogue"),
    ("b200_gemm_bf16", _gemm_args(epi=2), "GELU_GRAD needs aux"),
    ("b200_gemm_bf16", _gemm_args(epi=3, aux=A, ldaux=128, f32=1), "STASH needs aux and a bf16 output"),
    ("b200_gemm_bf16", _gemm_args(res=MIS, ldres=128), "epilogue operands"),
    ("b200_gemm_bf16", _gemm_args(gate=A, gate_stride=128, rows_per_gate=0), "epilogue operands"),
    ("b200_gemm_bf16", _gemm_args(block_n=96), "block_n"),
    ("b200_gemm_bf16", _gemm_args(split_k=4), "split_k > 1 needs a plain fp32 output"),
    ("b200_gemm_bf16", _gemm_args(split_k=4, f32=1, bias=A), "split_k > 1 needs a plain fp32 output"),
    ("b200_gemm_bf16_batched", [A, 64, 0, A, 64, 0, None, 0, None, 0, 0, A, 128, 0, 128, 128, 64, None, 0, 0, None, None],
     "bad group description"),
    ("b200_fa_fwd", [A, 64, A, 64, A, 64, A, 64, None, None, 1, 1, 128, 128, 128, 0.125, None], "head_dim"),
    ("b200_fa_fwd", [A, 64, A, 64, A, 64, A, 64, None, None, 1, 0, 128, 128, 64, 0.125, None], "bad shape"),
    ("b200_fa_fwd", [A, 64, None, 64, A, 64, A, 64, None, None, 1, 1, 128, 128, 64, 0.125, None], "null pointer"),
    ("b200_fa_fwd", [A, 64, A, 64, A, 64, A, 64, None, None, 1, 1, 128, 0, 64, 0.125, None], "no keys"),
    ("b200_fa_fwd", [A, 60, A, 64, A, 64, A, 64, None, None, 1, 1, 128, 128, 64, 0.125, None], "16-byte aligned"),
    ("b200_fa_fwd", [A, 64, A, 64, A, 64, A, 64, None, None, 1, 2, 128, 128, 64, 0.125, None], "pitch < H*64"),
    ("b200_fa_bwd", [A, 64, A, 64, A, 64, A, 64, A, A, None, A, 64, A, 64, A, 64, 1, 1, 128, 128, 32, 0.125, None, 0, None],
     "head_dim"),
    ("b200_fa_bwd", [A, 64, A, 64, A, 64, A, 64, None, A, None, A, 64, A, 64, A, 64, 1, 1, 128, 128, 64, 0.125, None, 0,
                     None], "null pointer"),
    ("b200_fa_bwd", [A, 64, A, 64, A, 64, A, 64, A, A, None, A, 62, A, 64, A, 64, 1, 1, 128, 128, 64, 0.125, None, 0, None],
     "16-byte aligned"),
    ("b200_norm_mod_fwd", [A, 2048, A, 2048, None, None, 0, 4, 4096, 4, 1e-6, 0, None], "D must be a multiple of 8 and <= 2048"),
    ("b200_norm_mod_fwd", [None, 2048, A, 2048, None, None, 0, 4, 2048, 4, 1e-6, 0, None], "null pointer or bad shape"),
    ("b200_norm_mod_fwd", [A, 2048, A, 2048, None, None, 0, 4, 2048, 0, 1e-6, 0, None], "rows_per_mod must be positive"),
    ("b200_norm_mod_bwd", [A, 2048, A, 2048, None, 0, None, 0, A, 2048, None, 0, 4, 2044, 4, 1e-6, 0, None],
     "D must be a multiple of 8"),
    ("b200_qknorm_rope_fwd", [A, 2048, A, 2048, A, A, A, None, 2048, A, 2048, A, 2048, 4, 4, 2048, 1e-5, None],
     "cos/sin must come together"),
    ("b200_qknorm_rope_fwd", [A, 2048, A, 2048, None, A, None, None, 0, A, 2048, A, 2048, 4, 4, 2048, 1e-5, None],
     "qknorm_rope_fwd"),
    ("b200_qknorm_rope_bwd", [A, 2048, 1, A, 2048, 0, A, 2048, A, 2048, A, A, None, None, 0, A, 2048, A, 2048, None, 0, None,
                              0, 4, 4, 4096, 1e-5, None], "D must be a multiple of 8 and <= 2048"),
    ("b200_rf_noise", [A, A, None, A, A, 2, 1024, None], "null pointer or bad shape"),
    ("b200_rf_noise", [A, A, A, A, A, 2, 1020, None], "rf_noise"),
    ("b200_rf_loss", [A, A, A, A, 1024, 1.0, A, 16, None], "workspace too small"),
    ("b200_rf_loss", [A, A, A, A, 0, 1.0, A, 1 << 20, None], "null pointer or empty input"),
    ("b200_lerp_condition", [A, A, A, 1, 64, 128, 16, 0.85, 0.5, 32, 64, None], "lerp_condition"),
    ("b200_guidance_step", [A, A, None, 1, A, 0, None, A, 1, 64, 128, 0, 0, 0, 0, None, 0, None], "x_next is null"),
    ("b200_guidance_step", [A, A, None, 0, A, 0, None, A, 1, 64, 100, 0, 0, 0, 0, None, 0, None], "multiple of 8"),
    ("b200_adamw_step", [A, A, -1, A, A, A, None], "bad block count"),
    ("b200_adamw_step", [A, None, 1, A, A, A, None], "null pointer"),
    ("b200_adamw_step", [MIS, A, 1, A, A, A, None], "8-byte aligned"),
    ("b200_rowscale", [A, 2048, A, 2048, A, 2048, 4, 2048, 0, None], "rowscale"),
    ("b200_colsum", [A, 2048, None, 4, 2048, None], "colsum"),
    ("b200_colsum_groups", [A, 2048, None, 0, A, 6, 2048, 4, None, 0, None], "multiple of rows_per_group"),
    ("b200_attn_merge", [A, 64, A, None, 64, A, A, 64, 1, 1, 128, 1, None], "attn_merge"),
    ("b200_attn_delta", [A, 60, A, 64, A, 1, 1, 128, None], "attn_delta"),
    ("b200_attn_delta_zero", [A, 64, A, 64, A, A, 32, 1, 1, 128, None], "attn_delta"),
]


@pytest.mark.parametrize("name,args,fragment", CONTRACT_CASES, ids=[f"{c[0]}-{i}" for i, c in enumerate(CONTRACT_CASES)])
def test_every_entry_point_refuses_contract_violations(name, args, fragment):
    """include/b200ltx.h: a negative return = argument / shape / alignment violation, nothing launched, message in
    b200_last_error().  These run on a box without a GPU: were a case to get past the host-side checks it would reach
    the launch and come back POSITIVE (a cudaError), failing the assertion."""
    from b200_ltx import lib
    h = lib.load()
    fn = getattr(h, name)
    assert len(args) == len(lib.PROTOTYPES[name][1]), (name, len(args), len(lib.PROTOTYPES[name][1]))
    rc = fn(*args)
    msg = h.b200_last_error().decode()
    assert rc < 0, (name, rc, msg)
    assert fragment in msg, (name, fragment, msg)
    with pytest.raises(lib.B200Error, match="contract violation"):
        lib.check(rc, name)


def test_empty_problems_are_a_successful_no_op():
    """Zero-sized inputs (an empty batch, no rows) return 0 without touching a device: the reference's modules accept
    empty tensors, and a ragged last shard may be empty."""
    from b200_ltx import lib
    h = lib.load()
    assert h.b200_gemm_bf16(*_gemm_args(M=0)) == 0
    assert h.b200_gemm_bf16(*_gemm_args(N=0)) == 0
    assert h.b200_fa_bwd(A, 64, A, 64, A, 64, A, 64, A, A, None, A, 64, A, 64, A, 64, 0, 1, 128, 128, 64, 0.125, None, 0,
                         None) == 0
    assert h.b200_fa_bwd_workspace_bytes(0, 32, 6144, 6144) == 0
    assert h.b200_fa_fwd_workspace_bytes(1, 32, 6144, 6144) >= 0
    assert h.b200_gemm_workspace_bytes() > 16384 and h.b200_rf_loss_workspace_bytes() > 0
    assert h.b200_adamw_chunk_elems() > 0 and h.b200_guidance_step_workspace_bytes(4) > 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="only where no launch can happen (stand-in addresses)")
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_random_arguments_never_crash_the_host_side(seed):
    """8000 random argument lists per seed over twelve entry points (tests/native/abi_fuzz.py, in a subprocess so that a
    crash is a test failure, not the end of the run): only contract violations, empty no-ops and 'no device' errors."""
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "native", "abi_fuzz.py"), str(seed), "8000"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.returncode, r.stderr[-800:])
    counts = eval(r.stdout.strip().splitlines()[-1])     # {'neg': .., 'zero': .., 'pos': ..}
    assert sum(counts.values()) == 8000 and counts["neg"] > 6000 and counts["zero"] > 100 and counts["pos"] > 50, counts


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only behaviour")
def test_no_cpu_fallback():
    from b200_ltx import lib, ops
    with pytest.raises(lib.B200Error):
        lib.require_device()
    with pytest.raises(lib.B200Error):
        ops.gemm(torch.zeros(8, 8, dtype=torch.bfloat16), torch.zeros(8, 8, dtype=torch.bfloat16))


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: no module of the product package (nor the C sources) may import, load or
    execute anything under oracle/.  Checked on the source text and on the modules an import of the whole package
    pulls in."""
    import ast
    import sys
    pkg = os.path.join(ROOT, "video-generation-for-human-avatars_b200")
    banned = {"ref_block", "ref_sampling", "ref_import", "make_golden", "diffusers_shim", "peft_shim", "oracle"}
    for name in sorted(os.listdir(pkg)):
        if not name.endswith(".py"):
            continue
        tree = ast.parse(open(os.path.join(pkg, name)).read())
        for node in ast.walk(tree):
            mods = []
            if isinstance(node, ast.Import):
                mods = [a.name for a in node.names]
            elif isinstance(node, ast.ImportFrom):
                mods = [node.module or ""]
            for m in mods:
                assert not (set(m.split(".")) & banned), f"{name} imports {m}"
    for name in os.listdir(os.path.join(pkg, "csrc")):
        text = open(os.path.join(pkg, "csrc", name)).read()
        assert "oracle/" not in text or name.endswith(".md"), name
    before = set(sys.modules)
    import b200_ltx.api  # noqa: F401
    import b200_ltx.optim  # noqa: F401
    import b200_ltx.ring  # noqa: F401
    pulled = {m.split(".")[0] for m in set(sys.modules) - before}
    assert not (pulled & banned), pulled & banned
