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
