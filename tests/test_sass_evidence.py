"""What the shipped library is made of, read from its SASS (cuobjdump, no GPU needed).

The kernels are hand-written for sm_100a: the contractions must be tcgen05 (`UTCHMMA`, with `.2CTA` for the CTA-pair GEMM),
accumulators must travel through TMEM (`LDTM` / `STTM`), tiles through TMA (`UTMALDG` / `UTMASTG`, and `UTMAREDG` for the
dQ reduce-add of the attention backward), and no kernel may fall back to the warp-level `HMMA` path (mma.sync / wmma) or
to `LDGSTS` (cp.async) staging.  Mnemonics as listed in /opt/skills/guides/B200_PROFILING.md."""
import collections
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def parse_sass():
    from b200_ltx import build, lib
    if not os.path.exists(lib.LIB_PATH):
        build.build()
    out = subprocess.run(["cuobjdump", "-sass", lib.LIB_PATH], capture_output=True, text=True, timeout=300).stdout
    per = collections.OrderedDict()
    name = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            per[name] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and name:
            op = m.group(1)
            per[name][op.split(".")[0]] += 1
            if op.startswith("UTCHMMA.2CTA"):
                per[name]["UTCHMMA.2CTA"] += 1
            if op.startswith("MUFU.EX2"):
                per[name]["MUFU.EX2"] += 1
    return per


@pytest.fixture(scope="module")
def sass():
    return parse_sass()


def _kernels(per, fragment):
    hit = {k: v for k, v in per.items() if fragment in k}
    assert hit, fragment
    return hit


@pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not available")
def test_contractions_are_tcgen05_with_tmem_and_tma(sass):
    assert len(sass) >= 40                                   # every kernel of the library was disassembled
    for name, ops in sass.items():
        assert ops["HMMA"] == 0 and ops["IMMA"] == 0, f"{name}: warp-level mma.sync path"
        assert ops["LDGSTS"] == 0, f"{name}: cp.async staging"
    gemms = _kernels(sass, "gemm_kernel")
    assert len(gemms) == 16                                  # BN {64,128,256} x layouts + the four CTA-pair variants
    for name, ops in gemms.items():
        assert ops["UTCHMMA"] >= 4 and ops["UTMALDG"] >= 2 and ops["LDTM"] >= 1, (name, dict(ops))
    pair = {k: v for k, v in gemms.items() if v["UTCHMMA.2CTA"] > 0}
    assert len(pair) == 4 and all(v["UTMASTG"] >= 1 for v in pair.values())      # cta_group::2 + TMA stores
    fwd = _kernels(sass, "fa_fwd_db_kernel")
    for name, ops in fwd.items():
        assert ops["UTCHMMA"] >= 8 and ops["UTMALDG"] >= 3 and ops["LDTM"] >= 2 and ops["STTM"] >= 1, (name, dict(ops))
        assert ops["MUFU.EX2"] >= 8 and ops["FFMA2"] >= 8, (name, dict(ops))   # exponentials + packed fp32x2 softmax math
    bwd = _kernels(sass, "fa_bwd_kernel")
    for name, ops in bwd.items():
        assert ops["UTCHMMA"] >= 24 and ops["UTMALDG"] >= 4 and ops["LDTM"] >= 4 and ops["STTM"] >= 2, (name, dict(ops))
        assert ops["UTMAREDG"] >= 2, (name, dict(ops))        # dQ partials: TMA reduce-add into the fp32 accumulator


@pytest.mark.skipif(shutil.which("cuobjdump") is None or shutil.which("c++filt") is None,
                    reason="cuobjdump / c++filt not available")
def test_sass_summary_in_profiles_is_current(sass):
    """profiles/r2_sass_mnemonics.md is generated from the same parse (tools/sass_summary.py): the committed table must
    be the table of the library as it builds now."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import sass_summary
    want = sass_summary.table(sass)
    have = open(os.path.join(ROOT, "profiles", "r2_sass_mnemonics.md")).read()
    assert want in have
