"""How the attention kernels deal (batch, head, tile) work items to CTAs, checked on the host.

attn_fwd.cu / attn_bwd.cu size their grids from small cost models (`fa_fwd_plan`: the items of a sparsely filled last wave
are split along the keys; `fa_bwd_tail_plan`: along the query walk; `fa_bwd_splits`: few key tiles -> several CTAs per key
tile) and every CTA decodes its item, its part and its range of the walk from `blockIdx` alone.  The planning functions and
the decode blocks are cut out of the sources as they stand and driven over ~15 k shapes (tests/native/attn_deal_harness.cpp):
every item computed, ranges non-empty, disjoint and tiling the walk, one workspace slot per part."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "video-generation-for-human-avatars_b200", "csrc")
HARNESS = os.path.join(ROOT, "tests", "native", "attn_deal_harness.cpp")


def _cut(text, begin, end, include_end=False):
    a = text.index(begin)
    b = text.index(end, a)
    return text[a:b + (len(end) if include_end else 0)]


def _sources():
    fwd = open(os.path.join(CSRC, "attn_fwd.cu")).read()
    bwd = open(os.path.join(CSRC, "attn_bwd.cu")).read()
    return {
        "/*@@FA_FWD_PLAN@@*/": _cut(fwd, "static void fa_fwd_plan(", "\n}\n", True),
        "/*@@FA_BWD_PLANS@@*/": _cut(bwd, "static void fa_bwd_tail_plan(", "\n}\n", True) + "\n"
                                + _cut(bwd, "static int fa_bwd_splits(", "\n}\n", True),
        "/*@@FA_FWD_DECODE@@*/": _cut(fwd, "  int item = blockIdx.x, part = -1, j_begin = 0, j_end = p.kv_tiles;",
                                      "b = item / (p.q_tiles * p.H);", True),
        "/*@@FA_BWD_DECODE@@*/": _cut(bwd, "  int kt, h, b, split, splits = p.q_splits, tail_slot = -1;",
                                      "const int T = (int)((int64_t)(split + 1) * p.q_tiles / splits) - t0;", True),
    }


def _build(tmp_path, parts, name):
    text = open(HARNESS).read()
    for marker, body in parts.items():
        assert marker in text
        text = text.replace(marker, body)
    cpp = tmp_path / f"{name}.cpp"
    cpp.write_text(text)
    exe = tmp_path / name
    r = subprocess.run(["g++", "-O1", "-std=c++17", "-o", str(exe), str(cpp)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return exe


@pytest.mark.skipif(shutil.which("g++") is None, reason="g++ not available")
def test_every_attention_work_item_is_dealt_exactly_once(tmp_path):
    exe = _build(tmp_path, _sources(), "deal")
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.startswith("OK "), r.stdout[-500:]
    assert int(r.stdout.split()[1]) > 10000


@pytest.mark.skipif(shutil.which("g++") is None, reason="g++ not available")
@pytest.mark.parametrize("marker,old,new", [
    ("/*@@FA_FWD_DECODE@@*/", "part = e % p.parts;", "part = e % p.parts; if (part == 1) part = 0;"),
    ("/*@@FA_BWD_DECODE@@*/", "splits = p.tail_parts;", "splits = p.tail_parts + 1;"),
    ("/*@@FA_BWD_PLANS@@*/", "int max_pr = q_tiles / 8 < 4 ? q_tiles / 8 : 4;", "int max_pr = 4 * q_tiles + 4;"),
])
def test_the_harness_catches_a_broken_deal(tmp_path, marker, old, new):
    parts = _sources()
    assert old in parts[marker]
    parts[marker] = parts[marker].replace(old, new)
    exe = _build(tmp_path, parts, "mutant")
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert r.returncode != 0 and "FAIL" in r.stdout, r.stdout[-300:]
