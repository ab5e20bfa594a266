"""The GEMM's persistent tile walk, checked exhaustively on the host.

`gemm_kernel` (csrc/gemm.cu) is persistent: every CTA (or CTA pair) derives the list of (tile, k-block range) units it
computes from its index alone -- grouped tile rasterisation, split-K slices, batch groups, and the stream-K deal of the
last partial wave with its owner / helper roles.  A unit visited twice or never is a wrong GEMM only for the shapes that
hit it.  The definitions are cut out of gemm.cu as they stand, compiled for the host (`__device__` defined away) and
driven over ~64 k parameter combinations: every k block of every tile exactly once, tile_coords a bijection, every
stream-K tile with exactly one epilogue owner and at most four helpers whose partial is their first unit (what the
kernel's flag protocol assumes), under the host's own admission rule for stream-K (gemm_impl)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GEMM = os.path.join(ROOT, "video-generation-for-human-avatars_b200", "csrc", "gemm.cu")
HARNESS = os.path.join(ROOT, "tests", "native", "gemm_walk_harness.cpp")


def _walk_source():
    src = open(GEMM).read()
    a = src.index("struct GemmParams {")
    b = src.index("__device__ __forceinline__ float tanh_fast")
    body = src[a:b]
    for needed in ("tile_coords(const GemmParams& p", "struct GemmUnit", "struct GemmWalk", "bool next(GemmUnit& u)"):
        assert needed in body, needed
    return body


def _build(tmp_path, body, name):
    cpp = tmp_path / f"{name}.cpp"
    cpp.write_text(open(HARNESS).read().replace("/*@@GEMM_WALK_SOURCE@@*/", body))
    exe = tmp_path / name
    r = subprocess.run(["g++", "-O1", "-std=c++17", "-o", str(exe), str(cpp)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    return exe


@pytest.mark.skipif(shutil.which("g++") is None, reason="g++ not available")
def test_tile_walk_covers_every_k_block_exactly_once(tmp_path):
    exe = _build(tmp_path, _walk_source(), "walk")
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.startswith("OK "), r.stdout[-500:]
    assert int(r.stdout.split()[1]) > 50000


@pytest.mark.skipif(shutil.which("g++") is None, reason="g++ not available")
@pytest.mark.parametrize("old,new", [
    ("    work += P;", "    work += P + (work > 40);"),                                      # a skipped work item
    ("u.kb_end = min(kb_all, u.kb_begin + (it_end - it));", "u.kb_end = min(kb_all, u.kb_begin + (it_end - it) + 1);"),
    ("tm = first + (in - tn * rows);", "tm = first + (in - tn * rows) % 2;"),                # rasterisation collision
])
def test_the_harness_catches_a_broken_walk(tmp_path, old, new):
    """The check must be able to fail: three one-line mutations of the walk are each detected."""
    body = _walk_source()
    assert old in body
    exe = _build(tmp_path, body.replace(old, new), "mutant")
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert r.returncode != 0 and "FAIL" in r.stdout, r.stdout[-300:]
