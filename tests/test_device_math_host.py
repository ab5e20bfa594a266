"""Scalar device math whose accuracy claims need no GPU: the FMA-pipe 2^x of the attention forward (attn_fwd.cu:
`ex2_poly`, a degree-3 polynomial standing in for `ex2.approx` on one pair of every eight exponentials) and the GELU-tanh /
GELU' pair of the GEMM epilogues (gemm.cu).  The functions are cut out of the sources verbatim and compiled for the host."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "video-generation-for-human-avatars_b200", "csrc")


def _cut(text, begin, end):
    a = text.index(begin)
    return text[a:text.index(end, a) + len(end)]


def _run(tmp_path, name, defines=()):
    fwd = open(os.path.join(CSRC, "attn_fwd.cu")).read()
    gemm = open(os.path.join(CSRC, "gemm.cu")).read()
    text = open(os.path.join(ROOT, "tests", "native", "device_math_harness.cpp")).read()
    ex2 = _cut(fwd, "#ifdef B200_EX2_MINIMAX", "#endif\n") + _cut(fwd, "__device__ __forceinline__ float ex2_poly(float x) {", "\n}\n")
    text = text.replace("/*@@EX2_POLY@@*/", ex2)
    text = text.replace("/*@@GELU@@*/", _cut(gemm, "__device__ __forceinline__ float gelu_tanh(float x) {", "\n}\n")
                        + _cut(gemm, "__device__ __forceinline__ float gelu_tanh_grad(float x) {", "\n}\n"))
    cpp = tmp_path / f"{name}.cpp"
    cpp.write_text(text)
    exe = tmp_path / name
    # -ffp-contract=off: fmaf() is the only fused operation, as on the device
    r = subprocess.run(["g++", "-O1", "-std=c++17", "-ffp-contract=off", *[f"-D{d}" for d in defines], "-o", str(exe), str(cpp)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300).stdout
    return out, float(re.search(r"ex2_poly max_rel (\S+)", out).group(1)), float(re.search(r"gelu max_abs (\S+)", out).group(1))


@pytest.mark.skipif(shutil.which("g++") is None, reason="g++ not available")
def test_polynomial_exp2_and_gelu_derivative(tmp_path):
    out, rel, gelu = _run(tmp_path, "math")
    # Taylor coefficients: 7.9e-4 at |f| = 0.5 (the source comment records it); what it feeds is P rounded to bf16, half
    # an ulp of which is 2^-9 = 2e-3
    assert rel < 8.5e-4, out
    assert "ex2_poly clamp 1 one 1" in out, out
    assert gelu < 2e-5, out


@pytest.mark.skipif(shutil.which("g++") is None, reason="g++ not available")
def test_minimax_constants_behind_the_experiment_switch(tmp_path):
    """-DB200_EX2_MINIMAX (a build.py variant, not the default: no GPU run of the parity tests stands behind it yet):
    the same three FMAs with the minimax constants are eight times closer to exp2."""
    out, rel, _ = _run(tmp_path, "math_minimax", ["B200_EX2_MINIMAX"])
    assert rel < 1.1e-4, out
    assert "ex2_poly clamp 1 one 1" in out, out
