"""Compile csrc/*.cu for sm_100a into libb200ltx.so (in-tree, next to this file).

nvcc cross-compiles without a GPU, so this runs in the CPU build container; the resulting .so is
git-ignored but travels to the GPU box with the repo snapshot."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
OBJ_DIR = os.path.join(PKG_DIR, "build")
LIB_PATH = os.path.join(PKG_DIR, "libb200ltx.so")
SOURCES = ["api.cu", "elementwise.cu", "guidance.cu", "adamw.cu", "gemm.cu", "attn_fwd.cu", "attn_bwd.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "--cudart", "shared", "-Xcompiler", "-fPIC"]


def _newer(src, dst):
    return (not os.path.exists(dst)) or os.path.getmtime(src) > os.path.getmtime(dst)


def _headers_mtime():
    return max(os.path.getmtime(os.path.join(CSRC, f)) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh")))


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ_DIR, exist_ok=True)
    hdr = _headers_mtime()
    jobs = []
    for s in SOURCES:
        src, obj = os.path.join(CSRC, s), os.path.join(OBJ_DIR, s[:-3] + ".o")
        if force or _newer(src, obj) or hdr > os.path.getmtime(obj):
            jobs.append([NVCC, *NVCC_FLAGS, "-c", src, "-o", obj])

    def run(cmd):
        if verbose:
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    with ThreadPoolExecutor(max_workers=min(4, len(jobs) or 1)) as ex:
        list(ex.map(run, jobs))
    objs = [os.path.join(OBJ_DIR, s[:-3] + ".o") for s in SOURCES]
    if jobs or not os.path.exists(LIB_PATH):
        run([NVCC, "-shared", "--cudart", "shared", "-o", LIB_PATH, *objs,
             "-Xlinker", "-rpath", "-Xlinker", "/usr/local/cuda/lib64"])
    return LIB_PATH


def build_variant(name: str, defines, sources=("attn_fwd.cu", "attn_bwd.cu")) -> str:
    """Kernel experiments: libb200ltx_<name>.so = the default objects with `sources` recompiled under extra -D flags.
    Selected at run time with B200LTX_LIB=<path> (lib.py); never loaded by default."""
    build()
    vdir = os.path.join(OBJ_DIR, "variant_" + name)
    os.makedirs(vdir, exist_ok=True)
    objs = []
    for s in SOURCES:
        obj = os.path.join(OBJ_DIR, s[:-3] + ".o")
        if s in sources:
            obj = os.path.join(vdir, s[:-3] + ".o")
            cmd = [NVCC, *NVCC_FLAGS, *[f"-D{d}" for d in defines], "-c", os.path.join(CSRC, s), "-o", obj]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        objs.append(obj)
    out = os.path.join(PKG_DIR, f"libb200ltx_{name}.so")
    r = subprocess.run([NVCC, "-shared", "--cudart", "shared", "-o", out, *objs, "-Xlinker", "-rpath", "-Xlinker",
                        "/usr/local/cuda/lib64"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    return out


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "variant":   # python -m b200_ltx.build variant <name> D1 D2=3 ...
        print(build_variant(sys.argv[2], sys.argv[3:]))
    else:
        print(build(force="--force" in sys.argv, verbose=True))
