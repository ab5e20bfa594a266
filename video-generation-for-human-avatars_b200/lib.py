"""ctypes binding of libb200ltx.so (C ABI declared in include/b200ltx.h).

There is no fallback: if the shared library is missing this module raises, and on a device that is
not sm_100 every launch returns an error that `check()` turns into an exception."""
import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int64, c_void_p

import torch  # noqa: F401  (loads libcudart.so.12 that libb200ltx links against)

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200LTX_LIB") or os.path.join(PKG_DIR, "libb200ltx.so")  # env override: kernel experiments

P, I, L, F = c_void_p, c_int, c_int64, c_float

# name -> (restype, argtypes); order mirrors include/b200ltx.h
PROTOTYPES = {
    "b200_version": (I, []),
    "b200_last_error": (c_char_p, []),
    "b200_device_check": (I, []),
    "b200_gemm_bf16": (I, [P, L, I, P, L, I, P, L, P, L, I, P, L, I, I, I, I, I, P, P, L, L, P, L, P, L, I, I, P]),
    "b200_gemm_workspace_bytes": (L, []),
    "b200_gemm_bf16_ws": (I, [P, L, I, P, L, I, P, L, P, L, I, P, L, I, I, I, I, I, P, P, L, L, P, L, P, L, I, I, P, L, P]),
    "b200_gemm_bf16_batched": (I, [P, L, I, P, L, I, P, L, P, L, I, P, L, I, I, I, I, P, I, I, P, P]),
    "b200_fa_fwd": (I, [P, L, P, L, P, L, P, L, P, P, I, I, I, I, I, F, P]),
    "b200_fa_fwd_workspace_bytes": (L, [I, I, I, I]),
    "b200_fa_fwd_ws": (I, [P, L, P, L, P, L, P, L, P, P, P, P, L, I, I, I, I, I, F, P, L, P]),
    "b200_attn_merge": (I, [P, L, P, P, L, P, P, L, I, I, I, I, P]),
    "b200_attn_delta": (I, [P, L, P, L, P, I, I, I, P]),
    "b200_attn_delta_zero": (I, [P, L, P, L, P, P, L, I, I, I, P]),
    "b200_fa_bwd_workspace_bytes": (L, [I, I, I, I]),
    "b200_fa_bwd": (I, [P, L, P, L, P, L, P, L, P, P, P, P, L, P, L, P, L, I, I, I, I, I, F, P, L, P]),
    "b200_norm_mod_fwd": (I, [P, L, P, L, P, P, L, L, I, L, F, I, P]),
    "b200_norm_mod_bwd": (I, [P, L, P, L, P, L, P, L, P, L, P, L, L, I, L, F, I, P]),
    "b200_qknorm_rope_fwd": (I, [P, L, P, L, P, P, P, P, L, P, L, P, L, L, L, I, F, P]),
    "b200_qknorm_rope_bwd": (I, [P, L, I, P, L, I, P, L, P, L, P, P, P, P, L, P, L, P, L, P, L, P, L, L, L, I, F, P]),
    "b200_rf_noise": (I, [P, P, P, P, P, L, L, P]),
    "b200_rf_loss_workspace_bytes": (L, []),
    "b200_rf_loss": (I, [P, P, P, P, L, F, P, L, P]),
    "b200_lerp_condition": (I, [P, P, P, I, I, I, I, F, F, I, I, P]),
    "b200_guidance_step_workspace_bytes": (L, [I]),
    "b200_guidance_step": (I, [P, P, P, I, P, I, P, P, I, L, I, I, I, I, I, P, L, P]),
    "b200_adamw_chunk_elems": (I, []),
    "b200_adamw_step": (I, [P, P, I, P, P, P, P]),
    "b200_rowscale": (I, [P, L, P, L, P, L, L, I, L, P]),
    "b200_colsum": (I, [P, L, P, L, I, P]),
    "b200_colsum_groups_workspace_bytes": (L, [L, I, L]),
    "b200_colsum_groups": (I, [P, L, P, L, P, L, I, L, P, L, P]),
}

_lib = None


class B200Error(RuntimeError):
    pass


def load() -> ctypes.CDLL:
    """Load libb200ltx.so (once).  Raises if it has not been built (python -m b200_ltx.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise B200Error(f"{LIB_PATH} is missing: run `python __graft_entry__.py build` "
                            "(there is no CPU or PyTorch fallback for the b200_ltx kernels)")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().b200_last_error()
        kind = "contract violation" if rc < 0 else "CUDA error"
        raise B200Error(f"{what}: {kind} {rc}: {msg.decode() if msg else ''}")


def require_device() -> None:
    if not torch.cuda.is_available():
        raise B200Error("b200_ltx needs a CUDA device (sm_100a); no CPU fallback exists")
    check(load().b200_device_check(), "b200_device_check")
