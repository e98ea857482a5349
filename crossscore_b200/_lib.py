"""ctypes binding of libcrossscore_sm100a.so (the C ABI in include/crossscore_b200.h).

There is no fallback: if the library is missing or the device is not a B200 the import-time /
call-time errors are raised to the caller.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("XS_LIB_PATH") or os.path.join(HERE, "libcrossscore_sm100a.so")  # override: development builds

DT_BF16, DT_F32, DT_TF32, DT_F16 = 0, 1, 2, 3
ACT_NONE, ACT_GELU, ACT_RELU, ACT_LEAKY = 0, 1, 2, 3
OP_PATCH_EMBED = 1
OP_SCORE_POSTPROCESS = 2

_p, _i, _f, _ll, _sz = C.c_void_p, C.c_int, C.c_float, C.c_longlong, C.c_size_t

# name -> (restype, argtypes); must list every symbol declared in include/crossscore_b200.h
SIGNATURES = {
    "xs_version": (_i, []),
    "xs_last_error": (C.c_char_p, []),
    "xs_device_check": (_i, []),
    "xs_workspace_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "xs_patch_embed": (_i, [_p, _p, _p, _p, _p, _sz, _i, _i, _i, _i, _p]),
    "xs_embed_cls_pos_ln": (_i, [_p, _i, _p, _p, _p, _p, _p, _f, _p, _i, _i, _i, _p]),
    "xs_layernorm": (_i, [_p, _p, _p, _p, _p, _f, _p, _p, _i, _i, _p]),
    "xs_final_ln_drop_cls_add_pe": (_i, [_p, _p, _p, _p, _f, _p, _p, _p, _p, _i, _i, _i, _i, _p]),
    "xs_pe_resample_bilinear_ac": (_i, [_p, _p, _i, _i, _i, _i, _i, _p]),
    "xs_pos_embed_resample_bicubic": (_i, [_p, _p, _i, _i, _i, _i, _i, _p]),
    "xs_pos_embed_resample_bicubic_steps": (_i, [_p, _p, _i, _i, _i, _i, _i, _f, _f, _p]),
    "xs_gemm_bias_act": (_i, [_p, _i, _p, _i, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p]),
    "xs_gemm_bias_residual": (_i, [_p, _i, _p, _i, _p, _p, _i, _i, _i, _i, _i, _p]),
    "xs_gemm_bias_residual_ln": (_i, [_p, _i, _p, _i, _p, _p, _i, _p, _p, _f, _p, _i, _i, _i, _i, _i, _p]),
    "xs_gemm_bias_residual_stats": (_i, [_p, _i, _p, _i, _p, _p, _i, _p, _i, _p, _f, _i, _i, _i, _i, _p]),
    "xs_row_stats": (_i, [_p, _p, _p, _f, _i, _p]),
    "xs_gemm_ln_folded": (_i, [_p, _i, _p, _i, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _p]),
    "xs_flash_attn": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _ll, _ll, _ll, _ll, _i, _i, _i, _f, _i, _p]),
    "xs_attn_set_optimistic": (None, [_i]),
    "xs_attn_set_layout": (None, [_i]),
    "xs_lse_merge": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _ll, _ll, _i, _p]),
    "xs_head_score_jigsaw": (_i, [_p, _i, _p, _i, _p, _p, _i, _i, _i, _i, _i, _f, _i, _p]),
    "xs_lse_merge_peers": (_i, [_p, _ll, _ll, _p, _p, _i, _i, _i, _i, _i, _i, _p]),
    "xs_preprocess_u8_resize_normalize": (_i, [_p, _i, _i, _i, _p, _i, _i, C.POINTER(C.c_float), _p]),
    "xs_score_postprocess": (_i, [_p, _i, _i, _i, _p, _p, _i, _p, _f, _f, _p, _sz, _p]),
    "xs_attn_probs_one_head": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _ll, _ll, _ll, _ll, _f, _i, _p]),
}

_lib = None


class XsError(RuntimeError):
    pass


def load():
    """dlopen the kernel library (raises if it has not been built: run crossscore_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise XsError(
                f"{LIB_PATH} not found: build it with `python -m crossscore_b200.build` "
                "(there is no CPU / PyTorch fallback for the CrossScore hot path)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def last_error() -> str:
    return load().xs_last_error().decode("utf-8", "replace")


def check(rc: int, what: str):
    if rc != 0:
        kind = "invalid argument" if rc < 0 else f"CUDA error {rc}"
        raise XsError(f"{what}: {kind}: {last_error()}")


def call(name: str, *args):
    """Call an int-returning entry point and raise XsError on a non-zero status."""
    check(getattr(load(), name)(*args), name)


_launches = 0


def count_launch(n: int = 1):
    global _launches
    _launches += n


def launches() -> int:
    return _launches
