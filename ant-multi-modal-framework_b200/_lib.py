"""ctypes binding of libb200mm.so (the C-ABI declared in include/b200mm.h).

The library is the ONLY implementation of the arithmetic: if it is missing, or a tensor is not a CUDA tensor, the call
raises — there is no CPU or eager-PyTorch fallback on the product path.
"""
import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int32, c_int64, c_uint64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200mm.so")

ACT_NONE, ACT_QUICKGELU, ACT_GELU_ERF = 0, 1, 2


class GemmArgs(Structure):
    _fields_ = [
        ("A", c_void_p), ("lda", c_int64), ("a_mn", c_int32),
        ("B", c_void_p), ("ldb", c_int64), ("b_mn", c_int32),
        ("D", c_void_p), ("ldd", c_int64), ("d_f32", c_int32),
        ("M", c_int64), ("N", c_int64), ("K", c_int64),
        ("alpha", c_float),
        ("bias", c_void_p), ("act", c_int32),
        ("aux_out", c_void_p),
        ("dact_in", c_void_p), ("ld_dact", c_int64),
        ("residual", c_void_p), ("ldr", c_int64),
        ("splits", c_int32),
        ("workspace", c_void_p), ("workspace_bytes", c_int64),
        ("drop_p", c_float), ("drop_seed", c_uint64),
    ]


_P = c_void_p
_F = POINTER(c_float)
_I64P = c_void_p  # int64 device arrays are passed as raw addresses

# name -> (restype, argtypes); every symbol of include/b200mm.h
SIGNATURES = {
    "b200mm_last_error": (c_char_p, []),
    "b200mm_version": (c_int32, []),
    "b200mm_check_device": (c_int32, []),
    "b200mm_gemm_workspace_bytes": (c_int64, [c_int64, c_int64, c_int32]),
    "b200mm_gemm_bf16": (c_int32, [POINTER(GemmArgs), _P]),
    "b200mm_layernorm_fwd": (c_int32, [_P, _P, _P, c_int64, _P, _P, _P, _P, _P, _P, c_int64, c_int32, c_float, _P]),
    "b200mm_layernorm_bwd": (c_int32, [_P, _P, _P, _P, _P, _P, _P, _P, _P, c_int64, c_int32, _P]),
    "b200mm_embed_layernorm_fwd": (c_int32, [_P, _I64P, _P, c_int64, _P, _I64P, _P, _P, _P, _P, _P, _P, c_int64, c_int32, c_float, _P]),
    "b200mm_attention_fwd": (c_int32, [_P, c_int64, c_int32, c_int32, c_int32, _P, c_int64, _P, _P, c_int32, c_int32, c_int32, c_int32, c_float, _P]),
    "b200mm_attention_bwd": (c_int32, [_P, c_int64, c_int32, c_int32, c_int32, _P, _P, c_int64, _P, _P, _P, _P, c_int32, c_int32, c_int32, c_int32, c_float, _P]),
    "b200mm_attention_fwd_dropout": (c_int32, [_P, c_int64, c_int32, c_int32, c_int32, _P, c_int64, _P, _P, c_int32, c_int32, c_int32, c_int32, c_float,
                                               c_float, c_uint64, _P]),
    "b200mm_attention_bwd_dropout": (c_int32, [_P, c_int64, c_int32, c_int32, c_int32, _P, _P, c_int64, _P, _P, _P, _P, c_int32, c_int32, c_int32,
                                               c_int32, c_float, c_float, c_uint64, _P]),
    "b200mm_attention_dropout_mask": (c_int32, [_P, c_int32, c_int32, c_int32, c_float, c_uint64, _P]),
    "b200mm_dropout": (c_int32, [_P, c_int64, _P, c_int64, c_int64, c_int32, c_float, c_uint64, _P]),
    "b200mm_contrast_num_tiles": (c_int32, [c_int64]),
    "b200mm_contrast_lse_partials": (c_int32, [_P, c_int64, _P, c_int64, c_int32, c_int64, c_int64, c_int64, c_float, c_int64, _P, _P, _P, _P]),
    "b200mm_contrast_lse_merge": (c_int32, [_P, _P, c_int32, _P, _P, c_int32, _P, c_int32, _P, _P, c_int64, _P]),
    "b200mm_contrast_softgrad": (c_int32, [_P, c_int64, _P, c_int64, c_int32, c_int64, c_int64, c_int64, c_int64, c_float, c_int64, _P, c_float, c_float, c_int32, _P, c_int64, _P, _P]),
    "b200mm_contrast_lse_partials_pair": (c_int32, [_P, c_int64, _P, c_int64, _P, c_int64, _P, c_int64, c_int64, c_int64, c_int64, c_float, _P, c_int64,
                                                    _P, _P, _P, _P, _P, _P, _P]),
    "b200mm_contrast_softgrad_pair": (c_int32, [_P, c_int64, _P, c_int64, _P, c_int64, _P, c_int64, c_int64, c_int64, c_int64, c_int64, c_float, _P,
                                                c_int64, _P, _P, _P, _P, c_float, _P, c_float, c_int32, c_int32, c_int32, c_int32, _P, _P, c_int64,
                                                _P, _P]),
    "b200mm_contrast_rank": (c_int32, [_P, c_int64, _P, c_int64, c_int32, c_int64, c_int64, c_int64, c_float, c_int64, _P, _P, _P, _P]),
    "b200mm_masked_mean_fwd": (c_int32, [_P, _P, _P, _P, c_int64, c_int32, c_int32, _P]),
    "b200mm_masked_mean_bwd": (c_int32, [_P, _P, _P, _P, c_int64, c_int32, c_int32, _P]),
    "b200mm_act_layernorm_fwd": (c_int32, [_P, c_int32, _P, _P, _P, _P, _P, c_int64, c_int32, c_float, _P]),
    "b200mm_act_layernorm_bwd": (c_int32, [_P, _P, c_int32, _P, _P, _P, _P, _P, _P, c_int64, c_int32, _P]),
    "b200mm_mask_rows": (c_int32, [_P, _P, _P, c_int64, c_int32, _P]),
    "b200mm_gather_rows": (c_int32, [_P, _I64P, _P, c_int64, c_int64, c_int32, _P]),
    "b200mm_relu_fwd": (c_int32, [_P, _P, c_int64, _P]),
    "b200mm_relu_bwd": (c_int32, [_P, _P, _P, c_int64, _P]),
    "b200mm_mil_nce_matrix_fwd": (c_int32, [_P, c_int64, _P, _P, _P, c_int32, _P]),
    "b200mm_mil_nce_matrix_bwd": (c_int32, [_P, c_int64, _P, _P, _P, _P, c_int32, _P]),
    "b200mm_xpos_apply": (c_int32, [_P, c_int64, c_int32, c_int32, _P, _P, _P, _P, c_int64, c_int32, c_int32, c_int32, c_int32, _P]),
    "b200mm_act_fwd": (c_int32, [_P, _P, c_int64, c_int32, _P]),
    "b200mm_rowsum_periodic": (c_int32, [_P, _P, c_int64, c_int32, c_int64, _P]),
    "b200mm_scatter_add_rows": (c_int32, [_P, _I64P, _P, c_int64, c_int32, c_int64, c_int64, _P]),
    "b200mm_cast_f32_bf16": (c_int32, [_P, _P, c_int64, c_float, _P]),
    "b200mm_rownorm_fwd": (c_int32, [_P, _P, _P, c_int64, c_int32, c_float, _P]),
    "b200mm_rownorm_bwd": (c_int32, [_P, _P, _P, _P, c_int64, c_int32, _P]),
    "b200mm_im2row": (c_int32, [_P, _P, c_int64, c_int32, c_int32, c_int32, c_int32, c_int32, _P]),
    "b200mm_rowdot": (c_int32, [_P, _P, _P, c_int64, c_int32, c_float, _P]),
    "b200mm_ema_update": (c_int32, [_P, _P, c_int32, c_int64, c_float, _P]),
}

_lib = None


def load():
    """Loads libb200mm.so (once). Raises with a build hint if it is missing — never falls back."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise RuntimeError(
            f"b200mm: native library {LIB_PATH} not found. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C ant-multi-modal-framework_b200/csrc`. There is no CPU/PyTorch fallback."
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class B200mmError(RuntimeError):
    pass


def check(rc, what):
    if rc != 0:
        msg = load().b200mm_last_error()
        raise B200mmError(f"{what} failed (rc={rc}): {msg.decode() if msg else '?'}")


# ---------------------------------------------------------------------------------------------------------------------
# optional per-call CUDA-event timing of every C-ABI launch (bench.py --profile); off by default, zero overhead when off
# ---------------------------------------------------------------------------------------------------------------------
_PROFILE_RECORDS = None


def enable_profile(records):
    """Every compute entry point is wrapped so that (name, start_event, stop_event) is appended to `records`."""
    global _PROFILE_RECORDS
    import torch

    lib = load()
    _PROFILE_RECORDS = records
    for name in SIGNATURES:
        if name in ("b200mm_last_error", "b200mm_version", "b200mm_check_device", "b200mm_gemm_workspace_bytes", "b200mm_contrast_num_tiles"):
            continue
        raw = getattr(lib, "_raw_" + name, None) or getattr(lib, name)
        setattr(lib, "_raw_" + name, raw)

        def make(raw_fn, nm):
            def wrapped(*args):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                rc = raw_fn(*args)
                e1.record()
                _PROFILE_RECORDS.append((nm, e0, e1))
                return rc

            return wrapped

        setattr(lib, name, make(raw, name))


def disable_profile():
    global _PROFILE_RECORDS
    lib = load()
    for name in SIGNATURES:
        raw = getattr(lib, "_raw_" + name, None)
        if raw is not None:
            setattr(lib, name, raw)
    _PROFILE_RECORDS = None
