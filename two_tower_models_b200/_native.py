"""ctypes binding of libtt_b200.so (the C ABI declared in include/tt_b200.h).

There is deliberately no fallback: if the library is missing, or a compute entry point is called
without a CUDA device, a RuntimeError is raised.
"""
import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int32, c_int64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
# TT_B200_LIB selects another build of the same sources (bring-up builds with extra -D flags); never a fallback
LIB_PATH = os.environ.get("TT_B200_LIB") or os.path.join(_HERE, "csrc", "libtt_b200.so")

_lib = None

P = c_void_p
I64 = c_int64
I32 = c_int32
F32 = c_float

class GemmProblem(ctypes.Structure):
    """struct tt_gemm_problem (include/tt_b200.h)."""

    _fields_ = [
        ("A", c_void_p), ("lda", c_int64), ("B", c_void_p), ("ldb", c_int64),
        ("M", c_int64), ("N", c_int64), ("K", c_int64),
        ("bias", c_void_p), ("relu_mask_bf16", c_void_p), ("ld_mask", c_int64),
        ("c_f32", c_void_p), ("ldc_f32", c_int64), ("c_bf16", c_void_p), ("ldc_bf16", c_int64),
        ("colsum_f32", c_void_p), ("alpha", c_float),
        ("a_mn_major", c_int32), ("b_mn_major", c_int32), ("relu", c_int32), ("accumulate", c_int32), ("split_k", c_int32),
    ]


class CastProblem(ctypes.Structure):
    """struct tt_cast_problem."""

    _fields_ = [("src", c_void_p), ("rows", c_int64), ("cols", c_int64), ("ld_src", c_int64),
                ("dst_bf16", c_void_p), ("ld_dst", c_int64), ("dst_cols", c_int64)]


class GatherProblem(ctypes.Structure):
    """struct tt_gather_problem."""

    _fields_ = [("table", c_void_p), ("table_rows", c_int64), ("dim", c_int64), ("ids", c_void_p), ("n", c_int64),
                ("dst_bf16", c_void_p), ("ld_dst", c_int64)]


class TowerProblem(ctypes.Structure):
    """struct tt_tower_problem."""

    _fields_ = [
        ("ids", c_void_p), ("table", c_void_p), ("table_rows", c_int64), ("feats", c_void_p), ("ld_feats", c_int64),
        ("w0_bf16", c_void_p), ("ldw0", c_int64), ("b0", c_void_p), ("w1_bf16", c_void_p), ("ldw1", c_int64),
        ("b1", c_void_p), ("wt_bf16", c_void_p), ("ldwt", c_int64), ("bt", c_void_p),
        ("feats_bf16", c_void_p), ("ld_feats16", c_int64), ("h_bf16", c_void_p), ("ldh", c_int64),
        ("x_bf16", c_void_p), ("ldx", c_int64), ("emb_f32", c_void_p), ("ld_emb", c_int64),
        ("emb_bf16", c_void_p), ("ld_emb16", c_int64),
        ("rows", c_int64), ("F", c_int64), ("D", c_int64), ("DI", c_int64), ("hidden", c_int64),
    ]


class TowerBwdProblem(ctypes.Structure):
    """struct tt_tower_bwd_problem."""

    _fields_ = [
        ("demb_bf16", c_void_p), ("ld_demb", c_int64), ("ids", c_void_p), ("table_rows", c_int64),
        ("wt_bf16", c_void_p), ("ldwt", c_int64), ("w1_bf16", c_void_p), ("ldw1", c_int64),
        ("h_bf16", c_void_p), ("ldh", c_int64), ("dx_bf16", c_void_p), ("lddx", c_int64),
        ("dh_bf16", c_void_p), ("lddh", c_int64), ("table_grad", c_void_p), ("dxsum", c_void_p), ("db0", c_void_p),
        ("rows", c_int64), ("D", c_int64), ("DI", c_int64), ("hidden", c_int64),
    ]


class AdamTensor(ctypes.Structure):
    """struct tt_adam_tensor."""

    _fields_ = [("param", c_void_p), ("grad", c_void_p), ("exp_avg", c_void_p), ("exp_avg_sq", c_void_p),
                ("numel", c_int64)]


# name -> (restype, argtypes); mirrors include/tt_b200.h one to one
SIGNATURES = {
    "tt_abi_version": (I32, []),
    "tt_last_error": (c_char_p, []),
    "tt_device_sm_count": (I32, []),
    "tt_launch_count": (I64, []),
    "tt_profile_enable": (None, [I32]),
    "tt_profile_report": (I32, [P, I64]),
    "tt_profile_report_graph": (I32, [P, I64, I32]),
    "tt_profile_null_span": (I32, [P]),
    "tt_cast_rows_bf16": (I32, [P, I64, I64, I64, P, I64, I64, P]),
    "tt_gather_rows_bf16": (I32, [P, I64, I64, P, I64, P, I64, P, P]),
    "tt_gather_rows_f32": (I32, [P, I64, I64, P, I64, P, I64, P, P]),
    "tt_scatter_add_rows": (I32, [P, P, I64, P, I64, I64, P, I64, P]),
    "tt_colsum": (I32, [P, P, I64, I64, I64, P, P]),
    "tt_gemm_bf16": (I32, [P, I64, I32, P, I64, I32, I64, I64, I64, P, I32, P, I64, F32, P, I64, P, I64, I32, I32, P, P]),
    "tt_gemm_bf16_batched": (I32, [P, I32, P]),
    "tt_cast_rows_bf16_batched": (I32, [P, I32, P]),
    "tt_gather_rows_bf16_batched": (I32, [P, I32, P, P]),
    "tt_tower_fwd_supported": (I32, [I64, I64, I64, I64]),
    "tt_tower_fwd": (I32, [P, I32, P, P]),
    "tt_tower_bwd_chain": (I32, [P, I32, P]),
    "tt_inbatch_ce_workspace_bytes": (I64, [I64, I64, I64]),
    "tt_inbatch_ce_fwd": (I32, [P, I64, P, I64, I64, I64, I64, I64, P, P, P, I64, P]),
    "tt_inbatch_ce_bwd": (I32, [P, I64, P, I64, I64, I64, I64, I64, P, P, P, I64, P, I64, P, I64, P, I64, P, P, P, I64, P]),
    "tt_inbatch_ce_fwd_parts": (I32, [P, I64, P, I32, I64, I64, I64, I64, I64, I64, P, P, P, I64, P]),
    "tt_inbatch_ce_bwd_parts": (I32, [P, I64, P, I32, I64, I64, I64, I64, I64, I64, P, P, P, I64, P, I64, P, I64, P, I64,
                                      P, P, P, I64, P]),
    "tt_inbatch_ce_loss_fwd": (I32, [P, I64, P, I64, I64, I64, I64, I64, P, I64, P, I64, P, P, P, P, P, P, I64, P]),
    "tt_inbatch_ce_bwd_scaled": (I32, [P, I64, P, I64, I64, I64, I64, I64, P, P, P, P, P, I64, P, I64, P, I64, P, I64, P, P,
                                       P, I64, P]),
    "tt_adam_step": (I32, [P, I32, c_double, c_double, c_double, F32, F32, P, P, P]),
    "tt_weighted_loss": (I32, [P, P, I64, P, I64, I64, P, P, P]),
    "tt_inbatch_ce_loss_fwd_sharded": (I32, [P, I64, P, I64, I64, I64, I64, I64, P, I64, P, I64, P, P, P, P, P, I64, P]),
    "tt_sharded_loss_finalize": (I32, [P, I32, I64, P, P, P]),
    "tt_set_sm_limit": (I32, [I32]),
    "tt_inbatch_ce_attach_zero_fill": (I32, [P, I64, P, I64]),
    "tt_history_last_supported": (I32, [I64, I64, I64]),
    "tt_history_last_fwd": (I32, [P, I64, P, I64, I64, I64, I64, P, P, P]),
    "tt_history_last_bwd1": (I32, [P, I64, P, P, I64, I64, I64, I64, P, P, P]),
    "tt_history_last_bwd2": (I32, [P, P, P, P, P, I64, I64, I64, I64, P, I64, P, P]),
    "tt_mips_workspace_bytes": (I64, [I64, I64, I64, I64]),
    "tt_mips_topk": (I32, [P, I64, P, I64, P, I64, P, I64, I64, I64, I64, I64, P, P, P, I64, P]),
    "tt_history_gather_pool": (I32, [P, I64, I64, P, I64, I64, P, P, I64, P, I64, P, P]),
    "tt_history_scatter_grad": (I32, [P, I64, P, I64, P, I64, I64, I64, P, I64, P]),
    "tt_attn_fwd": (I32, [P, I64, I64, I64, I64, I64, I64, P, I64, P]),
    "tt_attn_bwd": (I32, [P, I64, P, I64, I64, I64, I64, I64, I64, P, I64, P]),
}


def lib():
    """Load (once) and return the shared library; raise if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python two_tower_models_b200/csrc/build.py` "
            "(two_tower_models_b200 has no CPU or PyTorch fallback)"
        )
    l = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(l, name)  # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args
    if l.tt_abi_version() != 1:
        raise RuntimeError("libtt_b200.so ABI version mismatch")
    _lib = l
    return l


def check(rc, what=""):
    if rc != 0:
        msg = lib().tt_last_error().decode(errors="replace")
        raise RuntimeError(f"tt_b200 {what} failed (rc={rc}): {msg}")
