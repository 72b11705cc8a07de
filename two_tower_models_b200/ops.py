"""Tensor-level wrappers over the C ABI and the autograd Functions built from them.

PyTorch is used here for device memory (torch.empty), the current CUDA stream and autograd
bookkeeping only; every FLOP of the hot path runs in libtt_b200.so.  Nothing in this file has a CPU
or library fallback: tensors must live on a CUDA device.
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import torch

from . import _native

_BF16 = torch.bfloat16
_CHECK_IDS = os.environ.get("TT_B200_CHECK_IDS", "0") == "1"


def _r8(n: int) -> int:
    return (n + 7) // 8 * 8


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError(
                "two_tower_models_b200 runs on CUDA (sm_100a) only: got a CPU tensor; there is no CPU fallback"
            )


def _f32c(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()


def _i64c(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.int64:
        t = t.long()
    return t if t.is_contiguous() else t.contiguous()


# --------------------------------------------------------------------------------------------
# raw ops
# --------------------------------------------------------------------------------------------
def cast_rows_bf16(src: torch.Tensor, out: Optional[torch.Tensor] = None, col_offset: int = 0,
                   src_col0: int = 0, cols: Optional[int] = None) -> torch.Tensor:
    """bf16 copy of fp32 src[:, src_col0:src_col0+cols] into out[:, col_offset:...] (pitch-padded to 8)."""
    _need_cuda(src)
    src = _f32c(src)
    rows, total = src.shape
    cols = total - src_col0 if cols is None else cols
    if out is None:
        out = torch.empty((rows, _r8(cols)), dtype=_BF16, device=src.device)
        dst_cols = out.shape[1]
    else:
        dst_cols = cols
    L = _native.lib()
    _native.check(
        L.tt_cast_rows_bf16(src.data_ptr() + 4 * src_col0, rows, cols, src.stride(0),
                            out.data_ptr() + 2 * col_offset, out.stride(0), dst_cols, _stream()),
        "cast_rows_bf16",
    )
    return out


def gemm(A: torch.Tensor, B: torch.Tensor, M: int, N: int, K: int, *, a_mn=False, b_mn=False,
         bias: Optional[torch.Tensor] = None, relu=False, relu_mask: Optional[torch.Tensor] = None,
         alpha: float = 1.0, out32: Optional[torch.Tensor] = None, out16: Optional[torch.Tensor] = None,
         accumulate=False, split_k=0) -> None:
    """C = alpha*A*B^T (+bias)(relu)(mask); A,B bf16 2-D views (last stride 1); see tt_gemm_bf16."""
    L = _native.lib()
    _native.check(
        L.tt_gemm_bf16(A.data_ptr(), A.stride(0), int(a_mn), B.data_ptr(), B.stride(0), int(b_mn), M, N, K,
                       _ptr(bias), int(relu), _ptr(relu_mask), relu_mask.stride(0) if relu_mask is not None else 0,
                       alpha, _ptr(out32), out32.stride(0) if out32 is not None else 0,
                       _ptr(out16), out16.stride(0) if out16 is not None else 0,
                       int(accumulate), split_k, _stream()),
        "gemm_bf16",
    )


def colsum(src: torch.Tensor, cols: int) -> torch.Tensor:
    out = torch.zeros(cols, dtype=torch.float32, device=src.device)
    L = _native.lib()
    if src.dtype == _BF16:
        rc = L.tt_colsum(src.data_ptr(), None, src.shape[0], cols, src.stride(0), out.data_ptr(), _stream())
    else:
        rc = L.tt_colsum(None, src.data_ptr(), src.shape[0], cols, src.stride(0), out.data_ptr(), _stream())
    _native.check(rc, "colsum")
    return out


_oob_flags = {}


def _oob_flag(device) -> torch.Tensor:
    f = _oob_flags.get(device)
    if f is None:
        f = torch.zeros(1, dtype=torch.int32, device=device)
        _oob_flags[device] = f
    return f


def _maybe_check_ids(device, what):
    if _CHECK_IDS and int(_oob_flag(device).item()) != 0:
        _oob_flag(device).zero_()
        raise IndexError(f"{what}: index out of range in embedding lookup")


def gather_rows(table: torch.Tensor, ids: torch.Tensor, out: torch.Tensor, col_offset: int = 0) -> None:
    """out[:, col_offset:col_offset+dim] = table[ids] (out is bf16 or fp32)."""
    L = _native.lib()
    n, dim = ids.numel(), table.shape[1]
    flag = _oob_flag(table.device)
    if out.dtype == _BF16:
        rc = L.tt_gather_rows_bf16(table.data_ptr(), table.shape[0], dim, ids.data_ptr(), n,
                                   out.data_ptr() + 2 * col_offset, out.stride(0), flag.data_ptr(), _stream())
    else:
        rc = L.tt_gather_rows_f32(table.data_ptr(), table.shape[0], dim, ids.data_ptr(), n,
                                  out.data_ptr() + 4 * col_offset, out.stride(0), flag.data_ptr(), _stream())
    _native.check(rc, "gather_rows")
    _maybe_check_ids(table.device, "gather_rows")


def scatter_add_rows(src: torch.Tensor, ids: torch.Tensor, dim: int, table_rows: int, col_offset: int = 0,
                     grad: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Dense embedding gradient [table_rows, dim] (+)= rows of src[:, col_offset:col_offset+dim]."""
    if grad is None:
        grad = torch.zeros((table_rows, dim), dtype=torch.float32, device=src.device)
    L = _native.lib()
    es = src.element_size()
    p = src.data_ptr() + es * col_offset
    if src.dtype == _BF16:
        rc = L.tt_scatter_add_rows(p, None, src.stride(0), ids.data_ptr(), ids.numel(), dim, grad.data_ptr(), table_rows, _stream())
    else:
        rc = L.tt_scatter_add_rows(None, p, src.stride(0), ids.data_ptr(), ids.numel(), dim, grad.data_ptr(), table_rows, _stream())
    _native.check(rc, "scatter_add_rows")
    return grad


# --------------------------------------------------------------------------------------------
# packed bf16 weights (refreshed when the fp32 master parameter changes)
# --------------------------------------------------------------------------------------------
class PackedWeights:
    """bf16 shadow copies of fp32 master weights, re-packed lazily on parameter version change."""

    def __init__(self):
        self._cache = {}

    def get(self, key, param: torch.Tensor, segments=None) -> torch.Tensor:
        """segments: list of (src_col0, cols, dst_col0) placing column blocks at 8-aligned offsets."""
        ver = (param.data_ptr(), param._version, tuple(param.shape), str(param.device))
        hit = self._cache.get(key)
        if hit is not None and hit[0] == ver:
            return hit[1]
        rows, cols = param.shape
        with torch.no_grad():
            src = _f32c(param.detach())
            if segments is None:
                out = cast_rows_bf16(src)
            else:
                width = max(d0 + _r8(c) for (_, c, d0) in segments)
                out = hit[1] if (hit is not None and hit[1].shape == (rows, width)) else torch.zeros(
                    (rows, width), dtype=_BF16, device=param.device)
                for (s0, c, d0) in segments:
                    cast_rows_bf16(src, out=out, col_offset=d0, src_col0=s0, cols=c)
        self._cache[key] = (ver, out)
        return out


# --------------------------------------------------------------------------------------------
# tower: [id_emb | MLP(feats) | extra] -> Linear     (reference src/two_tower_base_retrieval.py:112-219)
# --------------------------------------------------------------------------------------------
def _mlp_forward(feats16, F, w0_16, b0, w1_16, b1, D, out16=None, out16_col=0, out32=None):
    """H = relu(feats W0^T + b0) (bf16, kept for backward);  Fe = H W1^T + b1 -> out16 view / out32."""
    rows = feats16.shape[0]
    hid = w0_16.shape[0]
    H16 = torch.empty((rows, hid), dtype=_BF16, device=feats16.device)
    gemm(feats16, w0_16, rows, hid, F, bias=b0, relu=True, out16=H16)
    o16 = None if out16 is None else out16[:, out16_col:]
    gemm(H16, w1_16, rows, D, hid, bias=b1, out16=o16, out32=out32)
    return H16


def _mlp_backward(dFe16, D, feats16, F, H16, w0_16, w1_16, need_dfeats=False):
    """Gradients of the feature MLP given dFe (bf16 view [rows, >=D])."""
    rows = feats16.shape[0]
    hid = H16.shape[1]
    dev = feats16.device
    dW1 = torch.zeros((D, hid), dtype=torch.float32, device=dev)
    gemm(dFe16, H16, D, hid, rows, a_mn=True, b_mn=True, out32=dW1, accumulate=True)
    db1 = colsum(dFe16, D)
    dH16 = torch.empty((rows, hid), dtype=_BF16, device=dev)
    gemm(dFe16, w1_16, rows, hid, D, b_mn=True, relu_mask=H16, out16=dH16)
    dW0 = torch.zeros((hid, F), dtype=torch.float32, device=dev)
    gemm(dH16, feats16, hid, F, rows, a_mn=True, b_mn=True, out32=dW0, accumulate=True)
    db0 = colsum(dH16, hid)
    dfeats = None
    if need_dfeats:
        dfeats = torch.empty((rows, F), dtype=torch.float32, device=dev)
        gemm(dH16, w0_16, rows, F, hid, b_mn=True, out32=dfeats)
    return dW0, db0, dW1, db1, dfeats


class TowerFunction(torch.autograd.Function):
    """emb = [table[ids] | MLP(feats) | extra] @ Wt^T + bt, fused on the device.

    The concatenation is never materialised in fp32: the three blocks are written as bf16 column
    segments (8-aligned offsets) of one operand buffer X and the tower Linear is a single GEMM over it.
    Returns fp32 emb [B, DI]; a bf16 copy is attached as `emb._tt_bf16` for the scoring kernels.
    """

    @staticmethod
    def forward(ctx, ids, feats, extra, table, w0, b0, w1, b1, wt, bt, packed: PackedWeights, tag: str):
        _need_cuda(ids, feats, table, w0, wt)
        ids = _i64c(ids)
        feats = _f32c(feats)
        B, F = feats.shape
        D = table.shape[1]
        DI = wt.shape[0]
        E = 0 if extra is None else extra.shape[1]
        D8 = _r8(D)
        KT = 2 * D8 + _r8(E)
        assert wt.shape[1] == 2 * D + E, "tower weight does not match [id_emb | feat_emb | extra]"
        segs = [(0, D, 0), (D, D, D8)] + ([(2 * D, E, 2 * D8)] if E else [])
        w0_16 = packed.get(tag + ".w0", w0)
        w1_16 = packed.get(tag + ".w1", w1)
        wt_16 = packed.get(tag + ".wt", wt, segments=segs)
        feats16 = cast_rows_bf16(feats)
        dense = (D8 == D) and (_r8(E) == E)
        X16 = (torch.empty if dense else torch.zeros)((B, KT), dtype=_BF16, device=feats.device)
        gather_rows(table, ids, X16, 0)
        H16 = _mlp_forward(feats16, F, w0_16, _f32c(b0), w1_16, _f32c(b1), D, out16=X16, out16_col=D8)
        if E:
            cast_rows_bf16(_f32c(extra), out=X16, col_offset=2 * D8)
        emb = torch.empty((B, DI), dtype=torch.float32, device=feats.device)
        emb16 = torch.empty((B, _r8(DI)), dtype=_BF16, device=feats.device)
        gemm(X16, wt_16, B, DI, KT, bias=_f32c(bt), out32=emb, out16=emb16)
        ctx.save_for_backward(ids, feats16, H16, X16, w0_16, w1_16, wt_16)
        ctx.dims = (B, F, D, DI, E, D8, KT, table.shape[0])
        ctx.need_dfeats = feats.requires_grad
        ctx.has_extra = extra is not None
        emb._tt_bf16 = emb16
        return emb

    @staticmethod
    def backward(ctx, demb):
        ids, feats16, H16, X16, w0_16, w1_16, wt_16 = ctx.saved_tensors
        B, F, D, DI, E, D8, KT, table_rows = ctx.dims
        dev = demb.device
        demb16 = getattr(demb, "_tt_bf16", None)
        if demb16 is None:
            demb16 = cast_rows_bf16(_f32c(demb))
        # tower Linear
        dWt_p = torch.zeros((DI, KT), dtype=torch.float32, device=dev)
        gemm(demb16, X16, DI, KT, B, a_mn=True, b_mn=True, out32=dWt_p, accumulate=True)
        dbt = colsum(demb16, DI)
        dX16 = torch.empty((B, KT), dtype=_BF16, device=dev)
        gemm(demb16, wt_16, B, KT, DI, b_mn=True, out16=dX16)
        if D8 == D and _r8(E) == E:
            dWt = dWt_p
        else:
            parts = [dWt_p[:, :D], dWt_p[:, D8:D8 + D]] + ([dWt_p[:, 2 * D8:2 * D8 + E]] if E else [])
            dWt = torch.cat(parts, dim=1)
        # id embedding (dense gradient, duplicates accumulate)
        dtable = scatter_add_rows(dX16, ids, D, table_rows, col_offset=0)
        # feature MLP
        dW0, db0, dW1, db1, dfeats = _mlp_backward(dX16[:, D8:], D, feats16, F, H16, w0_16, w1_16, ctx.need_dfeats)
        dextra = dX16[:, 2 * D8:2 * D8 + E].float() if ctx.has_extra else None
        return None, dfeats, dextra, dtable, dW0, db0, dW1, db1, dWt, dbt, None, None


class EmbeddingFunction(torch.autograd.Function):
    """nn.Embedding lookup with a dense gradient (used by get_user_embedding / process_user_features)."""

    @staticmethod
    def forward(ctx, ids, table):
        _need_cuda(ids, table)
        ids_flat = _i64c(ids).reshape(-1)
        out = torch.empty((ids_flat.numel(), table.shape[1]), dtype=torch.float32, device=table.device)
        gather_rows(_f32c(table), ids_flat, out)
        ctx.save_for_backward(ids_flat)
        ctx.table_shape = tuple(table.shape)
        return out.reshape(*ids.shape, table.shape[1])

    @staticmethod
    def backward(ctx, dout):
        (ids_flat,) = ctx.saved_tensors
        rows, dim = ctx.table_shape
        d = _f32c(dout.reshape(-1, dim))
        return None, scatter_add_rows(d, ids_flat, dim, rows)


class FeatureMLPFunction(torch.autograd.Function):
    """Linear(F,256) -> ReLU -> Linear(256,D) as two fused GEMM+bias(+ReLU) launches."""

    @staticmethod
    def forward(ctx, feats, w0, b0, w1, b1, packed: PackedWeights, tag: str):
        _need_cuda(feats, w0, w1)
        feats = _f32c(feats)
        B, F = feats.shape
        D = w1.shape[0]
        w0_16 = packed.get(tag + ".w0", w0)
        w1_16 = packed.get(tag + ".w1", w1)
        feats16 = cast_rows_bf16(feats)
        out = torch.empty((B, D), dtype=torch.float32, device=feats.device)
        H16 = _mlp_forward(feats16, F, w0_16, _f32c(b0), w1_16, _f32c(b1), D, out32=out)
        ctx.save_for_backward(feats16, H16, w0_16, w1_16)
        ctx.dims = (F, D)
        ctx.need_dfeats = feats.requires_grad
        return out

    @staticmethod
    def backward(ctx, dout):
        feats16, H16, w0_16, w1_16 = ctx.saved_tensors
        F, D = ctx.dims
        d16 = cast_rows_bf16(_f32c(dout))
        dW0, db0, dW1, db1, dfeats = _mlp_backward(d16, D, feats16, F, H16, w0_16, w1_16, ctx.need_dfeats)
        return dfeats, dW0, db0, dW1, db1, None, None


class LinearFunction(torch.autograd.Function):
    """y = x W^T + b on the tcgen05 GEMM (x fp32 in, fp32 out + bf16 shadow)."""

    @staticmethod
    def forward(ctx, x, w, b, packed: PackedWeights, tag: str):
        _need_cuda(x, w)
        x = _f32c(x)
        B, K = x.shape
        N = w.shape[0]
        w16 = packed.get(tag, w)
        x16 = cast_rows_bf16(x)
        y = torch.empty((B, N), dtype=torch.float32, device=x.device)
        y16 = torch.empty((B, _r8(N)), dtype=_BF16, device=x.device)
        gemm(x16, w16, B, N, K, bias=None if b is None else _f32c(b), out32=y, out16=y16)
        ctx.save_for_backward(x16, w16)
        ctx.dims = (B, K, N)
        ctx.has_bias = b is not None
        ctx.need_dx = x.requires_grad
        y._tt_bf16 = y16
        return y

    @staticmethod
    def backward(ctx, dy):
        x16, w16 = ctx.saved_tensors
        B, K, N = ctx.dims
        dy16 = getattr(dy, "_tt_bf16", None)
        if dy16 is None:
            dy16 = cast_rows_bf16(_f32c(dy))
        dW = torch.zeros((N, K), dtype=torch.float32, device=dy.device)
        gemm(dy16, x16, N, K, B, a_mn=True, b_mn=True, out32=dW, accumulate=True)
        db = colsum(dy16, N) if ctx.has_bias else None
        dx = None
        if ctx.need_dx:
            dx = torch.empty((B, K), dtype=torch.float32, device=dy.device)
            gemm(dy16, w16, B, K, N, b_mn=True, out32=dx)
        return dx, dW, db, None, None


# --------------------------------------------------------------------------------------------
# in-batch sampled-softmax cross entropy   (reference :287, :301, :310-312)
# --------------------------------------------------------------------------------------------
def _ce_workspace(B, N, d, device):
    nbytes = int(_native.lib().tt_inbatch_ce_workspace_bytes(B, N, d))
    return torch.empty(nbytes, dtype=torch.uint8, device=device)


def inbatch_ce_forward_raw(U16, V16, B, N, d, target_offset=0):
    ce = torch.empty(B, dtype=torch.float32, device=U16.device)
    lse = torch.empty(B, dtype=torch.float32, device=U16.device)
    ws = _ce_workspace(B, N, d, U16.device)
    _native.check(
        _native.lib().tt_inbatch_ce_fwd(U16.data_ptr(), U16.stride(0), V16.data_ptr(), V16.stride(0), B, N, d,
                                        target_offset, ce.data_ptr(), lse.data_ptr(), ws.data_ptr(), ws.numel(), _stream()),
        "inbatch_ce_fwd",
    )
    return ce, lse


def inbatch_ce_backward_raw(U16, V16, B, N, d, target_offset, lse, g, want_bf16=True):
    dev = U16.device
    dU = torch.empty((B, d), dtype=torch.float32, device=dev)
    dV = torch.empty((N, d), dtype=torch.float32, device=dev)
    dU16 = torch.empty((B, _r8(d)), dtype=_BF16, device=dev) if want_bf16 else None
    dV16 = torch.empty((N, _r8(d)), dtype=_BF16, device=dev) if want_bf16 else None
    ws = _ce_workspace(B, N, d, dev)
    _native.check(
        _native.lib().tt_inbatch_ce_bwd(
            U16.data_ptr(), U16.stride(0), V16.data_ptr(), V16.stride(0), B, N, d, target_offset,
            lse.data_ptr(), g.data_ptr(), dU.data_ptr(), dU.stride(0), _ptr(dU16), dU16.stride(0) if want_bf16 else 0,
            dV.data_ptr(), dV.stride(0), _ptr(dV16), dV16.stride(0) if want_bf16 else 0,
            ws.data_ptr(), ws.numel(), _stream()),
        "inbatch_ce_bwd",
    )
    return dU, dV, dU16, dV16


class InBatchCEFunction(torch.autograd.Function):
    """ce[B] = cross_entropy(U V^T, arange(B) + target_offset, reduction='none'), fused (no [B,N] matrix)."""

    @staticmethod
    def forward(ctx, U, V, target_offset: int = 0):
        _need_cuda(U, V)
        B, d = U.shape
        N = V.shape[0]
        if V.shape[1] != d:
            raise RuntimeError(f"user/item embedding dims differ: {d} vs {V.shape[1]}")
        U16 = getattr(U, "_tt_bf16", None)
        V16 = getattr(V, "_tt_bf16", None)
        if U16 is None:
            U16 = cast_rows_bf16(_f32c(U))
        if V16 is None:
            V16 = cast_rows_bf16(_f32c(V))
        ce, lse = inbatch_ce_forward_raw(U16, V16, B, N, d, target_offset)
        ctx.save_for_backward(U16, V16, lse)
        ctx.dims = (B, N, d, target_offset)
        return ce

    @staticmethod
    def backward(ctx, g):
        U16, V16, lse = ctx.saved_tensors
        B, N, d, off = ctx.dims
        dU, dV, dU16, dV16 = inbatch_ce_backward_raw(U16, V16, B, N, d, off, lse, _f32c(g))
        dU._tt_bf16 = dU16
        dV._tt_bf16 = dV16
        return dU, dV, None


def inbatch_cross_entropy(U: torch.Tensor, V: torch.Tensor, target_offset: int = 0) -> torch.Tensor:
    return InBatchCEFunction.apply(U, V, target_offset)
