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
# Out-of-range embedding ids: the gather kernels clamp them and raise a device flag (nn.Embedding raises IndexError).
# TT_B200_CHECK_IDS=1 reads the flag after EVERY lookup (a host sync per call); by default it is read every
# _CHECK_IDS_EVERY-th lookup outside stream capture - bad ids surface within a few steps at ~no cost - and on demand
# through check_ids().  TT_B200_CHECK_IDS=0 turns the periodic check off.
_CHECK_IDS = os.environ.get("TT_B200_CHECK_IDS", "")
_CHECK_IDS_EVERY = 64
_check_ids_calls = 0


class KernelTimer:
    """Optional CUDA-event spans around named native calls (bench.py's per-kernel roofline numbers).

    Events are recorded on the launching (current) stream, so the spans measure device time of exactly
    the kernels enqueued inside them.  Disabled (None) by default: zero overhead on the product path."""

    def __init__(self):
        self.spans = {}

    def begin(self, name):
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record(torch.cuda.current_stream())
        self.spans.setdefault(name, []).append((e0, e1))
        return e1

    def totals_ms(self):
        torch.cuda.synchronize()
        return {k: (sum(a.elapsed_time(b) for a, b in v), len(v)) for k, v in self.spans.items()}


TIMER: Optional[KernelTimer] = None


class _span:
    def __init__(self, name):
        self.name = name
        self.e1 = None

    def __enter__(self):
        if TIMER is not None:
            self.e1 = TIMER.begin(self.name)

    def __exit__(self, *exc):
        if self.e1 is not None:
            self.e1.record(torch.cuda.current_stream())
        return False


def profile_kernels(on) -> None:
    """Bracket every kernel the library launches with CUDA events: True / 1 = eager launches only, 2 = launches under
    stream capture too (external event nodes; read them with profile_graph_report after a replay)."""
    _native.lib().tt_profile_enable(int(on))


def profile_null_span() -> None:
    """An empty kernel inside a span ("null_kernel" in the reports): the overhead of the event pair of a span."""
    _native.check(_native.lib().tt_profile_null_span(_stream()), "profile_null_span")


def profile_graph_report(clear: bool = False) -> dict:
    """{kernel name: (total ms, launches)} of the LAST replay of the graphs captured while profile_kernels(2) was on."""
    import ctypes

    buf = ctypes.create_string_buffer(1 << 16)
    _native.check(_native.lib().tt_profile_report_graph(buf, len(buf), 1 if clear else 0), "profile_graph_report")
    out = {}
    for line in buf.value.decode().splitlines():
        name, ms, n = line.rsplit(" ", 2)
        out[name] = (float(ms), int(n))
    return out


def profile_report() -> dict:
    """{kernel name: (total ms, launches)} since the last report; synchronises the device."""
    import ctypes

    buf = ctypes.create_string_buffer(1 << 16)
    _native.check(_native.lib().tt_profile_report(buf, len(buf)), "profile_report")
    out = {}
    for line in buf.value.decode().splitlines():
        name, ms, n = line.rsplit(" ", 2)
        out[name] = (float(ms), int(n))
    return out


class sm_reserve:
    """Context manager: the persistent kernels enqueued inside size their grids for `reserve` fewer SMs (left to a
    collective running beside them)."""

    def __init__(self, reserve: int):
        self.reserve = int(reserve)

    def __enter__(self):
        if self.reserve > 0:
            sms = torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count
            self.prev = _native.lib().tt_set_sm_limit(max(sms - self.reserve, 1))
        return self

    def __exit__(self, *exc):
        if self.reserve > 0:
            _native.lib().tt_set_sm_limit(self.prev)
        return False


def launch_count() -> int:
    """Kernels launched by libtt_b200.so in this process so far."""
    return int(_native.lib().tt_launch_count())


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


def attach_shadow(t: torch.Tensor, t16: Optional[torch.Tensor], colsum: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Attach the bf16 operand copy (and optionally the fp32 column sums) of `t` as attributes, stamped with the
    tensor's (data_ptr, version): a consumer only trusts them while `t` is still the very tensor they describe
    (autograd may accumulate another gradient into it in place, which bumps the version)."""
    if t16 is not None:
        t._tt_bf16 = t16
    if colsum is not None:
        t._tt_colsum = colsum
    t._tt_stamp = (t.data_ptr(), t._version)
    return t


def shadow_of(t: torch.Tensor, name: str = "_tt_bf16") -> Optional[torch.Tensor]:
    """The attribute attached by attach_shadow, or None when it is absent or stale."""
    s = getattr(t, name, None)
    if s is None or getattr(t, "_tt_stamp", None) != (t.data_ptr(), t._version):
        return None
    return s


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
         accumulate=False, split_k=0, colsum: Optional[torch.Tensor] = None) -> None:
    """C = alpha*A*B^T (+bias)(relu)(mask); A,B bf16 2-D views (last stride 1); see tt_gemm_bf16."""
    L = _native.lib()
    _native.check(
        L.tt_gemm_bf16(A.data_ptr(), A.stride(0), int(a_mn), B.data_ptr(), B.stride(0), int(b_mn), M, N, K,
                       _ptr(bias), int(relu), _ptr(relu_mask), relu_mask.stride(0) if relu_mask is not None else 0,
                       alpha, _ptr(out32), out32.stride(0) if out32 is not None else 0,
                       _ptr(out16), out16.stride(0) if out16 is not None else 0,
                       int(accumulate), split_k, _ptr(colsum), _stream()),
        "gemm_bf16",
    )


def colsum(src: torch.Tensor, cols: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[c] (+)= sum_r src[r, c]; a caller-provided `out` must be initialised."""
    if out is None:
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


def check_ids(device=None, what="embedding lookup") -> None:
    """Raise IndexError if a lookup since the last check saw an id outside its table (synchronises the device)."""
    devs = list(_oob_flags) if device is None else [device]
    for dev in devs:
        f = _oob_flags.get(dev)
        if f is not None and int(f.item()) != 0:
            f.zero_()
            raise IndexError(f"{what}: index out of range in embedding lookup")


def _maybe_check_ids(device, what):
    global _check_ids_calls
    if _CHECK_IDS == "0":
        return
    _check_ids_calls += 1
    if _CHECK_IDS == "1" or _check_ids_calls % _CHECK_IDS_EVERY == 1:
        if torch.cuda.is_current_stream_capturing():
            return
        check_ids(device, what)


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


def gather_rows_new(table: torch.Tensor, ids: torch.Tensor) -> torch.Tensor:
    """fp32 table[ids] as a fresh [n, dim] tensor."""
    _need_cuda(table, ids)
    ids = _i64c(ids).reshape(-1)
    out = torch.empty((ids.numel(), table.shape[1]), dtype=torch.float32, device=table.device)
    gather_rows(_f32c(table), ids, out)
    return out


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
        self._force = False

    def invalidate(self):
        """Re-pack every weight on its next use (the buffers are kept and overwritten in place)."""
        for k, (ver, out) in list(self._cache.items()):
            self._cache[k] = (None, out)

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
                if hit is not None and hit[1].shape == (rows, _r8(cols)) and hit[1].device == param.device:
                    out = hit[1]
                    cast_rows_bf16(src, out=out, cols=cols)
                else:
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


def _zero_arena(sizes, device):
    """One zero-filled fp32 buffer carved into 16-byte aligned pieces (one memset instead of one per gradient)."""
    offs, total = [], 0
    for n in sizes:
        offs.append(total)
        total += (n + 3) // 4 * 4
    flat = torch.zeros(total, dtype=torch.float32, device=device)
    return [flat[o:o + n] for o, n in zip(offs, sizes)]


def _mlp_backward(dFe16, db1, D, feats16, F, H16, w0_16, w1_16, need_dfeats=False, bufs=None):
    """Gradients of the feature MLP given dFe (bf16 view [rows, >=D]) and its fp32 column sums db1."""
    rows = feats16.shape[0]
    hid = H16.shape[1]
    dev = feats16.device
    if bufs is None:
        bufs = _zero_arena([D * hid, hid, hid * F], dev)
    dW1, db0, dW0 = bufs[0].view(D, hid), bufs[1], bufs[2].view(hid, F)
    gemm(dFe16, H16, D, hid, rows, a_mn=True, b_mn=True, out32=dW1, accumulate=True)
    dH16 = torch.empty((rows, hid), dtype=_BF16, device=dev)
    gemm(dFe16, w1_16, rows, hid, D, b_mn=True, relu_mask=H16, out16=dH16, colsum=db0)  # db0 from the fp32 accumulators
    gemm(dH16, feats16, hid, F, rows, a_mn=True, b_mn=True, out32=dW0, accumulate=True)
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
    def forward(ctx, ids, feats, extra, table, w0, b0, w1, b1, wt, bt, packed: PackedWeights, tag: str, row_exchange=None):
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
        dense = (D8 == D) and (_r8(E) == E)
        X16 = (torch.empty if dense else torch.zeros)((B, KT), dtype=_BF16, device=feats.device)
        emb = torch.empty((B, DI), dtype=torch.float32, device=feats.device)
        emb16 = torch.empty((B, _r8(DI)), dtype=_BF16, device=feats.device)
        if tower_fused_supported(F, D, DI, w0.shape[0], E):
            feats16 = torch.empty((B, F), dtype=_BF16, device=feats.device)
            H16 = torch.empty((B, w0.shape[0]), dtype=_BF16, device=feats.device)
            tower_forward_fused([dict(ids=ids, feats=feats, table=_f32c(table), w0_16=w0_16, w1_16=w1_16, wt_16=wt_16,
                                      b0=_f32c(b0), b1=_f32c(b1), bt=_f32c(bt), feats16=feats16, H16=H16, X16=X16,
                                      emb=emb, emb16=emb16, B=B, F=F, D=D, DI=DI, hid=w0.shape[0])])
        else:
            feats16 = cast_rows_bf16(feats)
            gather_rows(table, ids, X16, 0)
            H16 = _mlp_forward(feats16, F, w0_16, _f32c(b0), w1_16, _f32c(b1), D, out16=X16, out16_col=D8)
            if E:
                cast_rows_bf16(_f32c(extra), out=X16, col_offset=2 * D8)
            gemm(X16, wt_16, B, DI, KT, bias=_f32c(bt), out32=emb, out16=emb16)
        ctx.save_for_backward(ids, feats16, H16, X16, w0_16, w1_16, wt_16)
        ctx.dims = (B, F, D, DI, E, D8, KT, table.shape[0])
        ctx.need_dfeats = feats.requires_grad
        ctx.has_extra = extra is not None
        ctx.row_exchange = row_exchange
        return attach_shadow(emb, emb16)

    @staticmethod
    def backward(ctx, demb):
        ids, feats16, H16, X16, w0_16, w1_16, wt_16 = ctx.saved_tensors
        B, F, D, DI, E, D8, KT, table_rows = ctx.dims
        dev = demb.device
        demb16 = shadow_of(demb)
        if demb16 is None:
            demb16 = cast_rows_bf16(_f32c(demb))
        hid = H16.shape[1]
        arena = _zero_arena([DI * KT, KT, DI, D * hid, hid, hid * F], dev)  # every accumulated gradient, one memset
        # tower Linear
        dWt_p = arena[0].view(DI, KT)
        gemm(demb16, X16, DI, KT, B, a_mn=True, b_mn=True, out32=dWt_p, accumulate=True)
        dbt = colsum(_f32c(demb), DI, out=arena[2])  # fp32 source: item-side bias gradients are analytically zero sums
        dX16 = torch.empty((B, KT), dtype=_BF16, device=dev)
        dXsum = arena[1]
        gemm(demb16, wt_16, B, KT, DI, b_mn=True, out16=dX16, colsum=dXsum)
        if D8 == D and _r8(E) == E:
            dWt = dWt_p
        else:
            parts = [dWt_p[:, :D], dWt_p[:, D8:D8 + D]] + ([dWt_p[:, 2 * D8:2 * D8 + E]] if E else [])
            dWt = torch.cat(parts, dim=1)
        # id embedding (dense gradient, duplicates accumulate)
        if ctx.row_exchange is not None:
            # data parallel: all-gather the touched (ids, row gradients) of every rank and apply all of them,
            # so the dense [hash, D] gradient is already the global sum (no 4*hash*D-byte all-reduce)
            ids_all, rows_all = ctx.row_exchange(ids, dX16[:, :D8].contiguous())
            dtable = scatter_add_rows(rows_all, ids_all, D, table_rows, col_offset=0)
        else:
            dtable = scatter_add_rows(dX16, ids, D, table_rows, col_offset=0)
        # feature MLP
        dW0, db0, dW1, db1, dfeats = _mlp_backward(dX16[:, D8:], dXsum[D8:D8 + D], D, feats16, F, H16, w0_16, w1_16,
                                                   ctx.need_dfeats, bufs=arena[3:6])
        dextra = dX16[:, 2 * D8:2 * D8 + E].float() if ctx.has_extra else None
        return None, dfeats, dextra, dtable, dW0, db0, dW1, db1, dWt, dbt, None, None, None


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
        dout = _f32c(dout)
        d16 = cast_rows_bf16(dout)
        dW0, db0, dW1, db1, dfeats = _mlp_backward(d16, colsum(dout, D), D, feats16, F, H16, w0_16, w1_16,
                                                   ctx.need_dfeats)
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
        return attach_shadow(y, y16)

    @staticmethod
    def backward(ctx, dy):
        x16, w16 = ctx.saved_tensors
        B, K, N = ctx.dims
        dy16 = shadow_of(dy)
        if dy16 is None:
            dy16 = cast_rows_bf16(_f32c(dy))
        dW = torch.zeros((N, K), dtype=torch.float32, device=dy.device)
        gemm(dy16, x16, N, K, B, a_mn=True, b_mn=True, out32=dW, accumulate=True)
        db = colsum(_f32c(dy), N) if ctx.has_bias else None
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
    _attach_pending_zero()
    with _span("inbatch_ce_fwd"):
        _native.check(
            _native.lib().tt_inbatch_ce_fwd(U16.data_ptr(), U16.stride(0), V16.data_ptr(), V16.stride(0), B, N, d,
                                            target_offset, ce.data_ptr(), lse.data_ptr(), ws.data_ptr(), ws.numel(),
                                            _stream()),
            "inbatch_ce_fwd",
        )
    return ce, lse


def inbatch_ce_backward_raw(U16, V16, B, N, d, target_offset, lse, g, want_bf16=True, colsums=None, g_scale=None,
                            g_scale2=None, want=("dU", "dV")):
    """colsums: optional zero-initialised fp32 [2, d] receiving the column sums of dU and dV (d <= 128).
    g_scale, g_scale2: optional one-element fp32 device tensors multiplied into g inside the kernels.
    want: which gradients to compute (each is one pass over the score tiles); the others come back as None."""
    dev = U16.device
    wu, wv = "dU" in want, "dV" in want
    dU = torch.empty((B, d), dtype=torch.float32, device=dev) if wu else None
    dV = torch.empty((N, d), dtype=torch.float32, device=dev) if wv else None
    dU16 = torch.empty((B, _r8(d)), dtype=_BF16, device=dev) if (want_bf16 and wu) else None
    dV16 = torch.empty((N, _r8(d)), dtype=_BF16, device=dev) if (want_bf16 and wv) else None
    ws = _ce_workspace(B, N, d, dev)
    with _span("inbatch_ce_bwd"):
        _native.check(
            _native.lib().tt_inbatch_ce_bwd_scaled(
                U16.data_ptr(), U16.stride(0), V16.data_ptr(), V16.stride(0), B, N, d, target_offset,
                lse.data_ptr(), g.data_ptr(), _ptr(g_scale), _ptr(g_scale2), _ptr(dU), d if wu else 0, _ptr(dU16),
                dU16.stride(0) if dU16 is not None else 0, _ptr(dV), d if wv else 0, _ptr(dV16),
                dV16.stride(0) if dV16 is not None else 0,
                None if (colsums is None or not wu) else colsums.data_ptr(),
                None if (colsums is None or not wv) else colsums.data_ptr() + 4 * d,
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
        U16 = shadow_of(U)
        V16 = shadow_of(V)
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
        cs = torch.zeros((2, d), dtype=torch.float32, device=g.device) if d <= 128 else None
        dU, dV, dU16, dV16 = inbatch_ce_backward_raw(U16, V16, B, N, d, off, lse, _f32c(g), colsums=cs)
        # fp32 column sums = bias gradients of the tower Linears, saves two reduction launches
        attach_shadow(dU, dU16, None if cs is None else cs[0])
        attach_shadow(dV, dV16, None if cs is None else cs[1])
        return dU, dV, None


def inbatch_cross_entropy(U: torch.Tensor, V: torch.Tensor, target_offset: int = 0) -> torch.Tensor:
    return InBatchCEFunction.apply(U, V, target_offset)


class InBatchWeightedLossFunction(torch.autograd.Function):
    """loss = mean(ce(U V^T, diagonal) * clamp(labels @ w, 1e-6) / max(.)): the whole of compute_training_loss with the
    identity hook (reference :279-347) in two launches forward (scores+softmax statistics, merge+weights+mean) and
    three backward (dU pass, dV pass, merge); the incoming d loss is applied inside the backward kernels."""

    @staticmethod
    def forward(ctx, U, V, labels, weights):
        _need_cuda(U, V, labels, weights)
        B, d = U.shape
        N = V.shape[0]
        if V.shape[1] != d:
            raise RuntimeError(f"user/item embedding dims differ: {d} vs {V.shape[1]}")
        U16 = shadow_of(U)
        V16 = shadow_of(V)
        if U16 is None:
            U16 = cast_rows_bf16(_f32c(U))
        if V16 is None:
            V16 = cast_rows_bf16(_f32c(V))
        labels, weights = _f32c(labels), _f32c(weights)
        dev = U16.device
        out = torch.empty(3 * B + 1, dtype=torch.float32, device=dev)  # ce | lse | g | g_norm (saved for backward)
        ce, lse, g, g_norm = out[:B], out[B:2 * B], out[2 * B:3 * B], out[3 * B:]
        loss = torch.empty((), dtype=torch.float32, device=dev)  # its own storage: the caller may modify it in place
        ws = _ce_workspace(B, N, d, dev)
        _attach_pending_zero()
        with _span("inbatch_ce_fwd"):
            _native.check(
                _native.lib().tt_inbatch_ce_loss_fwd(
                    U16.data_ptr(), U16.stride(0), V16.data_ptr(), V16.stride(0), B, N, d, 0, labels.data_ptr(),
                    labels.stride(0), weights.data_ptr(), labels.shape[1], ce.data_ptr(), lse.data_ptr(),
                    loss.data_ptr(), g.data_ptr(), g_norm.data_ptr(), ws.data_ptr(), ws.numel(), _stream()),
                "inbatch_ce_loss_fwd",
            )
        ctx.save_for_backward(U16, V16, lse, g, g_norm)
        ctx.dims = (B, N, d)
        return loss

    @staticmethod
    def backward(ctx, dloss):
        U16, V16, lse, g, g_norm = ctx.saved_tensors
        B, N, d = ctx.dims
        cs = torch.zeros((2, d), dtype=torch.float32, device=g.device) if d <= 128 else None
        dU, dV, dU16, dV16 = inbatch_ce_backward_raw(U16, V16, B, N, d, 0, lse, g, colsums=cs,
                                                     g_scale=_f32c(dloss).reshape(1), g_scale2=g_norm)
        attach_shadow(dU, dU16, None if cs is None else cs[0])
        attach_shadow(dV, dV16, None if cs is None else cs[1])
        return dU, dV, None, None


class WeightedLossFunction(torch.autograd.Function):
    """loss = mean(ce * clamp(labels @ w, 1e-6) / max(.)) in one launch (reference :322-343, identity hook)."""

    @staticmethod
    def forward(ctx, ce, labels, weights):
        _need_cuda(ce, labels, weights)
        ce, labels, weights = _f32c(ce), _f32c(labels), _f32c(weights)
        B, T = labels.shape
        loss = torch.empty((), dtype=torch.float32, device=ce.device)
        g = torch.empty(B, dtype=torch.float32, device=ce.device)
        _native.check(
            _native.lib().tt_weighted_loss(ce.data_ptr(), labels.data_ptr(), labels.stride(0), weights.data_ptr(), B, T,
                                           loss.data_ptr(), g.data_ptr(), _stream()),
            "weighted_loss",
        )
        ctx.save_for_backward(g)
        return loss

    @staticmethod
    def backward(ctx, dloss):
        (g,) = ctx.saved_tensors
        return g * dloss, None, None


# --------------------------------------------------------------------------------------------
# brute-force MIPS   (reference src/baseline_mips_module.py:57-72)
# --------------------------------------------------------------------------------------------
_mips_ws = {}


def mips_topk(query: torch.Tensor, corpus: torch.Tensor, corpus16: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """(indices int64 [Q,k], scores fp32 [Q,k]) = top-k of query @ corpus^T, scores descending."""
    _need_cuda(query, corpus, corpus16)
    q32 = _f32c(query.detach())
    c32 = _f32c(corpus)
    nq, d = q32.shape
    nc = c32.shape[0]
    q16 = shadow_of(query)
    if q16 is None:
        q16 = cast_rows_bf16(q32)
    L = _native.lib()
    nbytes = int(L.tt_mips_workspace_bytes(nq, nc, d, k))
    key = (str(q32.device), nbytes)
    ws = _mips_ws.get(key)
    if ws is None:
        _mips_ws.clear()
        ws = torch.empty(nbytes, dtype=torch.uint8, device=q32.device)
        _mips_ws[key] = ws
    idx = torch.empty((nq, k), dtype=torch.int64, device=q32.device)
    scores = torch.empty((nq, k), dtype=torch.float32, device=q32.device)
    with _span("mips_topk"):
        _native.check(
            L.tt_mips_topk(q16.data_ptr(), q16.stride(0), corpus16.data_ptr(), corpus16.stride(0), q32.data_ptr(),
                           q32.stride(0), c32.data_ptr(), c32.stride(0), nq, nc, d, k, idx.data_ptr(),
                           scores.data_ptr(), ws.data_ptr(), ws.numel(), _stream()),
            "mips_topk",
        )
    return idx, scores


# --------------------------------------------------------------------------------------------
# history encoder   (reference src/user_history_encoder.py:80-121)
# --------------------------------------------------------------------------------------------
def attn_forward(qkv16: torch.Tensor, nseq: int, H: int, D: int, heads: int, q_rows: int) -> torch.Tensor:
    out = torch.empty((nseq * q_rows, _r8(D)), dtype=_BF16, device=qkv16.device)
    with _span("attn_fwd"):
        _native.check(
            _native.lib().tt_attn_fwd(qkv16.data_ptr(), qkv16.stride(0), nseq, H, D, heads, q_rows, out.data_ptr(),
                                      out.stride(0), _stream()),
            "attn_fwd",
        )
    return out


def attn_backward(qkv16: torch.Tensor, dout16: torch.Tensor, nseq: int, H: int, D: int, heads: int,
                  q_rows: int) -> torch.Tensor:
    dqkv = torch.empty((nseq * H, qkv16.shape[1]), dtype=_BF16, device=qkv16.device)
    with _span("attn_bwd"):
        _native.check(
            _native.lib().tt_attn_bwd(qkv16.data_ptr(), qkv16.stride(0), dout16.data_ptr(), dout16.stride(0), nseq, H, D,
                                      heads, q_rows, dqkv.data_ptr(), dqkv.stride(0), _stream()),
            "attn_bwd",
        )
    return dqkv


class HistoryEncoderFunction(torch.autograd.Function):
    """summary[B, 2D] = [attention output of history row 0 | mean-pooled history].

    Two entry forms:  (ids int64 [B,H], table fp32 [rows, D])  - the embedding lookup is folded into the
    first kernel (TwoTowerWithUserHistoryEncoder), or  (None, x fp32 [B,H,D])  - UserHistoryEncoder.forward.
    Per layer: packed in-projection (tcgen05 GEMM + bias) -> per-head softmax(QK^T)V -> out-projection
    (GEMM + bias); no residual / LayerNorm / FFN, as in the reference.  The last layer only evaluates
    query row 0 of every sequence (the only row the reference consumes, :116).
    """

    @staticmethod
    def forward(ctx, ids, table_or_x, pe, heads: int, packed: PackedWeights, tag: str, *params):
        src = table_or_x
        _need_cuda(src, ids, pe)
        if ids is None:
            B, H, D = src.shape
            table = _f32c(src).reshape(B * H, D)
            ids_ = torch.arange(B * H, dtype=torch.int64, device=src.device).reshape(B, H)
        else:
            table = _f32c(src)
            ids_ = _i64c(ids)
            B, H = ids_.shape
            D = table.shape[1]
        L = len(params) // 4
        dev = table.device
        lib = _native.lib()
        D8, Q8 = _r8(D), _r8(3 * D)
        x16 = torch.empty((B * H, D8), dtype=_BF16, device=dev)
        summary = torch.empty((B, 2 * D), dtype=torch.float32, device=dev)
        flag = _oob_flag(dev)
        pe_ = None if pe is None else _f32c(pe)
        _native.check(
            lib.tt_history_gather_pool(table.data_ptr(), table.shape[0], D, ids_.data_ptr(), B, H, _ptr(pe_),
                                       x16.data_ptr(), x16.stride(0), summary.data_ptr() + 4 * D, summary.stride(0),
                                       flag.data_ptr(), _stream()),
            "history_gather_pool",
        )
        _maybe_check_ids(dev, "history lookup")
        saved = []
        recent = summary[:, :D]
        hd = D // heads if heads else 0
        # last layer: only query row 0 is consumed (reference :116) -> attention against the RAW layer input, no key / value
        # projection of the H rows (csrc/history_last.cu); other shapes take the generic kernels with q_rows = 1
        fast_last = (_HISTORY_LAST_FAST and L >= 1 and D8 == D and heads >= 1 and D % heads == 0
                     and bool(lib.tt_history_last_supported(H, D, heads)))
        for l in range(L):
            in_w, in_b, out_w, out_b = params[4 * l: 4 * l + 4]
            last = l == L - 1
            q_rows = 1 if last else H
            in_w16 = packed.get(f"{tag}.{l}.in", in_w)
            out_w16 = packed.get(f"{tag}.{l}.out", out_w)
            if last and fast_last:
                in_b32 = _f32c(in_b)
                c = float(hd) ** -0.5
                x0 = x16.view(B, H, D8)[:, 0]  # [B, D] view, row pitch H * D8
                q0_16 = torch.empty((B, D8), dtype=_BF16, device=dev)
                gemm(x0, in_w16[:D], B, D, D, bias=in_b32[:D], out16=q0_16)
                qt32 = torch.empty((B, heads * D), dtype=torch.float32, device=dev)
                gemm_batched([dict(A=q0_16[:, h * hd:(h + 1) * hd], B=in_w16[D + h * hd:D + (h + 1) * hd], M=B, N=D, K=hd,
                                   b_mn=True, alpha=c, out32=qt32[:, h * D:(h + 1) * D]) for h in range(heads)])
                z16 = torch.empty((B, heads * D), dtype=_BF16, device=dev)
                p32 = torch.empty((B, heads, H), dtype=torch.float32, device=dev)
                with _span("hist_last_fwd"):
                    _native.check(lib.tt_history_last_fwd(x16.data_ptr(), x16.stride(0), qt32.data_ptr(), B, H, D, heads,
                                                          z16.data_ptr(), p32.data_ptr(), _stream()), "history_last_fwd")
                o16 = torch.empty((B, D8), dtype=_BF16, device=dev)
                gemm_batched([dict(A=z16[:, h * D:(h + 1) * D], B=in_w16[2 * D + h * hd:2 * D + (h + 1) * hd], M=B, N=hd, K=D,
                                   bias=in_b32[2 * D + h * hd:2 * D + (h + 1) * hd], out16=o16[:, h * hd:(h + 1) * hd])
                              for h in range(heads)])
                gemm(o16, out_w16, B, D, D, bias=_f32c(out_b), out32=recent)
                saved += [x16, q0_16, qt32, z16, p32, o16, in_w16, out_w16]
                continue
            qkv16 = torch.empty((B * H, Q8), dtype=_BF16, device=dev)
            gemm(x16, in_w16, B * H, 3 * D, D, bias=_f32c(in_b), out16=qkv16)
            o16 = attn_forward(qkv16, B, H, D, heads, q_rows)
            saved += [x16, qkv16, o16, in_w16, out_w16]
            if last:
                gemm(o16, out_w16, B, D, D, bias=_f32c(out_b), out32=recent)
            else:
                y16 = torch.empty((B * H, D8), dtype=_BF16, device=dev)
                gemm(o16, out_w16, B * H, D, D, bias=_f32c(out_b), out16=y16)
                x16 = y16
        if L == 0:  # no attention layer: the "most recent" half is row 0 of the (position-encoded) input itself
            recent.copy_(x16.view(B, H, D8)[:, 0, :D])
        ctx.save_for_backward(ids_, *saved)
        ctx.dims = (B, H, D, heads, L, table.shape[0], ids is None, tuple(src.shape))
        ctx.fast_last = fast_last
        return summary

    @staticmethod
    def backward(ctx, dsummary):
        ids_, *saved = ctx.saved_tensors
        B, H, D, heads, L, table_rows, dense_x, src_shape = ctx.dims
        dev = dsummary.device
        lib = _native.lib()
        D8 = _r8(D)
        ds = _f32c(dsummary)
        dy16 = cast_rows_bf16(ds, cols=D)  # gradient of the last layer's row-0 output, [B, D8]
        dy_sum = colsum(ds[:, :D], D)      # its fp32 column sums (= last out-projection bias gradient)
        dmean = ds[:, D:]
        grads = [None] * (4 * L)
        hd = D // heads if heads else 0
        if L == 0:  # gradient of row 0 only
            full = torch.zeros((B, H, D8), dtype=_BF16, device=dev)
            full[:, 0] = dy16
            dy16 = full.view(B * H, D8)
        for l in range(L - 1, -1, -1):
            last = l == L - 1
            if last and ctx.fast_last:
                x16, q0_16, qt32, z16, p32, o16, in_w16, out_w16 = saved[5 * l: 5 * l + 8]
                c = float(hd) ** -0.5
                d_out_w = torch.zeros((D, D), dtype=torch.float32, device=dev)
                gemm(dy16, o16, D, D, B, a_mn=True, b_mn=True, out32=d_out_w, accumulate=True)
                d_out_b = dy_sum
                d_in_w = torch.zeros((3 * D, D), dtype=torch.float32, device=dev)
                d_in_b = torch.zeros(3 * D, dtype=torch.float32, device=dev)
                do16 = torch.empty((B, D8), dtype=_BF16, device=dev)
                gemm(dy16, out_w16, B, D, D, b_mn=True, out16=do16, colsum=d_in_b[2 * D:])  # + column sums = d bv
                dz32 = torch.empty((B, heads * D), dtype=torch.float32, device=dev)
                gemm_batched([dict(A=do16[:, h * hd:(h + 1) * hd], B=in_w16[2 * D + h * hd:2 * D + (h + 1) * hd], M=B, N=D, K=hd,
                                   b_mn=True, out32=dz32[:, h * D:(h + 1) * D]) for h in range(heads)])
                gemm_batched([dict(A=do16[:, h * hd:(h + 1) * hd], B=z16[:, h * D:(h + 1) * D], M=hd, N=D, K=B, a_mn=True, b_mn=True,
                                   out32=d_in_w[2 * D + h * hd:2 * D + (h + 1) * hd], accumulate=True) for h in range(heads)])
                ds32 = torch.empty((B, heads, H), dtype=torch.float32, device=dev)
                dqt16 = torch.empty((B, heads * D), dtype=_BF16, device=dev)
                with _span("hist_last_bwd"):
                    _native.check(lib.tt_history_last_bwd1(x16.data_ptr(), x16.stride(0), dz32.data_ptr(), p32.data_ptr(), B, H, D,
                                                           heads, ds32.data_ptr(), dqt16.data_ptr(), _stream()), "history_last_bwd1")
                dq0_16 = torch.empty((B, D8), dtype=_BF16, device=dev)
                gemm_batched([dict(A=dqt16[:, h * D:(h + 1) * D], B=in_w16[D + h * hd:D + (h + 1) * hd], M=B, N=hd, K=D, alpha=c,
                                   out16=dq0_16[:, h * hd:(h + 1) * hd], colsum=d_in_b[h * hd:(h + 1) * hd]) for h in range(heads)])
                gemm_batched([dict(A=q0_16[:, h * hd:(h + 1) * hd], B=dqt16[:, h * D:(h + 1) * D], M=hd, N=D, K=B, a_mn=True, b_mn=True,
                                   alpha=c, out32=d_in_w[D + h * hd:D + (h + 1) * hd], accumulate=True) for h in range(heads)])
                x0 = x16.view(B, H, D8)[:, 0]
                gemm(dq0_16, x0, D, D, B, a_mn=True, b_mn=True, out32=d_in_w[:D], accumulate=True)
                extra32 = torch.empty((B, D), dtype=torch.float32, device=dev)
                gemm(dq0_16, in_w16[:D], B, D, D, b_mn=True, out32=extra32)  # d q0 Wq: lands on row 0 of every sequence
                dx16 = torch.empty((B * H, D8), dtype=_BF16, device=dev)
                dy_sum = torch.zeros(D, dtype=torch.float32, device=dev)  # bias gradient of the layer below
                with _span("hist_last_bwd"):
                    _native.check(lib.tt_history_last_bwd2(dz32.data_ptr(), qt32.data_ptr(), p32.data_ptr(), ds32.data_ptr(),
                                                           extra32.data_ptr(), B, H, D, heads, dx16.data_ptr(), dx16.stride(0),
                                                           dy_sum.data_ptr(), _stream()), "history_last_bwd2")
                grads[4 * l: 4 * l + 4] = [d_in_w, d_in_b, d_out_w, d_out_b]
                dy16 = dx16
                continue
            x16, qkv16, o16, in_w16, out_w16 = saved[5 * l: 5 * l + 5]
            q_rows = 1 if last else H
            rows = B * q_rows
            d_out_w = torch.zeros((D, D), dtype=torch.float32, device=dev)
            gemm(dy16, o16, D, D, rows, a_mn=True, b_mn=True, out32=d_out_w, accumulate=True)
            d_out_b = dy_sum
            do16 = torch.empty((rows, D8), dtype=_BF16, device=dev)
            gemm(dy16, out_w16, rows, D, D, b_mn=True, out16=do16)
            dqkv16 = attn_backward(qkv16, do16, B, H, D, heads, q_rows)
            d_in_w = torch.zeros((3 * D, D), dtype=torch.float32, device=dev)
            gemm(dqkv16, x16, 3 * D, D, B * H, a_mn=True, b_mn=True, out32=d_in_w, accumulate=True)
            d_in_b = colsum(dqkv16, 3 * D)
            dx16 = torch.empty((B * H, D8), dtype=_BF16, device=dev)
            dy_sum = torch.zeros(D, dtype=torch.float32, device=dev)  # bias gradient of the layer below
            gemm(dqkv16, in_w16, B * H, D, 3 * D, b_mn=True, out16=dx16, colsum=dy_sum)
            grads[4 * l: 4 * l + 4] = [d_in_w, d_in_b, d_out_w, d_out_b]
            dy16 = dx16
        dtable = torch.zeros((table_rows, D), dtype=torch.float32, device=dev)
        _native.check(
            lib.tt_history_scatter_grad(dy16.data_ptr(), dy16.stride(0), dmean.data_ptr(), dmean.stride(0),
                                        ids_.data_ptr(), B, H, D, dtable.data_ptr(), table_rows, _stream()),
            "history_scatter_grad",
        )
        if dense_x:
            dtable = dtable.reshape(src_shape)
        return (None, dtable, None, None, None, None, *grads)


# --------------------------------------------------------------------------------------------
# batched launches: the same step of several towers in ONE kernel launch
# --------------------------------------------------------------------------------------------
def gemm_batched(problems) -> None:
    """problems: list of dicts with the keyword arguments of `gemm` plus A, B, M, N, K."""
    n = len(problems)
    arr = (_native.GemmProblem * n)()
    for i, p in enumerate(problems):
        g = arr[i]
        A, B = p["A"], p["B"]
        g.A, g.lda, g.B, g.ldb = A.data_ptr(), A.stride(0), B.data_ptr(), B.stride(0)
        g.M, g.N, g.K = p["M"], p["N"], p["K"]
        g.a_mn_major, g.b_mn_major = int(p.get("a_mn", False)), int(p.get("b_mn", False))
        bias, mask = p.get("bias"), p.get("relu_mask")
        g.bias = _ptr(bias)
        g.relu = int(p.get("relu", False))
        g.relu_mask_bf16 = _ptr(mask)
        g.ld_mask = mask.stride(0) if mask is not None else 0
        o32, o16 = p.get("out32"), p.get("out16")
        g.c_f32, g.ldc_f32 = _ptr(o32), (o32.stride(0) if o32 is not None else 0)
        g.c_bf16, g.ldc_bf16 = _ptr(o16), (o16.stride(0) if o16 is not None else 0)
        g.colsum_f32 = _ptr(p.get("colsum"))
        g.alpha = p.get("alpha", 1.0)
        g.accumulate, g.split_k = int(p.get("accumulate", False)), p.get("split_k", 0)
    _native.check(_native.lib().tt_gemm_bf16_batched(arr, n, _stream()), "gemm_bf16_batched")


def cast_batched(items) -> None:
    """items: list of (src fp32 2-D, src_col0, cols, out bf16 2-D, col_offset, dst_cols); one launch per 16 items."""
    for i0 in range(0, len(items), 16):
        chunk = items[i0:i0 + 16]
        arr = (_native.CastProblem * len(chunk))()
        for c, (src, src_col0, cols, out, col_offset, dst_cols) in zip(arr, chunk):
            c.src, c.rows, c.cols, c.ld_src = src.data_ptr() + 4 * src_col0, src.shape[0], cols, src.stride(0)
            c.dst_bf16, c.ld_dst, c.dst_cols = out.data_ptr() + 2 * col_offset, out.stride(0), dst_cols
        _native.check(_native.lib().tt_cast_rows_bf16_batched(arr, len(chunk), _stream()), "cast_rows_bf16_batched")


def gather_batched(items) -> None:
    """items: list of (table fp32, ids int64, out bf16, col_offset)."""
    arr = (_native.GatherProblem * len(items))()
    for g, (table, ids, out, col_offset) in zip(arr, items):
        g.table, g.table_rows, g.dim = table.data_ptr(), table.shape[0], table.shape[1]
        g.ids, g.n = ids.data_ptr(), ids.numel()
        g.dst_bf16, g.ld_dst = out.data_ptr() + 2 * col_offset, out.stride(0)
    flag = _oob_flag(items[0][0].device)
    _native.check(_native.lib().tt_gather_rows_bf16_batched(arr, len(items), flag.data_ptr(), _stream()),
                  "gather_rows_bf16_batched")
    _maybe_check_ids(items[0][0].device, "gather_rows")


def tower_fused_supported(F: int, D: int, DI: int, hid: int, E: int) -> bool:
    """Shapes the single-launch tower kernel takes (csrc/tower.cu); others use the per-layer launches."""
    if os.environ.get("TT_B200_FUSED_TOWER", "1") != "1" or E != 0:
        return False
    return bool(_native.lib().tt_tower_fwd_supported(F, D, DI, hid))


def tower_forward_fused(towers) -> None:
    """towers: dicts with ids, feats (fp32), table, w0_16, w1_16, wt_16, b0, b1, bt and the outputs feats16, H16,
    X16, emb, emb16: the whole forward of every tower in one launch (tt_tower_fwd)."""
    arr = (_native.TowerProblem * len(towers))()
    for t, d in zip(arr, towers):
        t.ids, t.table, t.table_rows = d["ids"].data_ptr(), d["table"].data_ptr(), d["table"].shape[0]
        t.feats, t.ld_feats = d["feats"].data_ptr(), d["feats"].stride(0)
        t.w0_bf16, t.ldw0, t.b0 = d["w0_16"].data_ptr(), d["w0_16"].stride(0), d["b0"].data_ptr()
        t.w1_bf16, t.ldw1, t.b1 = d["w1_16"].data_ptr(), d["w1_16"].stride(0), d["b1"].data_ptr()
        t.wt_bf16, t.ldwt, t.bt = d["wt_16"].data_ptr(), d["wt_16"].stride(0), d["bt"].data_ptr()
        t.feats_bf16, t.ld_feats16 = d["feats16"].data_ptr(), d["feats16"].stride(0)
        t.h_bf16, t.ldh = d["H16"].data_ptr(), d["H16"].stride(0)
        t.x_bf16, t.ldx = d["X16"].data_ptr(), d["X16"].stride(0)
        t.emb_f32, t.ld_emb = d["emb"].data_ptr(), d["emb"].stride(0)
        t.emb_bf16, t.ld_emb16 = d["emb16"].data_ptr(), d["emb16"].stride(0)
        t.rows, t.F, t.D, t.DI, t.hidden = d["B"], d["F"], d["D"], d["DI"], d["hid"]
    flag = _oob_flag(towers[0]["feats"].device)
    _native.check(_native.lib().tt_tower_fwd(arr, len(towers), flag.data_ptr(), _stream()), "tower_fwd")
    _maybe_check_ids(towers[0]["feats"].device, "gather_rows")


_HISTORY_LAST_FAST = os.environ.get("TT_B200_HISTORY_LAST_FAST", "1") == "1"  # row-0-only last encoder layer (history_last.cu)
_FUSED_TOWER_BWD = os.environ.get("TT_B200_FUSED_TOWER_BWD", "0") == "1"  # candidate, see csrc/tower_bwd.cu


def tower_backward_chain_fused(towers) -> None:
    """dX, dH, the dense table gradients and the db1 / db0 column sums of every tower in one launch
    (tt_tower_bwd_chain); towers: dicts with demb16, ids, wt_16, w1_16, H16, dX16, dH16, dtable_buf, dXsum, db0."""
    arr = (_native.TowerBwdProblem * len(towers))()
    for t, d in zip(arr, towers):
        t.demb_bf16, t.ld_demb = d["demb16"].data_ptr(), d["demb16"].stride(0)
        t.ids, t.table_rows = d["ids"].data_ptr(), d["table_rows"]
        t.wt_bf16, t.ldwt = d["wt_16"].data_ptr(), d["wt_16"].stride(0)
        t.w1_bf16, t.ldw1 = d["w1_16"].data_ptr(), d["w1_16"].stride(0)
        t.h_bf16, t.ldh = d["H16"].data_ptr(), d["H16"].stride(0)
        t.dx_bf16, t.lddx = d["dX16"].data_ptr(), d["dX16"].stride(0)
        t.dh_bf16, t.lddh = d["dH16"].data_ptr(), d["dH16"].stride(0)
        t.table_grad, t.dxsum, t.db0 = d["dtable_buf"].data_ptr(), d["dXsum"].data_ptr(), d["db0"].data_ptr()
        t.rows, t.D, t.DI, t.hidden = d["B"], d["D"], d["DI"], d["hid"]
    _native.check(_native.lib().tt_tower_bwd_chain(arr, len(towers), _stream()), "tower_bwd_chain")


_aux_streams = {}


def _aux_stream(device, idx: int = 0) -> "torch.cuda.Stream":
    st = _aux_streams.get((device, idx))
    if st is None:
        st = torch.cuda.Stream(device=device)
        _aux_streams[(device, idx)] = st
    return st


_PARALLEL_BACKWARD = os.environ.get("TT_B200_PARALLEL_BACKWARD", "1") == "1"
_LATE_ZERO_FILL = os.environ.get("TT_B200_LATE_ZERO_FILL", "1") == "1"
_pending_fills = []  # (device, event) of side-stream zero fills that the current stream has not joined yet


_FUSED_ZERO_FILL = os.environ.get("TT_B200_FUSED_ZERO_FILL", "1") == "1"
_pending_zero = []  # fp32 buffers waiting for a zero fill that the next in-batch CE forward launch will carry


def _attach_pending_zero() -> None:
    """Hand the waiting zero fills (the dense table gradients of the step) to the CE forward launch that follows: an idle
    warp of the scoring kernel streams the zeros out with TMA bulk stores (tt_inbatch_ce_attach_zero_fill)."""
    while _pending_zero:
        a = _pending_zero.pop()
        b = _pending_zero.pop() if _pending_zero else None
        if b is not None and b.device != a.device:
            _pending_zero.append(b)
            b = None
        _native.check(_native.lib().tt_inbatch_ce_attach_zero_fill(
            a.data_ptr(), a.numel() * 4, _ptr(b), 0 if b is None else b.numel() * 4), "inbatch_ce_attach_zero_fill")
        if _pending_zero:  # more than two buffers: the rest is filled by plain memsets
            for t in _pending_zero:
                t.zero_()
            _pending_zero.clear()


def join_pending_fills() -> None:
    """Make the current stream wait for the dense table-gradient zero fills that TowerSetFunction.forward started on
    the side stream.  train_forward calls this after the loss forward (the fills then ran beside the scoring kernels,
    and a CUDA-graph capture of a forward-only call still ends with every stream joined)."""
    while _pending_fills:
        device, ev = _pending_fills.pop()
        torch.cuda.current_stream(device).wait_event(ev)
    while _pending_zero:  # no CE forward launch picked them up (a custom loss): plain memsets
        _pending_zero.pop().zero_()


def _stage_weight(packed: PackedWeights, key, param, casts, segments=None) -> torch.Tensor:
    """bf16 operand copy of `param` from the cache; when stale, its cast is appended to `casts` (run later
    as part of one batched launch) instead of being launched here."""
    ver = (param.data_ptr(), param._version, tuple(param.shape), str(param.device))
    hit = packed._cache.get(key)
    if hit is not None and hit[0] == ver:
        return hit[1]
    rows, cols = param.shape
    src = _f32c(param.detach())
    if segments is None:
        segments = [(0, cols, 0)]
    width = max(d0 + _r8(c) for (_, c, d0) in segments)
    if hit is not None and hit[1].shape == (rows, width) and hit[1].device == param.device:
        out = hit[1]
    else:
        out = torch.zeros((rows, width), dtype=_BF16, device=param.device)
    for (s0, c, d0) in segments:
        casts.append((src, s0, c, out, d0, c))
    packed._cache[key] = (ver, out)
    return out


class TowerSetFunction(torch.autograd.Function):
    """Several towers ([id_emb | MLP(feats) | extra] -> Linear, reference :112-219) advanced in lock step: every
    stage (casts, gathers, each GEMM layer, each gradient GEMM) is ONE launch covering all towers, which halves
    the dependent chain of small launch-latency-bound kernels of a training step and fills the SMs.

    apply(specs, packed, *tensors): specs = [(tag, row_exchange)] per tower; tensors = 10 per tower:
    ids, feats, extra (or None), table, w0, b0, w1, b1, wt, bt.  Returns one fp32 embedding per tower
    (bf16 shadow attached as `_tt_bf16`)."""

    @staticmethod
    def forward(ctx, specs, packed: PackedWeights, *tensors):
        T = len(specs)
        tw, casts, gathers = [], [], []
        with torch.no_grad():
            for t in range(T):
                ids, feats, extra, table, w0, b0, w1, b1, wt, bt = tensors[10 * t: 10 * t + 10]
                tag = specs[t][0]
                _need_cuda(ids, feats, table, w0, wt)
                ids, feats = _i64c(ids), _f32c(feats)
                B, F = feats.shape
                D, DI, hid = table.shape[1], wt.shape[0], w0.shape[0]
                E = 0 if extra is None else extra.shape[1]
                D8 = _r8(D)
                KT = 2 * D8 + _r8(E)
                assert wt.shape[1] == 2 * D + E, "tower weight does not match [id_emb | feat_emb | extra]"
                segs = [(0, D, 0), (D, D, D8)] + ([(2 * D, E, 2 * D8)] if E else [])
                dev = feats.device
                d = dict(ids=ids, B=B, F=F, D=D, DI=DI, E=E, D8=D8, KT=KT, hid=hid, table_rows=table.shape[0],
                         need_dfeats=feats.requires_grad, has_extra=extra is not None, row_exchange=specs[t][1])
                d["w0_16"] = _stage_weight(packed, tag + ".w0", w0, casts)
                d["w1_16"] = _stage_weight(packed, tag + ".w1", w1, casts)
                d["wt_16"] = _stage_weight(packed, tag + ".wt", wt, casts, segments=segs)
                d["feats16"] = torch.empty((B, _r8(F)), dtype=_BF16, device=dev)
                d["fused"] = tower_fused_supported(F, D, DI, hid, E)
                if d["fused"]:
                    d["feats"], d["table"] = feats, _f32c(table)
                else:
                    casts.append((feats, 0, F, d["feats16"], 0, _r8(F)))
                dense = (D8 == D) and (_r8(E) == E)
                d["X16"] = (torch.empty if dense else torch.zeros)((B, KT), dtype=_BF16, device=dev)
                if E:
                    casts.append((_f32c(extra), 0, E, d["X16"], 2 * D8, E))
                if not d["fused"]:
                    gathers.append((_f32c(table), ids, d["X16"], 0))
                d["H16"] = torch.empty((B, hid), dtype=_BF16, device=dev)
                d["emb"] = torch.empty((B, DI), dtype=torch.float32, device=dev)
                d["emb16"] = torch.empty((B, _r8(DI)), dtype=_BF16, device=dev)
                d["b0"], d["b1"], d["bt"] = _f32c(b0), _f32c(b1), _f32c(bt)
                tw.append(d)
            # dense embedding-table gradients (the reference's nn.Embedding(sparse=False) semantics) need a zero
            # fill of hash x D floats per step (2 x 51 MB at the benchmark config).  It runs on a side stream: forked
            # AFTER the tower kernel by default, so that it shares the device with the loss kernels (which move 4 MB)
            # instead of competing with the tower kernel's gathers for HBM, and joined when the backward needs it.
            zero_jobs = [d for t, d in enumerate(tw) if ctx.needs_input_grad[2 + 10 * t + 3]]
            def start_fills():
                cur_ = torch.cuda.current_stream(zero_jobs[0]["feats16"].device)
                aux = _aux_stream(zero_jobs[0]["feats16"].device)
                aux.wait_stream(cur_)
                with torch.cuda.stream(aux):
                    for d in zero_jobs:
                        d["dtable"] = torch.zeros((d["table_rows"], d["D"]), dtype=torch.float32, device=cur_.device)
                    ev_ = torch.cuda.Event()
                    ev_.record(aux)
                return cur_, ev_

            # data parallel keeps the fills inside this function (the path that was measured on 2 GPUs)
            late_fill = _LATE_ZERO_FILL and (_FUSED_ZERO_FILL or all(spec[1] is None for spec in specs))
            if zero_jobs and not late_fill:
                cur, ev = start_fills()
            # data parallel, every tower exchanging row-wise through the same object: the ids of ALL towers travel in
            # one all-gather on a side stream now; the backward then needs a single all-gather of row gradients
            xch = specs[0][1]
            ctx.ids_handle = None
            if (xch is not None and hasattr(xch, "start_ids") and all(sp[1] is xch for sp in specs)
                    and len({(d["B"], d["D8"], d["D"]) for d in tw}) == 1):
                ctx.ids_handle = xch.start_ids(torch.cat([d["ids"] for d in tw]), T)
            if casts:
                cast_batched(casts)
            fused = [d for d in tw if d["fused"]]
            layered = [d for d in tw if not d["fused"]]
            by_shape = {}
            for d in fused:  # one launch for up to 4 towers of the same shape
                by_shape.setdefault((d["F"], d["D"], d["DI"]), []).append(d)
            for group in by_shape.values():
                for i0 in range(0, len(group), 4):
                    tower_forward_fused(group[i0:i0 + 4])
            for d in fused:  # inputs are not needed after the launch (backward uses the bf16 copies)
                d.pop("feats"), d.pop("table")
            if layered:
                gather_batched(gathers)
                gemm_batched([dict(A=d["feats16"], B=d["w0_16"], M=d["B"], N=d["hid"], K=d["F"], bias=d["b0"], relu=True,
                                   out16=d["H16"]) for d in layered])
                gemm_batched([dict(A=d["H16"], B=d["w1_16"], M=d["B"], N=d["D"], K=d["hid"], bias=d["b1"],
                                   out16=d["X16"][:, d["D8"]:]) for d in layered])
                gemm_batched([dict(A=d["X16"], B=d["wt_16"], M=d["B"], N=d["DI"], K=d["KT"], bias=d["bt"], out32=d["emb"],
                                   out16=d["emb16"]) for d in layered])
            if zero_jobs and not late_fill:  # join again (the fills ran beside the kernels above)
                cur.wait_event(ev)
                for d in zero_jobs:
                    d["dtable"].record_stream(cur)
            elif zero_jobs and _FUSED_ZERO_FILL and all((d["table_rows"] * d["D"]) % 4 == 0 for d in zero_jobs):
                # the fills ride in the in-batch CE forward launch that follows (TMA bulk stores from an idle warp of the
                # scoring kernel); join_pending_fills() falls back to memsets when no such launch comes
                for d in zero_jobs:
                    d["dtable"] = torch.empty((d["table_rows"], d["D"]), dtype=torch.float32, device=d["feats16"].device)
                    _pending_zero.append(d["dtable"])
            elif zero_jobs:  # fork now; the caller joins after the loss forward (join_pending_fills), else backward does
                cur, ev = start_fills()
                _pending_fills.append((cur.device, ev))
                for d in zero_jobs:
                    d["dtable"].record_stream(cur)
        ctx.tw = tw
        outs = []
        for d in tw:
            outs.append(attach_shadow(d["emb"], d["emb16"]))
        return tuple(outs)

    @staticmethod
    def backward(ctx, *dembs):
        tw = ctx.tw
        T = len(tw)
        dev = dembs[0].device
        join_pending_fills()  # no-op when train_forward already joined the table-gradient fills
        casts = []
        for d, demb in zip(tw, dembs):
            if demb is None:  # this tower's embedding did not reach the loss
                demb = torch.zeros_like(d["emb"])
            d16 = shadow_of(demb)
            demb = _f32c(demb)
            if d16 is None:
                d16 = torch.empty((d["B"], _r8(d["DI"])), dtype=_BF16, device=dev)
                casts.append((demb, 0, d["DI"], d16, 0, _r8(d["DI"])))
            d["demb"], d["demb16"], d["demb_colsum"] = demb, d16, shadow_of(demb, "_tt_colsum")
        if casts:
            cast_batched(casts)
        # every accumulated gradient of every tower lives in one zero-filled arena (one memset)
        sizes = []
        for d in tw:
            sizes += [d["DI"] * d["KT"], d["KT"], d["DI"], d["D"] * d["hid"], d["hid"], d["hid"] * d["F"]]
        arena = _zero_arena(sizes, dev)
        for t, d in enumerate(tw):
            a = arena[6 * t: 6 * t + 6]
            d["dWt_p"], d["dXsum"], d["dbt"] = a[0].view(d["DI"], d["KT"]), a[1], a[2]
            d["dW1"], d["db0"], d["dW0"] = a[3].view(d["D"], d["hid"]), a[4], a[5].view(d["hid"], d["F"])
            d["dX16"] = torch.empty((d["B"], d["KT"]), dtype=_BF16, device=dev)
            d["dH16"] = torch.empty((d["B"], d["hid"]), dtype=_BF16, device=dev)
            if d["demb_colsum"] is not None:
                d["dbt"] = d["demb_colsum"]  # already reduced (fp32) by the kernel that produced demb
            else:
                colsum(d["demb"], d["DI"], out=d["dbt"])  # fp32 source: analytically-zero sums stay at fp32 noise
        # The gradient kernels are small and latency-bound, and only dX -> dH -> dW0 is a true chain: the weight
        # gradients that need no dX (tower Linear) or only dX (MLP layer 1) and the embedding scatter-adds run on side
        # streams beside it (parallel branches of the captured graph), joined before the results are handed back.
        cur = torch.cuda.current_stream(dev)
        par = _PARALLEL_BACKWARD and all(d["row_exchange"] is None for d in tw)  # collectives stay on one stream
        s1, s2 = (_aux_stream(dev, 1), _aux_stream(dev, 2)) if par else (cur, cur)
        for d in tw:  # buffers are allocated on the main stream; zero-filled beside the forward kernels when possible
            pre = d.pop("dtable", None)
            d["dtable_buf"] = pre if pre is not None else torch.zeros((d["table_rows"], d["D"]), dtype=torch.float32, device=dev)
        chain = (_FUSED_TOWER_BWD and all(d["fused"] and d["row_exchange"] is None for d in tw)
                 and len({(d["D"], d["DI"]) for d in tw}) == 1 and len(tw) <= 4)
        if chain:
            # candidate path: dX, dH, both table scatter-adds and the two bias column sums in ONE launch, then the three
            # weight-gradient GEMMs of every tower in one batched split-K launch
            tower_backward_chain_fused(tw)
            gemm_batched([dict(A=d["demb16"], B=d["X16"], M=d["DI"], N=d["KT"], K=d["B"], a_mn=True, b_mn=True,
                               out32=d["dWt_p"], accumulate=True) for d in tw])
            gemm_batched([dict(A=d["dX16"][:, d["D8"]:], B=d["H16"], M=d["D"], N=d["hid"], K=d["B"], a_mn=True, b_mn=True,
                               out32=d["dW1"], accumulate=True) for d in tw])
            gemm_batched([dict(A=d["dH16"], B=d["feats16"], M=d["hid"], N=d["F"], K=d["B"], a_mn=True, b_mn=True,
                               out32=d["dW0"], accumulate=True) for d in tw])
            for d in tw:
                d["dtable_out"] = d.pop("dtable_buf")
        if not chain:
            if par:
                s1.wait_stream(cur)
            with torch.cuda.stream(s1):  # dWt = demb^T X (split-K over the batch)
                gemm_batched([dict(A=d["demb16"], B=d["X16"], M=d["DI"], N=d["KT"], K=d["B"], a_mn=True, b_mn=True,
                                   out32=d["dWt_p"], accumulate=True) for d in tw])
            # dX = demb Wt (+ fp32 column sums = bias gradient of the second MLP layer)
            gemm_batched([dict(A=d["demb16"], B=d["wt_16"], M=d["B"], N=d["KT"], K=d["DI"], b_mn=True, out16=d["dX16"],
                               colsum=d["dXsum"]) for d in tw])
            if par:
                s1.wait_stream(cur)
                s2.wait_stream(cur)
            with torch.cuda.stream(s1):  # dW1 = dFe^T H
                gemm_batched([dict(A=d["dX16"][:, d["D8"]:], B=d["H16"], M=d["D"], N=d["hid"], K=d["B"], a_mn=True, b_mn=True,
                                   out32=d["dW1"], accumulate=True) for d in tw])
            with torch.cuda.stream(s2):  # id embeddings (dense gradient, duplicates accumulate)
                if ctx.ids_handle is not None:  # one all-gather of the row gradients of every tower
                    xch = tw[0]["row_exchange"]
                    rows_all = xch.gather_rows(torch.cat([d["dX16"][:, :d["D8"]] for d in tw]))
                    for t, d in enumerate(tw):
                        d["dtable_out"] = scatter_add_rows(rows_all, xch.ids_for(ctx.ids_handle, t), d["D"], d["table_rows"],
                                                           col_offset=0, grad=d.pop("dtable_buf"))
                for d in tw:
                    if "dtable_out" in d:
                        continue
                    pre = d.pop("dtable_buf")
                    if d["row_exchange"] is not None:
                        ids_all, rows_all = d["row_exchange"](d["ids"], d["dX16"][:, :d["D8"]].contiguous())
                        d["dtable_out"] = scatter_add_rows(rows_all, ids_all, d["D"], d["table_rows"], col_offset=0, grad=pre)
                    else:
                        d["dtable_out"] = scatter_add_rows(d["dX16"], d["ids"], d["D"], d["table_rows"], col_offset=0, grad=pre)
            # dH = (dFe W1) masked by ReLU (+ column sums = bias gradient of layer 0)
            gemm_batched([dict(A=d["dX16"][:, d["D8"]:], B=d["w1_16"], M=d["B"], N=d["hid"], K=d["D"], b_mn=True,
                               relu_mask=d["H16"], out16=d["dH16"], colsum=d["db0"]) for d in tw])
            gemm_batched([dict(A=d["dH16"], B=d["feats16"], M=d["hid"], N=d["F"], K=d["B"], a_mn=True, b_mn=True,
                               out32=d["dW0"], accumulate=True) for d in tw])
            if par:
                cur.wait_stream(s1)
                cur.wait_stream(s2)
        grads = []
        for d in tw:
            D, D8, E, KT = d["D"], d["D8"], d["E"], d["KT"]
            if D8 == D and _r8(E) == E:
                dWt = d["dWt_p"]
            else:
                parts = [d["dWt_p"][:, :D], d["dWt_p"][:, D8:D8 + D]] + ([d["dWt_p"][:, 2 * D8:2 * D8 + E]] if E else [])
                dWt = torch.cat(parts, dim=1)
            dtable = d.pop("dtable_out")
            dfeats = None
            if d["need_dfeats"]:
                dfeats = torch.empty((d["B"], d["F"]), dtype=torch.float32, device=dev)
                gemm(d["dH16"], d["w0_16"], d["B"], d["F"], d["hid"], b_mn=True, out32=dfeats)
            dextra = d["dX16"][:, 2 * D8:2 * D8 + E].float() if d["has_extra"] else None
            db1 = d["dXsum"][D8:D8 + D]
            # no reference to a gradient may stay behind in ctx: AccumulateGrad adopts a fresh gradient in place only
            # when it holds the last reference, otherwise it clones it (one copy launch per parameter)
            dW0, db0, dW1, dbt = d.pop("dW0"), d.pop("db0"), d.pop("dW1"), d.pop("dbt")
            for k in ("dWt_p", "dXsum", "demb", "demb16", "demb_colsum", "dX16", "dH16"):
                d.pop(k, None)
            grads += [None, dfeats, dextra, dtable, dW0, db0, dW1, db1, dWt, dbt]
            del dW0, db0, dW1, dbt, db1, dWt, dtable
        del arena
        return (None, None, *grads)


def _part_array(parts):
    import ctypes

    return (ctypes.c_void_p * len(parts))(*[int(p) for p in parts])


def inbatch_ce_forward_parts(U16, v_ptrs, rows_per_part, ldv, B, N, d, target_offset=0):
    """ce, lse of U against V = [parts...] given as device pointers (possibly peer memory), see tt_inbatch_ce_fwd_parts."""
    ce = torch.empty(B, dtype=torch.float32, device=U16.device)
    lse = torch.empty(B, dtype=torch.float32, device=U16.device)
    ws = _ce_workspace(B, N, d, U16.device)
    with _span("inbatch_ce_fwd"):
        _native.check(
            _native.lib().tt_inbatch_ce_fwd_parts(U16.data_ptr(), U16.stride(0), _part_array(v_ptrs), len(v_ptrs),
                                                  rows_per_part, ldv, B, N, d, target_offset, ce.data_ptr(), lse.data_ptr(),
                                                  ws.data_ptr(), ws.numel(), _stream()),
            "inbatch_ce_fwd_parts",
        )
    return ce, lse


def inbatch_ce_backward_parts(U16, v_ptrs, rows_per_part, ldv, B, N, d, target_offset, lse, g):
    dev = U16.device
    dU = torch.empty((B, d), dtype=torch.float32, device=dev)
    dV = torch.empty((N, d), dtype=torch.float32, device=dev)
    dU16 = torch.empty((B, _r8(d)), dtype=_BF16, device=dev)
    ws = _ce_workspace(B, N, d, dev)
    cs = torch.zeros((2, d), dtype=torch.float32, device=dev) if d <= 128 else None
    with _span("inbatch_ce_bwd"):
        _native.check(
            _native.lib().tt_inbatch_ce_bwd_parts(
                U16.data_ptr(), U16.stride(0), _part_array(v_ptrs), len(v_ptrs), rows_per_part, ldv, B, N, d, target_offset,
                lse.data_ptr(), g.data_ptr(), dU.data_ptr(), dU.stride(0), dU16.data_ptr(), dU16.stride(0),
                dV.data_ptr(), dV.stride(0), None, 0, _ptr(cs), None, ws.data_ptr(), ws.numel(), _stream()),
            "inbatch_ce_bwd_parts",
        )
    attach_shadow(dU, dU16, None if cs is None else cs[0])
    return dU, dV
