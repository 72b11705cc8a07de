"""Batch-sharded (data-parallel) training of the two-tower loss over the GPUs of one box.

The reference has no distributed code; this module is the multi-GPU generalisation that BASELINE.json's
north_star defines: one process per GPU (torchrun), every rank owns B_loc rows of the global batch.

    forward   V_loc (bf16) --all-gather--> V_glob [N = world*B_loc, d]
              ce_i = logsumexp_j(U_i . V_glob_j) - U_i . V_glob_{i + rank*B_loc}      (fused CE kernel, target offset)
              w_i  = clamp(nuv_i, 1e-6) / max over the GLOBAL batch   (all-reduce MAX of one float)
              loss = sum_i ce_i w_i / N  (+ all-reduce SUM of the detached scalar for reporting)
    backward  dU_loc is local;  dV_glob (contribution of the local users to every item)
              --reduce-scatter(SUM)--> dV_loc;  dense parameter gradients --all-reduce(SUM)-- (the 1/N factor
              is already inside dL/dce, so gradients are summed, not averaged).

Collectives per step (6, of which the id all-gather and the reduce-scatter run beside compute): all-gather(V) ->
all-gather of the per-rank (max, sum) loss statistics -> reduce-scatter(dV), started after the dV pass and hidden
behind the dU pass -> one all-gather of the touched embedding rows' gradients of BOTH towers (their ids were gathered
on a side stream during the forward) -> one flat all-reduce of the dense parameter gradients.

With these collectives a world_size-W run reproduces the single-process reference on the concatenated batch
(same loss, same gradients up to fp reduction order) - tests/test_distributed_cpu.py checks that with gloo.

The numeric kernels are injected through a small `kernels` object so that the collective logic can be
exercised on CPU/gloo by the tests (with the CPU oracle standing in for the kernels); the product default
is the CUDA path in `ops` and raises on CPU tensors.
"""
from typing import Optional

import torch
import torch.distributed as dist

from . import ops


class _CudaKernels:
    """Product kernels: libtt_b200.so through ops (CUDA only)."""

    @staticmethod
    def operand(x: torch.Tensor) -> torch.Tensor:
        x16 = ops.shadow_of(x)
        return x16 if x16 is not None else ops.cast_rows_bf16(ops._f32c(x))

    @staticmethod
    def ce_forward(U_op, V_op, B, N, d, offset):
        return ops.inbatch_ce_forward_raw(U_op, V_op, B, N, d, offset)

    @staticmethod
    def ce_backward(U_op, V_op, B, N, d, offset, lse, g):
        dU, dV, dU16, _ = ops.inbatch_ce_backward_raw(U_op, V_op, B, N, d, offset, lse, g)
        ops.attach_shadow(dU, dU16)
        return dU, dV

    # the two passes as separate launches, so that the reduce-scatter of dV overlaps the dU pass
    @staticmethod
    def ce_backward_dv(U_op, V_op, B, N, d, offset, lse, g, g_scale=None, g_scale2=None):
        _, dV, _, dV16 = ops.inbatch_ce_backward_raw(U_op, V_op, B, N, d, offset, lse, g, g_scale=g_scale,
                                                    g_scale2=g_scale2, want=("dV",))
        return dV, dV16

    @staticmethod
    def ce_backward_du(U_op, V_op, B, N, d, offset, lse, g, g_scale=None, g_scale2=None):
        cs = torch.zeros((2, d), dtype=torch.float32, device=U_op.device) if d <= 128 else None
        dU, _, dU16, _ = ops.inbatch_ce_backward_raw(U_op, V_op, B, N, d, offset, lse, g, g_scale=g_scale,
                                                    g_scale2=g_scale2, want=("dU",), colsums=cs)
        ops.attach_shadow(dU, dU16, None if cs is None else cs[0])
        return dU


def _reduce_scatter_sum(full: torch.Tensor, rank: int, world: int, group) -> torch.Tensor:
    rows = full.shape[0] // world
    if dist.get_backend(group) == "gloo":  # gloo has no reduce_scatter: all-reduce and keep the own shard
        dist.all_reduce(full, op=dist.ReduceOp.SUM, group=group)
        return full[rank * rows:(rank + 1) * rows].clone()
    out = torch.empty((rows,) + tuple(full.shape[1:]), dtype=full.dtype, device=full.device)
    dist.reduce_scatter_tensor(out, full.contiguous(), op=dist.ReduceOp.SUM, group=group)
    return out


_RS_BF16 = __import__("os").environ.get("TT_B200_RS_BF16", "1") == "1"


def _start_reduce_scatter_dv(dV_all, dV_all16, rank, world, group, d):  # group: the (CTA-limited) overlap group
    """Start the reduce-scatter of the items' gradient (async on NCCL's stream) and return a closure that waits for it
    and returns the local fp32 dV [B, d] (bf16 operand copy attached).  By default the bf16 copy travels (half the
    bytes; the tower backward consumes dV as a bf16 operand anyway) - TT_B200_RS_BF16=0 sends fp32."""
    rows = dV_all.shape[0] // world
    if dV_all16 is not None and _RS_BF16:
        out16 = torch.empty((rows, dV_all16.shape[1]), dtype=dV_all16.dtype, device=dV_all16.device)
        work = dist.reduce_scatter_tensor(out16, dV_all16, op=dist.ReduceOp.SUM, group=group, async_op=True)

        def finish():
            work.wait()
            dV = out16[:, :d].float()
            return ops.attach_shadow(dV, out16)
    else:
        out = torch.empty((rows, d), dtype=dV_all.dtype, device=dV_all.device)
        work = dist.reduce_scatter_tensor(out, dV_all.contiguous(), op=dist.ReduceOp.SUM, group=group, async_op=True)

        def finish():
            work.wait()
            return out
    return finish


class ShardedInBatchCE(torch.autograd.Function):
    """ce[B_loc] of the local users against the all-gathered items (positives at column row + rank*B_loc)."""

    @staticmethod
    def forward(ctx, U, V, group, kernels, overlap=None):
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        B, d = U.shape
        if V.shape != (B, d):
            raise RuntimeError(f"user/item embeddings must both be [{B}, {d}] on every rank, got {tuple(V.shape)}")
        U_op, V_op = kernels.operand(U), kernels.operand(V)
        V_all = torch.empty((world * B,) + tuple(V_op.shape[1:]), dtype=V_op.dtype, device=V_op.device)
        dist.all_gather_into_tensor(V_all, V_op.contiguous(), group=group)
        ce, lse = kernels.ce_forward(U_op, V_all, B, world * B, d, rank * B)
        ctx.save_for_backward(U_op, V_all, lse)
        ctx.meta = (B, d, rank, world, group, kernels)
        ctx.overlap = overlap if overlap is not None else (group, 0)
        return ce

    @staticmethod
    def backward(ctx, g):
        U_op, V_all, lse = ctx.saved_tensors
        B, d, rank, world, group, kernels = ctx.meta
        g = g.contiguous().float()
        if hasattr(kernels, "ce_backward_dv") and dist.get_backend(group) != "gloo":
            og, octas = ctx.overlap
            dV_all, dV_all16 = kernels.ce_backward_dv(U_op, V_all, B, world * B, d, rank * B, lse, g)
            finish = _start_reduce_scatter_dv(dV_all, dV_all16, rank, world, og, d)  # runs beside the dU pass
            with ops.sm_reserve(octas):
                dU = kernels.ce_backward_du(U_op, V_all, B, world * B, d, rank * B, lse, g)
            return dU, finish(), None, None, None
        dU, dV_all = kernels.ce_backward(U_op, V_all, B, world * B, d, rank * B, lse, g)
        dV = _reduce_scatter_sum(dV_all, rank, world, group)
        return dU, dV, None, None, None


class ShardedWeightedLoss(torch.autograd.Function):
    """compute_training_loss with the identity hook, batch-sharded, CUDA kernels only: all-gather(V), ONE fused launch for
    scores + softmax statistics + label weights + this rank's (max nuv, sum ce nuv), one all-gather of those pairs and
    a one-thread kernel that folds them into the global loss and the scalar the backward kernels multiply into g."""

    @staticmethod
    def forward(ctx, U, V, labels, weights, group, overlap=None):
        from . import _native

        rank, world = dist.get_rank(group), dist.get_world_size(group)
        B, d = U.shape
        if V.shape != (B, d):
            raise RuntimeError(f"user/item embeddings must both be [{B}, {d}] on every rank, got {tuple(V.shape)}")
        U_op, V_op = _CudaKernels.operand(U), _CudaKernels.operand(V)
        dev = U_op.device
        N = world * B
        V_all = torch.empty((N,) + tuple(V_op.shape[1:]), dtype=V_op.dtype, device=dev)
        dist.all_gather_into_tensor(V_all, V_op.contiguous(), group=group)
        labels, weights = ops._f32c(labels), ops._f32c(weights)
        out = torch.empty(3 * B + 1, dtype=torch.float32, device=dev)  # ce | lse | g | g_norm (saved for backward)
        ce, lse, g, g_norm = out[:B], out[B:2 * B], out[2 * B:3 * B], out[3 * B:]
        loss = torch.empty((), dtype=torch.float32, device=dev)
        stats = torch.empty(4, dtype=torch.float32, device=dev)
        ws = ops._ce_workspace(B, N, d, dev)
        L = _native.lib()
        ops._attach_pending_zero()  # the dense table-gradient zero fills of the step ride in this launch
        _native.check(
            L.tt_inbatch_ce_loss_fwd_sharded(U_op.data_ptr(), U_op.stride(0), V_all.data_ptr(), V_all.stride(0), B, N, d,
                                             rank * B, labels.data_ptr(), labels.stride(0), weights.data_ptr(),
                                             labels.shape[1], ce.data_ptr(), lse.data_ptr(), g.data_ptr(), stats.data_ptr(),
                                             ws.data_ptr(), ws.numel(), ops._stream()),
            "inbatch_ce_loss_fwd_sharded")
        stats_all = torch.empty(2 * world, dtype=torch.float32, device=dev)
        dist.all_gather_into_tensor(stats_all, stats[:2], group=group)
        _native.check(L.tt_sharded_loss_finalize(stats_all.data_ptr(), world, N, loss.data_ptr(), g_norm.data_ptr(),
                                                 ops._stream()), "sharded_loss_finalize")
        ctx.save_for_backward(U_op, V_all, lse, g, g_norm)
        ctx.meta = (B, d, rank, world, group)
        ctx.overlap = overlap if overlap is not None else (group, 0)
        return loss

    @staticmethod
    def backward(ctx, dloss):
        U_op, V_all, lse, g, g_norm = ctx.saved_tensors
        B, d, rank, world, group = ctx.meta
        gs = ops._f32c(dloss).reshape(1)
        K = _CudaKernels
        og, octas = ctx.overlap
        dV_all, dV_all16 = K.ce_backward_dv(U_op, V_all, B, world * B, d, rank * B, lse, g, g_scale=gs, g_scale2=g_norm)
        finish = _start_reduce_scatter_dv(dV_all, dV_all16, rank, world, og, d)  # runs beside the dU pass
        with ops.sm_reserve(octas):
            dU = K.ce_backward_du(U_op, V_all, B, world * B, d, rank * B, lse, g, g_scale=gs, g_scale2=g_norm)
        return dU, finish(), None, None, None, None


class _PeerItemBuffers:
    """One symmetric-memory buffer per rank for the bf16 item embeddings: every rank writes its shard into its own
    buffer and the scoring kernels of all ranks read all shards IN PLACE over NVLink (TMA loads from the peers'
    memory), so the all-gather of V disappears into the kernel that consumes it."""

    def __init__(self, group, rows, pitch, device):
        import torch.distributed._symmetric_memory as symm

        pg = group if group is not None else dist.group.WORLD
        try:
            symm.enable_symm_mem_for_group(pg.group_name)
        except Exception:
            pass
        self.buf = symm.empty((rows, pitch), dtype=torch.bfloat16, device=device)
        self.hdl = symm.rendezvous(self.buf, pg)
        self.ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        self.generation = 0

    def publish(self, v16: torch.Tensor) -> int:
        self.hdl.barrier(channel=0)  # every rank is done reading the previous step's shards
        self.buf.copy_(v16)
        self.hdl.barrier(channel=1)  # every shard is in place
        self.generation += 1
        return self.generation


class PeerShardedInBatchCE(torch.autograd.Function):
    """ShardedInBatchCE without the all-gather: the item shards are read from the peers' symmetric buffers."""

    @staticmethod
    def forward(ctx, U, V, group, peers: _PeerItemBuffers):
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        B, d = U.shape
        U_op, V_op = _CudaKernels.operand(U), _CudaKernels.operand(V)
        gen = peers.publish(V_op)
        ce, lse = ops.inbatch_ce_forward_parts(U_op, peers.ptrs, B, peers.buf.stride(0), B, world * B, d, rank * B)
        ctx.save_for_backward(U_op, lse)
        ctx.meta = (B, d, rank, world, group, peers, gen)
        return ce

    @staticmethod
    def backward(ctx, g):
        U_op, lse = ctx.saved_tensors
        B, d, rank, world, group, peers, gen = ctx.meta
        if gen != peers.generation:
            raise RuntimeError("the peer item buffers were overwritten by a later forward; run backward before the next "
                               "forward or disable peer-memory scoring (TT_B200_PEER_CE=0)")
        dU, dV_all = ops.inbatch_ce_backward_parts(U_op, peers.ptrs, B, peers.buf.stride(0), B, world * B, d, rank * B,
                                                   lse, g.contiguous().float())
        dV = _reduce_scatter_sum(dV_all, rank, world, group)
        return dU, dV, None, None


class _RowExchange:
    """Sparse exchange of the id-embedding gradients: every rank contributes its touched (id, row gradient) pairs instead
    of all-reducing dense [hash, D] tables.  One object serves all towers of a step:
      start_ids(ids of all towers, concatenated)  - forward: all-gather on a side stream, long before it is needed;
      gather_rows(row gradients of all towers)     - backward: ONE all-gather for every tower;
      ids_for(handle, t)                           - tower t's ids in the gathered order, the other towers' slots masked
                                                     to -1 (the scatter-add kernel skips them without touching the row).
    Called as a function (ids, rows) it exchanges one tower on its own (ops.TowerFunction)."""

    def __init__(self, group, world):
        self.group, self.world = group, world
        self._masks = {}

    def __call__(self, ids, rows):
        world, group = self.world, self.group
        ids_all = torch.empty((world * ids.shape[0],), dtype=ids.dtype, device=ids.device)
        rows_all = torch.empty((world * rows.shape[0], rows.shape[1]), dtype=rows.dtype, device=rows.device)
        dist.all_gather_into_tensor(ids_all, ids.contiguous(), group=group)
        dist.all_gather_into_tensor(rows_all, rows, group=group)
        return ids_all, rows_all

    def _mask(self, towers, rows, t, device):
        key = (towers, rows, t, str(device))
        m = self._masks.get(key)
        if m is None:
            slot = torch.arange(towers * rows, device=device).div(rows, rounding_mode="floor").repeat(self.world)
            m = slot != t
            self._masks[key] = m
        return m

    def start_ids(self, ids_cat, towers):
        """ids_cat: int64 [towers * rows].  Returns a handle for ids_for()."""
        dev = ids_cat.device
        rows = ids_cat.shape[0] // towers
        masks = [self._mask(towers, rows, t, dev) for t in range(towers)]  # (built outside the side stream once)
        side = None
        if ids_cat.is_cuda:
            cur = torch.cuda.current_stream(dev)
            side = ops._aux_stream(dev, 3)
            side.wait_stream(cur)
        ctxm = torch.cuda.stream(side) if side is not None else __import__("contextlib").nullcontext()
        with ctxm:
            ids_all = torch.empty((self.world * ids_cat.shape[0],), dtype=ids_cat.dtype, device=dev)
            dist.all_gather_into_tensor(ids_all, ids_cat, group=self.group)
            per_tower = [ids_all.masked_fill(m, -1) for m in masks]
            ev = None
            if side is not None:
                ev = torch.cuda.Event()
                ev.record(side)
        if side is not None:
            ids_cat.record_stream(side)
        return per_tower, ev

    def ids_for(self, handle, t):
        per_tower, ev = handle
        if ev is not None:
            torch.cuda.current_stream(per_tower[t].device).wait_event(ev)
            per_tower[t].record_stream(torch.cuda.current_stream(per_tower[t].device))
        return per_tower[t]

    def gather_rows(self, rows_cat):
        rows_all = torch.empty((self.world * rows_cat.shape[0], rows_cat.shape[1]), dtype=rows_cat.dtype, device=rows_cat.device)
        dist.all_gather_into_tensor(rows_all, rows_cat.contiguous(), group=self.group)
        return rows_all


class DataParallelContext:
    def __init__(self, group=None, kernels=None, peer_memory=None):
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised (launch with torchrun, one process per GPU)")
        self.group = group
        self.kernels = kernels or _CudaKernels
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self._presynced = set()  # ids of parameters whose gradient is already the global sum after backward
        if peer_memory is None:
            import os

            peer_memory = os.environ.get("TT_B200_PEER_CE", "0") == "1"
        self.peer_memory = bool(peer_memory) and kernels is None
        self._peers = None
        self._row_exchange = None
        # Communicator for the collective that runs BESIDE a kernel (reduce-scatter of dV during the dU pass): limited to a
        # few CTAs, and the dU pass leaves that many SMs free (ops.sm_reserve) - with NCCL's default CTA count the 148-CTA
        # persistent kernel waits for the SMs NCCL holds and its tail grows by the collective's duration.
        self.overlap_group, self.overlap_ctas = group, 0
        import os as _os

        # Measured on 8 GPUs (profiles/r02_summary.md): 0 / 16 / 32 CTAs give 1.103 / 1.107 / 1.138 ms per step - the tail of
        # the dU kernel shrinks (307 -> 267 us) but the slower collective takes it back, so the default stays off.
        ctas = int(_os.environ.get("TT_B200_RS_CTAS", "0"))
        if kernels is None and ctas > 0 and dist.get_backend(group) == "nccl":
            try:
                opts = dist.ProcessGroupNCCL.Options()
                opts.config.max_ctas = ctas
                opts.config.min_ctas = 1
                ranks = dist.get_process_group_ranks(group) if group is not None else None
                self.overlap_group = dist.new_group(ranks=ranks, backend="nccl", pg_options=opts)
                self.overlap_ctas = ctas
            except Exception:  # older torch / NCCL: keep the main group, no SM reservation
                self.overlap_group, self.overlap_ctas = group, 0

    def compute_training_loss(self, model, user_embedding, item_embeddings, position, labels):
        """Sharded version of TwoTowerBaseRetrieval.compute_training_loss (reference :279-347); the
        debias_net_user_value hook is called exactly as in the single-GPU path."""
        if self.peer_memory and user_embedding.shape[0] % 128 == 0:
            B, dcols = item_embeddings.shape
            pitch = (dcols + 7) // 8 * 8
            if self._peers is None or tuple(self._peers.buf.shape) != (B, pitch):
                self._peers = _PeerItemBuffers(self.group, B, pitch, item_embeddings.device)
            ce = PeerShardedInBatchCE.apply(user_embedding, item_embeddings, self.group, self._peers)  # [B_loc]
        else:
            from .towers import TwoTowerBaseRetrieval

            if (self.kernels is _CudaKernels and user_embedding.is_cuda
                    and getattr(type(model), "debias_net_user_value", None) is TwoTowerBaseRetrieval.debias_net_user_value
                    and labels.dim() == 2 and labels.shape[1] == model.user_value_weights.shape[0]
                    and not labels.requires_grad and user_embedding.shape[0] <= 65536):
                # identity hook: weights, batch max and the weighted mean ride in the CE's merge kernel on every rank
                return ShardedWeightedLoss.apply(user_embedding, item_embeddings, labels, model.user_value_weights, self.group,
                                                 (self.overlap_group, self.overlap_ctas))
            ce = ShardedInBatchCE.apply(user_embedding, item_embeddings, self.group, self.kernels,
                                        (self.overlap_group, self.overlap_ctas))  # [B_loc]
        net_user_value = torch.sum(labels * model.user_value_weights, dim=-1)
        net_user_value, additional_loss = model.debias_net_user_value(
            net_user_value=net_user_value, position=position, user_embedding=user_embedding
        )
        net_user_value = torch.clamp(net_user_value, min=0.000001)
        # ONE small all-gather carries every rank's (max weight, sum ce * weight, additional loss): the batch-global max
        # (reference :339), the global mean (:343) and the global sum of the hook's additional loss (:346)
        has_aux = torch.is_tensor(additional_loss)
        s_loc = torch.sum(ce * net_user_value)
        vec = torch.stack([torch.max(net_user_value).detach().reshape(()).float(), s_loc.detach().float(),
                           (additional_loss.detach().reshape(()).float() if has_aux
                            else torch.zeros((), dtype=torch.float32, device=ce.device))])
        gathered = torch.empty(3 * self.world, dtype=torch.float32, device=ce.device)
        dist.all_gather_into_tensor(gathered, vec, group=self.group)
        gathered = gathered.view(self.world, 3)
        scale = 1.0 / (gathered[:, 0].max() * (ce.shape[0] * self.world))
        # gradient = this rank's share (gradients are summed across ranks afterwards; the additional loss is a sum over
        # the batch in the reference, so its local part enters unscaled); value = the global loss
        local = s_loc * scale + (additional_loss if has_aux else 0.0)
        value = gathered[:, 1].sum() * scale + (gathered[:, 2].sum() if has_aux else additional_loss)
        return local + (value - local.detach())

    def row_exchange(self, model, table):
        """Sparse exchange of an id-embedding table's gradient: every rank contributes its B_loc touched
        (id, row-gradient) pairs (2 MB at B_loc = 8192, D = 128) instead of all-reducing the dense [hash, D]
        gradient (51 MB).  Returns a callable for ops.TowerFunction, or None when the table also receives
        other, rank-local gradient contributions (the history lookup shares the item table) - then the dense
        all-reduce in sync_gradients handles it."""
        if getattr(model, "user_history_encoder", None) is not None and table is model.item_id_embedding_arch.weight:
            return None
        self._presynced.add(id(table))
        if self._row_exchange is None:
            self._row_exchange = _RowExchange(self.group, self.world)
        return self._row_exchange

    def sync_gradients(self, model) -> None:
        """Sum parameter gradients over the ranks (call after loss.backward()).  Dense parameters travel in
        one flat bucket; embedding tables whose gradient was already exchanged row-wise (row_exchange) are
        skipped, any other large gradient is all-reduced as a dense tensor."""
        small, big = [], []
        for p in model.parameters():
            if p.grad is None or id(p) in self._presynced:
                continue
            (big if p.grad.numel() >= (1 << 20) else small).append(p.grad)
        if small:
            flat = torch.cat([g.reshape(-1) for g in small])
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
            off = 0
            for g in small:
                n = g.numel()
                g.copy_(flat[off:off + n].view_as(g))
                off += n
        for g in big:
            dist.all_reduce(g, op=dist.ReduceOp.SUM, group=self.group)


def enable_data_parallel(model, group=None, kernels=None, peer_memory=None) -> DataParallelContext:
    """Shard `model.compute_training_loss` over the process group (see module docstring).  peer_memory=True (or
    TT_B200_PEER_CE=1) replaces the NCCL all-gather of the item embeddings by in-kernel TMA reads of the peers'
    symmetric-memory buffers."""
    ctx = DataParallelContext(group, kernels, peer_memory)
    model._dp = ctx
    return ctx


def disable_data_parallel(model) -> None:
    model._dp = None
