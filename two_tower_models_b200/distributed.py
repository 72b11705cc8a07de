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


def _reduce_scatter_sum(full: torch.Tensor, rank: int, world: int, group) -> torch.Tensor:
    rows = full.shape[0] // world
    if dist.get_backend(group) == "gloo":  # gloo has no reduce_scatter: all-reduce and keep the own shard
        dist.all_reduce(full, op=dist.ReduceOp.SUM, group=group)
        return full[rank * rows:(rank + 1) * rows].clone()
    out = torch.empty((rows,) + tuple(full.shape[1:]), dtype=full.dtype, device=full.device)
    dist.reduce_scatter_tensor(out, full.contiguous(), op=dist.ReduceOp.SUM, group=group)
    return out


class ShardedInBatchCE(torch.autograd.Function):
    """ce[B_loc] of the local users against the all-gathered items (positives at column row + rank*B_loc)."""

    @staticmethod
    def forward(ctx, U, V, group, kernels):
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
        return ce

    @staticmethod
    def backward(ctx, g):
        U_op, V_all, lse = ctx.saved_tensors
        B, d, rank, world, group, kernels = ctx.meta
        dU, dV_all = kernels.ce_backward(U_op, V_all, B, world * B, d, rank * B, lse, g.contiguous().float())
        dV = _reduce_scatter_sum(dV_all, rank, world, group)
        return dU, dV, None, None


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
            ce = ShardedInBatchCE.apply(user_embedding, item_embeddings, self.group, self.kernels)  # [B_loc]
        net_user_value = torch.sum(labels * model.user_value_weights, dim=-1)
        net_user_value, additional_loss = model.debias_net_user_value(
            net_user_value=net_user_value, position=position, user_embedding=user_embedding
        )
        net_user_value = torch.clamp(net_user_value, min=0.000001)
        gmax = torch.max(net_user_value).detach().clone()
        dist.all_reduce(gmax, op=dist.ReduceOp.MAX, group=self.group)  # batch-global max (reference :339)
        net_user_value = net_user_value / gmax
        n_global = ce.shape[0] * self.world
        local = torch.sum(ce * net_user_value) / n_global  # this rank's share of mean over the global batch
        total = local.detach().clone()
        dist.all_reduce(total, op=dist.ReduceOp.SUM, group=self.group)
        # value = global loss, gradient = this rank's share (gradients are summed across ranks afterwards)
        loss = local + (total - local.detach())
        if torch.is_tensor(additional_loss):
            return loss + additional_loss / self.world
        return loss + additional_loss

    def row_exchange(self, model, table):
        """Sparse exchange of an id-embedding table's gradient: every rank contributes its B_loc touched
        (id, row-gradient) pairs (2 MB at B_loc = 8192, D = 128) instead of all-reducing the dense [hash, D]
        gradient (51 MB).  Returns a callable for ops.TowerFunction, or None when the table also receives
        other, rank-local gradient contributions (the history lookup shares the item table) - then the dense
        all-reduce in sync_gradients handles it."""
        if getattr(model, "user_history_encoder", None) is not None and table is model.item_id_embedding_arch.weight:
            return None
        self._presynced.add(id(table))
        group, world = self.group, self.world

        def exchange(ids, rows):
            ids_all = torch.empty((world * ids.shape[0],), dtype=ids.dtype, device=ids.device)
            rows_all = torch.empty((world * rows.shape[0], rows.shape[1]), dtype=rows.dtype, device=rows.device)
            dist.all_gather_into_tensor(ids_all, ids.contiguous(), group=group)
            dist.all_gather_into_tensor(rows_all, rows, group=group)
            return ids_all, rows_all

        return exchange

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
