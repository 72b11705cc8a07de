"""world_size-2 gloo test of the batch-sharded loss (two_tower_models_b200/distributed.py) on CPU.

The collective logic (all-gather of item embeddings, target offset, batch-global max, reduce-scatter of dV,
summed gradients) is exercised with the CPU oracle standing in for the CUDA kernels; the result must equal
the single-process oracle on the concatenated global batch - the equivalence SURVEY 8(e) defines."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

def _spawn_with_free_port(fn, make_args, nprocs=2, attempts=4):
    """mp.spawn(fn, args=make_args(port)) on a free local port; the port can be taken between probing and binding
    (EADDRINUSE on a busy box), so a failed rendezvous is retried on another port."""
    last = None
    for _ in range(attempts):
        s = socket.socket()
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
        s.close()
        try:
            mp.spawn(fn, args=make_args(port), nprocs=nprocs, join=True)
            return
        except Exception as exc:  # noqa: BLE001 - only the address-in-use case is retried
            last = exc
            if "EADDRINUSE" not in str(exc) and "address already in use" not in str(exc).lower():
                raise
    raise last


class OracleKernels:
    """CPU stand-ins with the kernel interface of distributed._CudaKernels (fp32, exact)."""

    @staticmethod
    def operand(x):
        return x.detach()

    @staticmethod
    def ce_forward(U, V_all, B, N, d, offset):
        import oracle

        return oracle.inbatch_ce(U, V_all, offset)

    @staticmethod
    def ce_backward(U, V_all, B, N, d, offset, lse, g):
        import oracle

        return oracle.inbatch_ce_backward(U, V_all, lse, g, offset)


class _Towers(torch.nn.Module):
    def __init__(self, params, uvw):
        super().__init__()
        self.p = torch.nn.ParameterDict({k.replace(".", "/"): torch.nn.Parameter(v.clone()) for k, v in params.items()})
        self.user_value_weights = uvw
        self._dp = None

    def named(self):
        return {k.replace("/", "."): v for k, v in self.p.items()}

    def debias_net_user_value(self, net_user_value, position, user_embedding):
        return net_user_value, 0


def _problem():
    from test_gpu_models import _random_base_params, _random_batch

    p = _random_base_params(24, 24, 12, 12, 40, 40, seed=5)
    batch = _random_batch(64, 12, 12, 40, 40, 2, seed=6)
    return p, torch.tensor([1.0, 0.5]), batch


class _HookedTowers(_Towers):
    """Non-identity debias hook: position-dependent weights plus an additional loss that is a SUM over the batch (like
    the reference's mse_loss(reduction="sum") in src/two_tower_with_position_debiased_weights.py:101-103) and depends
    on a parameter and on the user embedding."""

    def debias_net_user_value(self, net_user_value, position, user_embedding):
        w = 1.0 / (1.0 + 0.1 * position.float())
        est = user_embedding[:, 0] * self.p["user_tower_arch/bias"][0] + 0.3
        return net_user_value * w + 0.05, torch.sum((est - net_user_value) ** 2)


def _hooked_single_process(p, uvw, batch):
    """The same model on the concatenated batch in one process (reference :279-347 with the hook above)."""
    import oracle

    m = _HookedTowers(p, uvw)
    P = m.named()
    u = oracle.base_user_embedding(P, batch["user_id"], batch["user_features"])
    v = oracle.base_item_embedding(P, batch["item_id"], batch["item_features"])
    ce, _ = oracle.inbatch_ce(u, v)
    nuv = torch.sum(batch["labels"] * uvw, dim=-1)
    nuv, aux = m.debias_net_user_value(nuv, batch["position"], u)
    nuv = torch.clamp(nuv, min=0.000001)
    nuv = nuv / torch.max(nuv)
    loss = torch.mean(ce * nuv) + aux
    loss.backward()
    return loss.detach(), {k: t.grad for k, t in P.items()}


def _worker(rank, world, port, out, hooked=False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle
        from two_tower_models_b200 import distributed as ttd

        torch.set_num_threads(1)
        p, uvw, batch = _problem()
        n = batch["user_id"].shape[0] // world
        sl = slice(rank * n, (rank + 1) * n)
        loc = {k: v[sl] for k, v in batch.items()}
        m = (_HookedTowers if hooked else _Towers)(p, uvw)
        ctx = ttd.enable_data_parallel(m, kernels=OracleKernels)
        P = m.named()
        u = oracle.base_user_embedding(P, loc["user_id"], loc["user_features"])
        v = oracle.base_item_embedding(P, loc["item_id"], loc["item_features"])
        loss = ctx.compute_training_loss(m, u, v, loc["position"], loc["labels"])
        loss.backward()
        ctx.sync_gradients(m)
        if rank == 0:
            torch.save({"loss": loss.detach(), "grads": {k: t.grad for k, t in P.items()}}, out)
    finally:
        dist.destroy_process_group()


def test_sharded_loss_equals_single_process_reference(tmp_path):
    import oracle
    from helpers import assert_close_fro

    out = str(tmp_path / "rank0.pt")
    _spawn_with_free_port(_worker, lambda port: (2, port, out))
    got = torch.load(out)
    p, uvw, batch = _problem()
    ref_loss, ref_grads = oracle.base_train_forward_with_grads(p, uvw, batch)
    assert abs(float(got["loss"]) - float(ref_loss)) <= 1e-6 * abs(float(ref_loss))
    for k, g in ref_grads.items():
        assert_close_fro(got["grads"][k], g, rtol=1e-5, atol=1e-7, what=k)  # atol: analytically-zero item-side biases


def test_sharded_loss_with_a_debias_hook_and_additional_loss(tmp_path):
    """The hook's additional loss is a sum over the batch: its gradient must NOT be divided by the world size, and the
    reported value must be the global sum (distributed.py scaled it by 1 / world)."""
    from helpers import assert_close_fro

    out = str(tmp_path / "rank0.pt")
    _spawn_with_free_port(_worker, lambda port: (2, port, out, True))
    got = torch.load(out)
    p, uvw, batch = _problem()
    ref_loss, ref_grads = _hooked_single_process(p, uvw, batch)
    assert abs(float(got["loss"]) - float(ref_loss)) <= 1e-5 * abs(float(ref_loss))
    for k, g in ref_grads.items():
        assert_close_fro(got["grads"][k], g, rtol=1e-4, atol=1e-6, what=k)


def test_requires_process_group():
    from two_tower_models_b200 import distributed as ttd

    if dist.is_initialized():
        pytest.skip("a process group is live")
    with pytest.raises(RuntimeError):
        ttd.DataParallelContext()
