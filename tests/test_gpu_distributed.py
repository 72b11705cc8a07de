"""2-GPU NCCL test of the batch-sharded loss with the real CUDA kernels (skipped with fewer than 2 GPUs): the sharded
run must reproduce the single-process ORACLE (fp32 CPU restatement of the reference) on the concatenated global batch -
loss and every parameter gradient - within the tolerances of tests/test_gpu_models.py."""
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

pytestmark = pytest.mark.gpu

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


def _build(p, uvw):
    import two_tower_models_b200 as tt

    d = p["user_id_embedding_arch.weight"].shape[1]
    m = tt.TwoTowerBaseRetrieval(10, p["user_id_embedding_arch.weight"].shape[0], d, p["user_features_arch.0.weight"].shape[1],
                                 p["item_id_embedding_arch.weight"].shape[0], d, p["item_features_arch.0.weight"].shape[1],
                                 uvw.tolist(), tt.BaselineMIPSModule(16, d))
    m.load_state_dict(p, strict=True)
    return m


def _problem():
    from test_gpu_models import _random_base_params, _random_batch

    p = _random_base_params(128, 128, 64, 64, 500, 500, seed=11)
    return p, torch.tensor([1.0, 0.5]), _random_batch(1024, 64, 64, 500, 500, 2, seed=12)


ORDER = ["user_id", "user_features", "user_history", "item_id", "item_features", "position", "labels"]


def _worker(rank, world, port, out, peer):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from two_tower_models_b200 import distributed as ttd

        p, uvw, batch = _problem()
        n = batch["user_id"].shape[0] // world
        loc = {k: v[rank * n:(rank + 1) * n].cuda() for k, v in batch.items()}
        m = _build(p, uvw).cuda()
        ctx = ttd.enable_data_parallel(m, peer_memory=peer)
        loss = m.train_forward(*[loc[k] for k in ORDER])
        loss.backward()
        ctx.sync_gradients(m)
        torch.cuda.synchronize()
        if rank == 0:
            torch.save({"loss": loss.detach().cpu(), "grads": {k: t.grad.cpu() for k, t in m.named_parameters()}}, out)
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("peer", [False, True], ids=["nccl_allgather", "peer_memory"])
def test_two_gpu_sharded_step_matches_oracle_on_concatenated_batch(tmp_path, peer):
    from helpers import assert_close_fro

    out = str(tmp_path / "rank0.pt")
    _spawn_with_free_port(_worker, lambda port: (2, port, out, peer))
    got = torch.load(out)
    import oracle

    p, uvw, batch = _problem()
    ref_loss, ref_grads = oracle.base_train_forward_with_grads(p, uvw, batch)  # one process, the whole 1024-row batch
    assert abs(float(got["loss"]) - float(ref_loss)) <= 1e-3 * abs(float(ref_loss)), (float(got["loss"]), float(ref_loss))
    for k, g in ref_grads.items():
        assert_close_fro(got["grads"][k], g, rtol=5e-2, atol=1e-4 if g.dim() == 1 else 2e-6, what=k)


def _hook_model(p, uvw):
    """Base model with a non-identity debias hook: position-dependent weights and an additional loss that is a SUM over the
    batch (like the reference's mse_loss(reduction="sum"), src/two_tower_with_position_debiased_weights.py:101-103) and
    depends on a parameter of its own and on the user embedding."""
    import two_tower_models_b200 as tt

    class Hooked(tt.TwoTowerBaseRetrieval):
        def __init__(self, *a, **k):
            super().__init__(*a, **k)
            self.position_scale = torch.nn.Parameter(torch.tensor([0.1, 0.3]))

        def debias_net_user_value(self, net_user_value, position, user_embedding):
            w = 1.0 / (1.0 + self.position_scale[0] * position.float())
            est = user_embedding[:, 0] * self.position_scale[1] + 0.3
            return net_user_value * w + 0.05, torch.sum((est - net_user_value) ** 2) * 1e-3

    d = p["user_id_embedding_arch.weight"].shape[1]
    m = Hooked(10, p["user_id_embedding_arch.weight"].shape[0], d, p["user_features_arch.0.weight"].shape[1],
               p["item_id_embedding_arch.weight"].shape[0], d, p["item_features_arch.0.weight"].shape[1],
               uvw.tolist(), tt.BaselineMIPSModule(16, d))
    m.load_state_dict(p, strict=False)
    return m


def _hook_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from two_tower_models_b200 import distributed as ttd

        p, uvw, batch = _problem()
        n = batch["user_id"].shape[0] // world
        loc = {k: v[rank * n:(rank + 1) * n].cuda() for k, v in batch.items()}
        m = _hook_model(p, uvw).cuda()
        ctx = ttd.enable_data_parallel(m)
        loss = m.train_forward(*[loc[k] for k in ORDER])
        loss.backward()
        ctx.sync_gradients(m)
        torch.cuda.synchronize()
        if rank == 0:
            torch.save({"loss": loss.detach().cpu(), "grads": {k: t.grad.cpu() for k, t in m.named_parameters()}}, out)
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_sharded_step_with_debias_hook_matches_single_gpu(tmp_path):
    """Non-identity hook with an additional loss: the generic sharded path - per-row ce, one all-gather of (max, sum,
    additional loss) - against the same model on ONE GPU with the whole batch (the CPU twin of this test, with the oracle
    as reference, is tests/test_distributed_cpu.py::test_sharded_loss_with_a_debias_hook_and_additional_loss)."""
    from helpers import assert_close_fro

    out = str(tmp_path / "rank0.pt")
    _spawn_with_free_port(_hook_worker, lambda port: (2, port, out))
    got = torch.load(out)
    p, uvw, batch = _problem()
    m = _hook_model(p, uvw).cuda()
    loss = m.train_forward(*[batch[k].cuda() for k in ORDER])
    loss.backward()
    assert abs(float(got["loss"]) - float(loss.detach())) <= 1e-4 * abs(float(loss.detach())), (float(got["loss"]), float(loss))
    for k, t in m.named_parameters():
        assert_close_fro(got["grads"][k], t.grad, rtol=2e-2, atol=1e-4 if t.dim() == 1 else 2e-6, what=k)
