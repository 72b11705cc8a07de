"""2-GPU NCCL test of the batch-sharded loss with the real CUDA kernels (skipped with fewer than 2 GPUs):
the sharded run must reproduce the single-GPU run of the same global batch."""
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
def test_two_gpu_sharded_step_matches_single_gpu(tmp_path, peer):
    from helpers import assert_close_fro

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "rank0.pt")
    mp.spawn(_worker, args=(2, port, out, peer), nprocs=2, join=True)
    got = torch.load(out)
    p, uvw, batch = _problem()
    m = _build(p, uvw).cuda()
    loss = m.train_forward(*[batch[k].cuda() for k in ORDER])
    loss.backward()
    assert abs(float(got["loss"]) - float(loss.detach())) <= 1e-4 * abs(float(loss.detach()))
    for k, t in m.named_parameters():
        assert_close_fro(got["grads"][k], t.grad, rtol=2e-2, atol=1e-4 if t.dim() == 1 else 2e-6, what=k)
