"""CPU-side checks of the boundary: the C-ABI library loads without a GPU and exports every symbol that
include/tt_b200.h declares; compute entry points fail loudly (no CPU fallback); host mirrors keep the
reference's names."""
import ctypes
import os
import re

import pytest
import torch

from helpers import load_golden, section

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "tt_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tt_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from two_tower_models_b200 import _native

    names = _declared()
    assert len(names) >= 18, names
    lib = ctypes.CDLL(_native.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/tt_b200.h but not exported"
    assert set(_native.SIGNATURES) == set(names), set(_native.SIGNATURES) ^ set(names)
    assert _native.lib().tt_abi_version() == 1


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only behaviour")
def test_no_cpu_fallback():
    import two_tower_models_b200 as tt
    from two_tower_models_b200 import _native, ops

    L = _native.lib()
    # a compute entry point without a CUDA device returns an error code and a message
    rc = L.tt_inbatch_ce_fwd(None, 8, None, 8, 128, 128, 64, 0, None, None, None, 1 << 30, None)
    assert rc != 0 and len(L.tt_last_error()) > 0
    with pytest.raises(RuntimeError):
        ops.inbatch_cross_entropy(torch.randn(8, 16), torch.randn(8, 16))
    m = tt.BaselineMIPSModule(16, 8)
    with pytest.raises(RuntimeError):
        m(torch.randn(2, 8), 3)
    enc = tt.UserHistoryEncoder(8, 4, 2, 1, True)
    with pytest.raises(RuntimeError):
        enc(torch.randn(2, 4, 8))


def test_host_mirrors_keep_reference_names():
    import two_tower_models_b200 as tt

    g = load_golden("hist_small.npz")
    p = section(g, "p:")
    H = g["in:user_history"].shape[1]
    m = tt.TwoTowerWithUserHistoryEncoder(
        5, p["user_id_embedding_arch.weight"].shape[0], p["user_id_embedding_arch.weight"].shape[1],
        p["user_features_arch.0.weight"].shape[1], H, p["item_id_embedding_arch.weight"].shape[0],
        p["item_id_embedding_arch.weight"].shape[1], p["item_features_arch.0.weight"].shape[1], [1.0, 0.5],
        tt.BaselineMIPSModule(16, p["item_id_embedding_arch.weight"].shape[1]),
    )
    assert set(m.state_dict().keys()) == set(p.keys())
    m.load_state_dict(p, strict=True)
    assert torch.equal(m.user_history_encoder.positional_embeddings, g["attr:positional_embeddings"])
    g2 = load_golden("base_reftest.npz")
    p2 = section(g2, "p:")
    b = tt.TwoTowerBaseRetrieval(10, 100, 50, 20, 150, 40, 30, [0.1, 0.2, 0.3], tt.BaselineMIPSModule(16, 40))
    assert {k: tuple(v.shape) for k, v in b.state_dict().items()} == {k: tuple(v.shape) for k, v in p2.items()}
    for name in ("get_user_embedding", "process_user_features", "compute_user_embedding", "compute_item_embeddings",
                 "forward", "debias_net_user_value", "compute_training_loss", "train_forward"):
        assert callable(getattr(b, name))


@pytest.mark.parametrize("kind,cls_name", [("position", "TwoTowerWithPositionDebiasedWeights"),
                                           ("user", "TwoTowerWithUserDebiasedWeights"),
                                           ("both", "TwoTowerWithDebiasing")])
def test_debias_mirrors_keep_reference_names_and_hook_math(kind, cls_name):
    """The debias drop-ins carry the reference subclasses' parameter names (state_dict of the golden, produced by the
    unmodified reference, loads strictly) and their hook bodies - plain tensor code - reproduce the oracle's
    restatement of the reference hooks on CPU tensors."""
    import oracle
    import two_tower_models_b200 as tt

    g = load_golden(f"debias_{kind}.npz")
    p, batch = section(g, "p:"), section(g, "in:")
    DU, DI = p["user_id_embedding_arch.weight"].shape[1], p["item_id_embedding_arch.weight"].shape[1]
    m = getattr(tt, cls_name)(
        5, p["user_id_embedding_arch.weight"].shape[0], DU, p["user_features_arch.0.weight"].shape[1],
        batch["user_history"].shape[1], p["item_id_embedding_arch.weight"].shape[0], DI,
        p["item_features_arch.0.weight"].shape[1], g["attr:user_value_weights"].tolist(), tt.BaselineMIPSModule(16, DI),
    )
    assert set(m.state_dict().keys()) == set(p.keys())
    m.load_state_dict(p, strict=True)
    nuv = oracle.net_user_value(batch["labels"], g["attr:user_value_weights"])
    u = g["out:user_embedding"]
    got_w, got_loss = m.debias_net_user_value(net_user_value=nuv, position=batch["position"], user_embedding=u)
    ref_w, ref_loss = oracle.DEBIAS_HOOKS[kind](p, nuv, batch["position"], u)
    assert torch.allclose(got_w, ref_w, rtol=1e-6, atol=1e-7) and torch.allclose(got_loss, ref_loss, rtol=1e-6)


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only behaviour")
def test_fused_adam_and_graph_step_refuse_cpu_tensors():
    import two_tower_models_b200 as tt
    from two_tower_models_b200.graph import GraphedTrainStep

    w = torch.nn.Parameter(torch.randn(8, 4))
    w.grad = torch.randn(8, 4)
    opt = tt.FusedAdam([w], lr=1e-3)
    with pytest.raises(RuntimeError):
        opt.step()
    with pytest.raises(ValueError):
        tt.FusedAdam([w], lr=-1.0)
    m = tt.TwoTowerBaseRetrieval(4, 20, 16, 8, 20, 16, 8, [1.0], tt.BaselineMIPSModule(16, 16))
    batch = dict(user_id=torch.zeros(4, dtype=torch.int64), user_features=torch.randn(4, 8),
                 user_history=torch.zeros(4, 2, dtype=torch.int64), item_id=torch.zeros(4, dtype=torch.int64),
                 item_features=torch.randn(4, 8), position=torch.zeros(4, dtype=torch.int64), labels=torch.ones(4, 1))
    with pytest.raises(RuntimeError):
        GraphedTrainStep(m, batch)
