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
