"""GPU parity of the debias_net_user_value drop-ins against the reference's own outputs.

The towers, the history encoder and the B x B loss run in the sm_100a kernels; the [B]-sized hook bodies are torch
code on the device.  Tolerances as in test_gpu_history.py (bf16 tensor-core operands vs the fp32 reference, tiny
batch): loss rel 2e-3, gradients rel-Frobenius 1e-1 with an absolute floor for analytically small bias gradients.
"""
import pytest
import torch

from helpers import assert_close_fro, load_golden, rel_fro, section

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind,cls_name", [("position", "TwoTowerWithPositionDebiasedWeights"),
                                           ("user", "TwoTowerWithUserDebiasedWeights"),
                                           ("both", "TwoTowerWithDebiasing")])
def test_debias_models_match_reference_golden(kind, cls_name):
    import two_tower_models_b200 as tt

    g = load_golden(f"debias_{kind}.npz")
    p, batch, grads = section(g, "p:"), section(g, "in:"), section(g, "grad:")
    DU = p["user_id_embedding_arch.weight"].shape[1]
    DI = p["item_id_embedding_arch.weight"].shape[1]
    mips = tt.BaselineMIPSModule(corpus_size=129, embedding_dim=DI)
    m = getattr(tt, cls_name)(
        num_items=5, user_id_hash_size=p["user_id_embedding_arch.weight"].shape[0], user_id_embedding_dim=DU,
        user_features_size=p["user_features_arch.0.weight"].shape[1], user_history_seqlen=batch["user_history"].shape[1],
        item_id_hash_size=p["item_id_embedding_arch.weight"].shape[0], item_id_embedding_dim=DI,
        item_features_size=p["item_features_arch.0.weight"].shape[1],
        user_value_weights=g["attr:user_value_weights"].tolist(), mips_module=mips,
    )
    assert set(m.state_dict().keys()) == set(p.keys())  # same parameter names as the reference subclass
    m.load_state_dict(p, strict=True)
    m = m.cuda()
    b = {k: v.cuda() for k, v in batch.items()}
    loss = m.train_forward(b["user_id"], b["user_features"], b["user_history"], b["item_id"], b["item_features"],
                           b["position"], b["labels"])
    loss.backward()
    assert abs(float(loss.detach()) - float(g["out:loss"])) <= 2e-3 * abs(float(g["out:loss"]))
    u = m.compute_user_embedding(b["user_id"], b["user_features"], b["user_history"])
    assert rel_fro(u, g["out:user_embedding"]) < 2e-2
    for k, prm in m.named_parameters():
        assert prm.grad is not None, k
        assert_close_fro(prm.grad, grads[k], rtol=1e-1, atol=1e-3 if prm.dim() == 1 else 1e-5, what=k)
    top = m(b["user_id"], b["user_features"], b["user_history"])
    assert top.shape == (batch["user_id"].shape[0], 5) and top.dtype == torch.int64
