"""Pin the oracle: replay every golden vector produced by the reference (tests/golden/gen_golden.py)
and the reference's own two known-answer tests (reference tests/test_user_history_enc.py:48-124)."""
import pytest
import torch

import oracle
from helpers import assert_close_fro, load_golden, rel_fro, section


def _layers(p, prefix=""):
    out, i = [], 0
    while f"{prefix}multihead_attn_layers.{i}.in_proj_weight" in p:
        b = f"{prefix}multihead_attn_layers.{i}."
        out.append((p[b + "in_proj_weight"], p[b + "in_proj_bias"], p[b + "out_proj.weight"], p[b + "out_proj.bias"]))
        i += 1
    return out


def test_encoder_known_answer_vectors():
    g = load_golden("encoder_kat.npz")
    x = g["x"]
    # expected values are the literals of the reference's own tests, atol as in the reference (1e-3)
    expected = {
        "nope": torch.tensor([[[0.8240, 0.7119], [1.0, 2.0]]]),
        "pe": torch.tensor([[[1.4978, 1.2425], [1.0, 2.0]]]),
    }
    for tag in ("nope", "pe"):
        p = section(g, f"{tag}:p:")
        pe = oracle.positional_encoding(3, 2) if tag == "pe" else None
        if tag == "pe":
            assert torch.equal(pe, g["pe:pe"])
        y = oracle.history_encoder(x, _layers(p), heads=1, pe=pe)
        assert torch.allclose(y, expected[tag], atol=1e-3), (tag, y)
        assert torch.allclose(y, g[f"{tag}:out"], atol=1e-6)


def _check_train(name, history=False):
    g = load_golden(name)
    p, batch, grads = section(g, "p:"), section(g, "in:"), section(g, "grad:")
    uvw = g["attr:user_value_weights"]
    if history:
        heads = int(g["attr:heads"])
        pe = g["attr:positional_embeddings"]
        H, D = pe.shape
        assert torch.equal(oracle.positional_encoding(H, D), pe)
        u = oracle.history_user_embedding(p, batch["user_id"], batch["user_features"], batch["user_history"], heads, pe)
        loss, og = oracle.history_train_forward_with_grads(p, uvw, batch, heads, pe)
    else:
        u = oracle.base_user_embedding(p, batch["user_id"], batch["user_features"])
        loss, og = oracle.base_train_forward_with_grads(p, uvw, batch)
    v = oracle.base_item_embedding(p, batch["item_id"], batch["item_features"])
    assert rel_fro(u, g["out:user_embedding"]) < 2e-6
    assert rel_fro(v, g["out:item_embeddings"]) < 2e-6
    assert abs(float(loss) - float(g["out:loss"])) <= 2e-6 * abs(float(g["out:loss"])) + 1e-7
    assert set(og) == set(grads)
    for k in grads:
        assert_close_fro(og[k], grads[k], rtol=2e-5, atol=2e-8, what=k)
    # closed-form CE backward == autograd of the reference formulation
    ref_u, ref_v = g["out:user_embedding"], g["out:item_embeddings"]
    ce, lse = oracle.inbatch_ce(ref_u, ref_v)
    w = torch.clamp(oracle.net_user_value(batch["labels"], uvw), min=1e-6)
    w = w / w.max()
    gi = w / ce.shape[0]
    uu = ref_u.clone().requires_grad_(True)
    vv = ref_v.clone().requires_grad_(True)
    torch.nn.functional.cross_entropy(uu @ vv.t(), torch.arange(ce.shape[0]), reduction="none").mul(gi).sum().backward()
    du, dv = oracle.inbatch_ce_backward(ref_u, ref_v, lse, gi)
    assert rel_fro(du, uu.grad) < 1e-5 and rel_fro(dv, vv.grad) < 1e-5


def test_base_reference_test_shapes():
    _check_train("base_reftest.npz")


def test_base_config1_small():
    _check_train("base_c1small.npz")


def test_history_model():
    _check_train("hist_small.npz", history=True)


def test_encoder_l2_forward_backward():
    g = load_golden("encoder_l2.npz")
    p = {k: v.clone().requires_grad_(True) for k, v in section(g, "p:").items()}
    x = g["in:x"].clone().requires_grad_(True)
    y = oracle.history_encoder(x, _layers(p), int(g["attr:heads"]), g["attr:pe"])
    assert rel_fro(y, g["out:y"]) < 2e-6
    (y * g["in:gout"]).sum().backward()
    assert rel_fro(x.grad, g["grad:x"]) < 2e-5
    for k, v in p.items():
        assert_close_fro(v.grad, g["grad:" + k], rtol=2e-5, atol=2e-8, what=k)


def test_mips_randn_and_exact_grid():
    g = load_golden("mips.npz")
    idx, sc, emb = oracle.mips_forward(g["randn:q"], g["randn:corpus"], 10)
    assert torch.equal(idx, g["randn:idx"])  # no exact ties in randn data
    assert torch.allclose(sc, g["randn:scores"], rtol=1e-6, atol=1e-6)
    assert torch.equal(emb, g["randn:emb"])
    idx2, sc2 = oracle.mips_topk(g["grid:q"], g["grid:corpus"], 20)
    assert torch.equal(sc2, g["grid:scores"])  # exact-grid: scores are bit-identical in any order
    # ties may be ordered differently by torch.topk; the multiset of (score, idx) must agree
    ref_pairs = sorted(zip(g["grid:scores"].flatten().tolist(), g["grid:idx"].flatten().tolist()))
    got_pairs = sorted(zip(sc2.flatten().tolist(), idx2.flatten().tolist()))
    same = sum(a == b for a, b in zip(ref_pairs, got_pairs))
    assert same >= 0.98 * len(ref_pairs)
    # gathered scores are self-consistent
    full = g["grid:q"] @ g["grid:corpus"].t()
    assert torch.equal(torch.gather(full, 1, idx2), sc2)


@pytest.mark.parametrize("kind", ["position", "user", "both"])
def test_debias_hook_models(kind):
    """The three debias_net_user_value subclasses: oracle restatement of the hooks vs the reference's loss and every
    parameter gradient (tests/golden/debias_*.npz, produced by the unmodified reference)."""
    g = load_golden(f"debias_{kind}.npz")
    p, batch, grads = section(g, "p:"), section(g, "in:"), section(g, "grad:")
    uvw, heads, pe = g["attr:user_value_weights"], int(g["attr:heads"]), g["attr:positional_embeddings"]
    loss, og = oracle.debias_train_forward_with_grads(p, uvw, batch, heads, pe, kind)
    assert abs(float(loss) - float(g["out:loss"])) <= 5e-6 * abs(float(g["out:loss"])) + 1e-6
    assert set(og) == set(grads)
    for k in grads:
        assert_close_fro(og[k], grads[k], rtol=5e-5, atol=5e-8, what=k)
