"""GPU parity of UserHistoryEncoder / TwoTowerWithUserHistoryEncoder against the reference's known-answer
tests, the reference-generated golden vectors and the CPU oracle.

Tolerances: the kernels keep activations in bf16 between GEMMs (fp32 accumulate, fp32 softmax); vs the fp32
reference: outputs rel-Frobenius <= 2e-2 per attention layer stack, gradients <= 6e-2 (<= 1e-1 for the
12-row golden batch).  The known-answer tests (reference atol 1e-3, which the fp32 oracle meets in
tests/test_oracle_golden.py) are replayed here at atol 3e-2: their inputs (1..4 plus positional terms) are
rounded to bf16 (spacing 2^-6 .. 2^-7 at that magnitude) before the in-projection.
"""
import pytest
import torch

import oracle
from helpers import assert_close_fro, load_golden, rel_fro, section

pytestmark = pytest.mark.gpu


def _layers(p, prefix=""):
    out, i = [], 0
    while f"{prefix}multihead_attn_layers.{i}.in_proj_weight" in p:
        b = f"{prefix}multihead_attn_layers.{i}."
        out.append((p[b + "in_proj_weight"], p[b + "in_proj_bias"], p[b + "out_proj.weight"], p[b + "out_proj.bias"]))
        i += 1
    return out


def test_encoder_known_answer_vectors():
    """reference tests/test_user_history_enc.py:48-124 (seed 42, DI=2, H=3, 1 head, 1 layer)."""
    import two_tower_models_b200 as tt

    g = load_golden("encoder_kat.npz")
    expected = {
        False: torch.tensor([[[0.8240, 0.7119], [1.0, 2.0]]]),
        True: torch.tensor([[[1.4978, 1.2425], [1.0, 2.0]]]),
    }
    for use_pe in (False, True):
        torch.manual_seed(42)
        enc = tt.UserHistoryEncoder(2, 3, 1, 1, use_pe)
        tag = "pe" if use_pe else "nope"
        for k, v in enc.state_dict().items():  # same init stream as the reference under the same seed
            assert torch.equal(v, g[f"{tag}:p:{k}"]), k
        out = enc.cuda()(g["x"].cuda()).cpu()
        assert out.shape == (1, 2, 2)
        assert torch.allclose(out, expected[use_pe], atol=3e-2), out
        assert torch.allclose(out, g[f"{tag}:out"], atol=3e-2)
        assert torch.equal(out[:, 1], g["x"].mean(1))  # mean-pool is exact fp32, taken before the PE


def test_encoder_shape_contract():
    """reference tests/test_user_history_enc.py:21-46: D=64, H=128, 4 heads, 12 layers."""
    import two_tower_models_b200 as tt

    torch.manual_seed(42)
    enc = tt.UserHistoryEncoder(64, 128, 4, 12, True).cuda()
    x = torch.randn(32, 128, 64, device="cuda")
    y = enc(x)
    assert y.shape == (32, 2, 64) and bool(torch.isfinite(y).all())
    assert enc.get_output_dim() == 128


def test_encoder_l2_golden_forward_backward():
    import two_tower_models_b200 as tt

    g = load_golden("encoder_l2.npz")
    p = section(g, "p:")
    heads = int(g["attr:heads"])
    B, H, D = g["in:x"].shape
    enc = tt.UserHistoryEncoder(D, H, heads, len(_layers(p)), True)
    enc.load_state_dict(p, strict=True)
    assert torch.equal(enc.positional_embeddings, g["attr:pe"])
    enc = enc.cuda()
    x = g["in:x"].cuda().requires_grad_(True)
    y = enc(x)
    assert rel_fro(y, g["out:y"]) < 2e-2, rel_fro(y, g["out:y"])
    (y * g["in:gout"].cuda()).sum().backward()
    assert_close_fro(x.grad, g["grad:x"], rtol=6e-2, what="dx")
    for k, prm in enc.named_parameters():
        assert_close_fro(prm.grad, g["grad:" + k], rtol=6e-2, atol=1e-5, what=k)


@pytest.mark.parametrize("B,H,D,heads,L", [(64, 50, 128, 4, 2), (33, 17, 64, 2, 3), (20, 128, 64, 4, 1), (9, 50, 40, 5, 2)])
def test_encoder_matches_oracle(B, H, D, heads, L):
    import two_tower_models_b200 as tt

    torch.manual_seed(B)
    enc = tt.UserHistoryEncoder(D, H, heads, L, True)
    with torch.no_grad():  # non-zero biases so that they are exercised
        for layer in enc.multihead_attn_layers:
            layer.in_proj_bias.normal_(0, 0.1)
            layer.out_proj.bias.normal_(0, 0.1)
    p = {k: v.detach().clone() for k, v in enc.state_dict().items()}
    x = torch.randn(B, H, D)
    gout = torch.randn(B, 2, D)
    pr = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    xr = x.clone().requires_grad_(True)
    yr = oracle.history_encoder(xr, _layers(pr), heads, oracle.positional_encoding(H, D))
    (yr * gout).sum().backward()
    enc = enc.cuda()
    xg = x.cuda().requires_grad_(True)
    y = enc(xg)
    (y * gout.cuda()).sum().backward()
    assert rel_fro(y, yr) < 2e-2, rel_fro(y, yr)
    assert_close_fro(xg.grad, xr.grad, rtol=6e-2, what="dx")
    for k, prm in enc.named_parameters():
        assert_close_fro(prm.grad, pr[k].grad, rtol=6e-2, atol=1e-5, what=k)


def _build_hist(p, g, H, L=3, heads=4):
    import two_tower_models_b200 as tt

    DU = p["user_id_embedding_arch.weight"].shape[1]
    DI = p["item_id_embedding_arch.weight"].shape[1]
    mips = tt.BaselineMIPSModule(corpus_size=g["attr:corpus"].shape[0], embedding_dim=DI)
    mips.corpus = g["attr:corpus"].clone()
    m = tt.TwoTowerWithUserHistoryEncoder(
        num_items=5, user_id_hash_size=p["user_id_embedding_arch.weight"].shape[0], user_id_embedding_dim=DU,
        user_features_size=p["user_features_arch.0.weight"].shape[1], user_history_seqlen=H,
        item_id_hash_size=p["item_id_embedding_arch.weight"].shape[0], item_id_embedding_dim=DI,
        item_features_size=p["item_features_arch.0.weight"].shape[1],
        user_value_weights=g["attr:user_value_weights"].tolist(), mips_module=mips,
        num_attention_heads=heads, num_attention_layers=L,
    )
    m.load_state_dict(p, strict=True)
    return m.cuda()


def test_history_model_matches_reference_golden():
    """TwoTowerWithUserHistoryEncoder.train_forward + backward vs the reference's own outputs."""
    g = load_golden("hist_small.npz")
    p, batch, grads = section(g, "p:"), section(g, "in:"), section(g, "grad:")
    H = batch["user_history"].shape[1]
    m = _build_hist(p, g, H, L=3, heads=int(g["attr:heads"]))
    assert set(m.state_dict().keys()) == set(p.keys())
    assert torch.equal(m.user_history_encoder.positional_embeddings.cpu(), g["attr:positional_embeddings"])
    b = {k: v.cuda() for k, v in batch.items()}
    loss = m.train_forward(b["user_id"], b["user_features"], b["user_history"], b["item_id"], b["item_features"],
                           b["position"], b["labels"])
    loss.backward()
    u = m.compute_user_embedding(b["user_id"], b["user_features"], b["user_history"])
    assert rel_fro(u, g["out:user_embedding"]) < 2e-2
    assert abs(float(loss.detach()) - float(g["out:loss"])) <= 2e-3 * abs(float(g["out:loss"]))
    for k, prm in m.named_parameters():
        assert prm.grad is not None, k
        assert_close_fro(prm.grad, grads[k], rtol=1e-1, atol=1e-4 if prm.dim() == 1 else 2e-6, what=k)
    top = m(b["user_id"], b["user_features"], b["user_history"])
    assert top.shape == (batch["user_id"].shape[0], 5) and top.dtype == torch.int64


def test_history_model_reference_test_shapes():
    """reference tests/test_two_tower_user_hist.py:24-63 (H=128, DU=50?, ids < 10): shape / range contract."""
    import two_tower_models_b200 as tt

    torch.manual_seed(0)
    mips = tt.BaselineMIPSModule(corpus_size=1001, embedding_dim=40)
    m = tt.TwoTowerWithUserHistoryEncoder(
        num_items=10, user_id_hash_size=100, user_id_embedding_dim=50, user_features_size=20, user_history_seqlen=128,
        item_id_hash_size=200, item_id_embedding_dim=40, item_features_size=30, user_value_weights=[0.1, 0.2, 0.3],
        mips_module=mips,
    ).cuda()
    B = 32
    uid = torch.randint(0, 100, (B,), device="cuda")
    uf = torch.randn(B, 20, device="cuda")
    hist = torch.randint(0, 10, (B, 128), device="cuda")
    top = m(uid, uf, hist)
    assert top.shape == (B, 10) and int(top.min()) >= 0 and int(top.max()) < 1001
    loss = m.train_forward(uid, uf, hist, torch.randint(0, 200, (B,), device="cuda"), torch.randn(B, 30, device="cuda"),
                           torch.randint(0, 100, (B,), device="cuda"), torch.randint(0, 2, (B, 3), device="cuda").float())
    loss.backward()
    assert isinstance(loss.item(), float) and m.item_id_embedding_arch.weight.grad is not None


ATTN_CASES = [
    # nseq, H, D, heads, q_rows
    (5, 50, 128, 4, 50),    # BASELINE config 3 shape: two sequences per 128-row tile, head_dim 32
    (7, 50, 128, 4, 1),     # last encoder layer: only row 0
    (3, 128, 64, 4, 128),   # reference unit-test shape: one sequence per tile, head_dim 16
    (4, 17, 64, 2, 17),     # head_dim 32, odd tile tail
    (2, 64, 128, 2, 64),    # head_dim 64
    (301, 50, 128, 4, 50),  # more tiles than one wave of CTAs would leave idle; odd sequence count
    (3, 10, 40, 5, 10),     # head_dim 8: CUDA-core fallback kernel
]


@pytest.mark.parametrize("nseq,H,D,heads,q_rows", ATTN_CASES)
def test_attention_core_forward_backward(nseq, H, D, heads, q_rows):
    """tt_attn_fwd / tt_attn_bwd against plain fp32 torch on the same bf16-rounded q|k|v."""
    from two_tower_models_b200 import ops

    g = torch.Generator().manual_seed(nseq * 7 + H)
    hd = D // heads
    qkv = (torch.randn(nseq * H, 3 * D, generator=g) * 0.7).to(torch.bfloat16)
    dout = (torch.randn(nseq * q_rows, D, generator=g) * 0.5).to(torch.bfloat16)
    x = qkv.float().reshape(nseq, H, 3, heads, hd).requires_grad_(True)
    q, k, v = x[:, :, 0].transpose(1, 2), x[:, :, 1].transpose(1, 2), x[:, :, 2].transpose(1, 2)  # [n, h, H, hd]
    s = (q[:, :, :q_rows] @ k.transpose(-1, -2)) / (hd ** 0.5)
    ref = (torch.softmax(s, dim=-1) @ v).transpose(1, 2).reshape(nseq * q_rows, D)
    ref.backward(dout.float())
    dref = x.grad.reshape(nseq * H, 3 * D)
    dev = torch.device("cuda:0")
    pad = ops._r8(3 * D)
    qkv_d = torch.zeros(nseq * H, pad, dtype=torch.bfloat16, device=dev)
    qkv_d[:, :3 * D] = qkv.to(dev)
    out = ops.attn_forward(qkv_d, nseq, H, D, heads, q_rows)
    torch.cuda.synchronize()
    assert rel_fro(out[:, :D].float(), ref) < 1e-2, rel_fro(out[:, :D].float(), ref)
    do_d = torch.zeros(nseq * q_rows, ops._r8(D), dtype=torch.bfloat16, device=dev)
    do_d[:, :D] = dout.to(dev)
    dqkv = ops.attn_backward(qkv_d, do_d, nseq, H, D, heads, q_rows)
    torch.cuda.synchronize()
    assert rel_fro(dqkv[:, :3 * D].float(), dref) < 2e-2, rel_fro(dqkv[:, :3 * D].float(), dref)
