"""GPU parity of the fused MIPS top-k (tt_mips_topk) and of the BaselineMIPSModule drop-in.

Bit-exact bar: on exact-grid inputs (values k/64, |k| <= 127: every product and partial sum is exact in
fp32 in any order) indices and scores must equal the oracle's stable (score desc, index asc) order.
On randn inputs the reference's own top-k (golden vectors) is reproduced: same indices wherever
neighbouring fp32 scores differ by more than their rounding noise; scores to 1e-5 relative.
"""
import pytest
import torch

import oracle
from helpers import load_golden

pytestmark = pytest.mark.gpu


def _grid(shape, g):
    return torch.randint(-127, 128, shape, generator=g).float() / 64.0


def _run(q, corpus, k):
    import two_tower_models_b200 as tt

    m = tt.BaselineMIPSModule(corpus_size=corpus.shape[0], embedding_dim=corpus.shape[1])
    m.corpus = corpus.clone()
    m = m.cuda()
    idx, sc, emb = m(q.cuda(), k)
    torch.cuda.synchronize()
    return idx.cpu(), sc.cpu(), emb.cpu()


GRID_CASES = [
    # nq, nc, d, k
    (32, 100, 50, 10),       # reference tests/test_baseline_mips_module.py shape
    (32, 1001, 40, 10),      # reference tests/test_two_tower_base_retrieval.py shape (C=1001, DI=40)
    (1, 300, 64, 5),
    (257, 777, 48, 20),
    (300, 5000, 128, 100),
    (600, 40000, 128, 100),  # several corpus parts per query block, merged in the finalize kernel
    (40, 3, 16, 3),          # k == corpus size
    (700, 257, 128, 64),
    (300, 3000, 256, 50),    # BASELINE config-5 embedding width: one query tile per CTA, four K boxes
    (130, 1500, 200, 10),    # 128 < d < 256, d not a multiple of 64
]


@pytest.mark.parametrize("nq,nc,d,k", GRID_CASES)
def test_mips_exact_grid_bit_exact(nq, nc, d, k):
    g = torch.Generator().manual_seed(nq * 31 + nc)
    q, c = _grid((nq, d), g), _grid((nc, d), g)
    idx, sc, emb = _run(q, c, k)
    ridx, rsc = oracle.mips_topk(q, c, k)
    assert idx.dtype == torch.int64 and idx.shape == (nq, k) and sc.shape == (nq, k) and emb.shape == (nq, k, d)
    assert torch.equal(sc, rsc), f"scores differ: max |diff| {float((sc - rsc).abs().max())}"
    assert torch.equal(idx, ridx), f"{int((idx != ridx).sum())} of {idx.numel()} indices differ"
    assert torch.equal(emb, c[ridx])


def test_mips_heavy_ties():
    """Many duplicate corpus rows: exact ties must come back in ascending index order."""
    g = torch.Generator().manual_seed(7)
    base = _grid((37, 64), g)
    c = base[torch.randint(0, 37, (3000,), generator=g)]
    q = _grid((130, 64), g)
    idx, sc, _ = _run(q, c, 50)
    ridx, rsc = oracle.mips_topk(q, c, 50)
    assert torch.equal(sc, rsc) and torch.equal(idx, ridx)


def test_mips_reference_golden_randn():
    g = load_golden("mips.npz")
    idx, sc, emb = _run(g["randn:q"], g["randn:corpus"], 10)
    assert torch.allclose(sc, g["randn:scores"], rtol=1e-5, atol=1e-5)
    same = (idx == g["randn:idx"])
    # a differing index is only acceptable where the two candidates' fp32 scores are within rounding noise
    if not bool(same.all()):
        full = g["randn:q"] @ g["randn:corpus"].t()
        a = torch.gather(full, 1, idx)
        b = torch.gather(full, 1, g["randn:idx"])
        assert float((a - b).abs()[~same].max()) < 1e-4
    assert torch.equal(emb, g["randn:corpus"][idx])
    idx2, sc2, _ = _run(g["grid:q"], g["grid:corpus"], 20)
    assert torch.equal(sc2, g["grid:scores"])  # the reference's own scores on exact-grid inputs, bit for bit


def test_mips_randn_recall_large():
    """randn corpus (bf16 screening is inexact here): the fp32 re-scored top-k must contain the fp32
    oracle's top-k except for candidates closer than the screening margin can separate."""
    g = torch.Generator().manual_seed(3)
    nq, nc, d, k = 512, 100_000, 128, 100
    q, c = torch.randn(nq, d, generator=g), torch.randn(nc, d, generator=g)
    idx, sc, _ = _run(q, c, k)
    ridx, rsc = oracle.mips_topk(q, c, k)
    hit = sum(len(set(a.tolist()) & set(b.tolist())) for a, b in zip(idx, ridx)) / (nq * k)
    assert hit >= 0.9999, hit
    assert torch.allclose(sc, rsc, rtol=2e-5, atol=2e-5)
    assert bool((sc[:, :-1] >= sc[:, 1:]).all())


def test_mips_module_contract_and_errors():
    import two_tower_models_b200 as tt

    m = tt.BaselineMIPSModule(corpus_size=100, embedding_dim=50)
    assert m.corpus.shape == (100, 50) and m.corpus_size == 100 and m.embedding_dim == 50
    assert list(m.state_dict().keys()) == []  # corpus is not part of the reference's state_dict either
    m = m.cuda()
    q = torch.randn(32, 50, device="cuda")
    idx, sc, emb = m(q, 10)  # (ids, scores, embeddings): reference order
    assert idx.shape == (32, 10) and sc.shape == (32, 10) and emb.shape == (32, 10, 50)
    assert int(idx.min()) >= 0 and int(idx.max()) < 100
    with pytest.raises(RuntimeError):
        m(q, 101)  # k > corpus size, as torch.topk
    with pytest.raises(RuntimeError):
        m(torch.randn(4, 7, device="cuda"), 3)


def test_mips_full_size_properties():
    """BASELINE config-4 corpus (1M x 128) with a query slice: sortedness, self-consistency of the returned
    scores with the returned indices, and agreement with the oracle on a sub-sample of the queries."""
    g = torch.Generator().manual_seed(11)
    nq, nc, d, k = 2048, 1_000_000, 128, 100
    q, c = _grid((nq, d), g), _grid((nc, d), g)
    idx, sc, _ = _run(q, c, k)
    assert bool((sc[:, :-1] >= sc[:, 1:]).all())
    assert torch.equal((q[:, None, :] * c[idx]).sum(-1), sc)  # exact-grid: any summation order is exact
    sub = torch.arange(0, nq, 64)
    ridx, rsc = oracle.mips_topk(q[sub], c, k)
    assert torch.equal(sc[sub], rsc) and torch.equal(idx[sub], ridx)


def test_mips_config4_all_queries_sampled_oracle():
    """BASELINE config 4 at its full size (65 536 queries x 1M corpus, d = 128, top-100) - the launch shape with one
    full round of query blocks plus a tail whose corpus is cut into parts and merged by the finalize kernel.  Every
    64th query is compared with the oracle: bit-exact indices and scores on exact-grid data (ties broken by index)."""
    g = torch.Generator().manual_seed(21)
    nq, nc, d, k = 65536, 1_000_000, 128, 100
    q, c = _grid((nq, d), g), _grid((nc, d), g)
    import two_tower_models_b200 as tt

    m = tt.BaselineMIPSModule(corpus_size=nc, embedding_dim=d)
    m.corpus = c.clone()
    m = m.cuda()
    idx, sc, emb = m(q.cuda(), k)
    del emb  # [65536, 100, 128] fp32: not needed on the host
    idx, sc = idx.cpu(), sc.cpu()
    assert idx.shape == (nq, k) and int(idx.min()) >= 0 and int(idx.max()) < nc
    assert bool((sc[:, :-1] >= sc[:, 1:]).all())
    sub = torch.arange(0, nq, 64)
    for lo in range(0, len(sub), 256):  # 256 oracle queries at a time: 1 GB of fp32 scores on the host
        s = sub[lo:lo + 256]
        ridx, rsc = oracle.mips_topk(q[s], c, k)
        assert torch.equal(sc[s], rsc) and torch.equal(idx[s], ridx), lo
