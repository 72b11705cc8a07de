#!/usr/bin/env python
"""Generate golden vectors by running the UNMODIFIED reference (read-only import).

Run in the build container only (the GPU box has no /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/gen_golden.py

Each .npz holds the seeded inputs, the reference module's parameters (state_dict keys
verbatim, prefixed "p:") plus its non-state attributes, and the reference outputs
(prefixed "out:" / "grad:").  tests/test_oracle_golden.py replays them against oracle/.
Nothing from the reference's source is copied; only its numerical outputs are stored.
"""

import os
import sys

import numpy as np
import torch

REF = os.environ.get("TT_REFERENCE", "/root/reference")
sys.dont_write_bytecode = True
sys.path.insert(0, REF)

from src.baseline_mips_module import BaselineMIPSModule  # noqa: E402
from src.two_tower_base_retrieval import TwoTowerBaseRetrieval  # noqa: E402
from src.two_tower_with_user_history_encoder import TwoTowerWithUserHistoryEncoder  # noqa: E402
from src.user_history_encoder import UserHistoryEncoder  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def _np(t):
    return t.detach().cpu().numpy()


def _save(name, d):
    path = os.path.join(OUT, name)
    np.savez_compressed(path, **{k: (_np(v) if torch.is_tensor(v) else np.asarray(v)) for k, v in d.items()})
    print(f"wrote {path}  ({os.path.getsize(path)/1024:.1f} KiB)")


def _batch(g, B, IU, II, H, uhash, ihash, T, hist_high=None):
    return dict(
        user_id=torch.randint(0, uhash, (B,), generator=g),
        user_features=torch.randn(B, IU, generator=g),
        user_history=torch.randint(0, hist_high or ihash, (B, H), generator=g),
        item_id=torch.randint(0, ihash, (B,), generator=g),
        item_features=torch.randn(B, II, generator=g),
        position=torch.randint(0, 100, (B,), generator=g),
        labels=torch.randint(0, 2, (B, T), generator=g).float(),
    )


def _run_train(model, batch):
    model.zero_grad()
    loss = model.train_forward(
        batch["user_id"], batch["user_features"], batch["user_history"], batch["item_id"],
        batch["item_features"], batch["position"], batch["labels"],
    )
    loss.backward()
    u = model.compute_user_embedding(batch["user_id"], batch["user_features"], batch["user_history"])
    v = model.compute_item_embeddings(batch["item_id"], batch["item_features"])
    d = {f"in:{k}": t for k, t in batch.items()}
    d.update({f"p:{k}": t for k, t in model.state_dict().items()})
    d.update({f"grad:{k}": p.grad for k, p in model.named_parameters()})
    d["out:loss"] = loss
    d["out:user_embedding"] = u
    d["out:item_embeddings"] = v
    d["attr:user_value_weights"] = model.user_value_weights
    return d


def gen_base(name, seed, B, DU, DI, IU, II, uhash, ihash, uvw, H=8):
    torch.manual_seed(seed)
    mips = BaselineMIPSModule(corpus_size=257, embedding_dim=DI)
    model = TwoTowerBaseRetrieval(
        num_items=10, user_id_hash_size=uhash, user_id_embedding_dim=DU, user_features_size=IU,
        item_id_hash_size=ihash, item_id_embedding_dim=DI, item_features_size=II,
        user_value_weights=uvw, mips_module=mips,
    )
    g = torch.Generator().manual_seed(seed + 1000)
    batch = _batch(g, B, IU, II, H, uhash, ihash, len(uvw))
    d = _run_train(model, batch)
    top = model(batch["user_id"], batch["user_features"], batch["user_history"])
    d["attr:corpus"] = mips.corpus
    d["out:forward_top_items"] = top
    _save(name, d)


def gen_history(name, seed, B, DU, DI, IU, II, H, uhash, ihash, uvw):
    torch.manual_seed(seed)
    mips = BaselineMIPSModule(corpus_size=129, embedding_dim=DI)
    model = TwoTowerWithUserHistoryEncoder(
        num_items=5, user_id_hash_size=uhash, user_id_embedding_dim=DU, user_features_size=IU,
        user_history_seqlen=H, item_id_hash_size=ihash, item_id_embedding_dim=DI, item_features_size=II,
        user_value_weights=uvw, mips_module=mips,
    )
    g = torch.Generator().manual_seed(seed + 1000)
    batch = _batch(g, B, IU, II, H, uhash, ihash, len(uvw))
    d = _run_train(model, batch)
    d["attr:positional_embeddings"] = model.user_history_encoder.positional_embeddings
    d["attr:heads"] = model.user_history_encoder.num_attention_heads
    d["attr:corpus"] = mips.corpus
    _save(name, d)


def gen_debias(name, kind, seed, B, DU, DI, IU, II, H, uhash, ihash, uvw):
    """The three debias_net_user_value subclasses (position / user / both), loss and every gradient."""
    import warnings

    from src.two_tower_with_debiasing import TwoTowerWithDebiasing
    from src.two_tower_with_position_debiased_weights import TwoTowerWithPositionDebiasedWeights
    from src.two_tower_with_user_debiased_weights import TwoTowerWithUserDebiasedWeights

    cls = {"position": TwoTowerWithPositionDebiasedWeights, "user": TwoTowerWithUserDebiasedWeights,
           "both": TwoTowerWithDebiasing}[kind]
    torch.manual_seed(seed)
    mips = BaselineMIPSModule(corpus_size=129, embedding_dim=DI)
    model = cls(
        num_items=5, user_id_hash_size=uhash, user_id_embedding_dim=DU, user_features_size=IU,
        user_history_seqlen=H, item_id_hash_size=ihash, item_id_embedding_dim=DI, item_features_size=II,
        user_value_weights=uvw, mips_module=mips,
    )
    g = torch.Generator().manual_seed(seed + 1000)
    batch = _batch(g, B, IU, II, H, uhash, ihash, len(uvw))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")  # the reference's own broadcasting UserWarning (two_tower_with_debiasing.py:110)
        d = _run_train(model, batch)
    d["attr:positional_embeddings"] = model.user_history_encoder.positional_embeddings
    d["attr:heads"] = model.user_history_encoder.num_attention_heads
    d["attr:kind"] = np.array(kind)
    _save(name, d)


def gen_encoder_kats():
    """The reference's two known-answer tests (tests/test_user_history_enc.py:48-124), seed 42."""
    d = {}
    x = torch.tensor([[[1, 2], [3, 4], [-1, 0]]], dtype=torch.float32)
    for tag, use_pe in (("nope", False), ("pe", True)):
        torch.manual_seed(42)
        enc = UserHistoryEncoder(2, 3, 1, 1, use_pe)
        out = enc(x)
        for k, t in enc.state_dict().items():
            d[f"{tag}:p:{k}"] = t
        d[f"{tag}:out"] = out
        if use_pe:
            d[f"{tag}:pe"] = enc.positional_embeddings
    d["x"] = x
    _save("encoder_kat.npz", d)


def gen_encoder(name, seed, B, H, D, heads, L, use_pe=True):
    torch.manual_seed(seed)
    enc = UserHistoryEncoder(D, H, heads, L, use_pe)
    x = torch.randn(B, H, D).requires_grad_(True)
    out = enc(x)
    gout = torch.randn(out.shape, generator=torch.Generator().manual_seed(seed + 7))
    (out * gout).sum().backward()
    d = {f"p:{k}": t for k, t in enc.state_dict().items()}
    d.update({f"grad:{k}": p.grad for k, p in enc.named_parameters()})
    d["in:x"] = x
    d["in:gout"] = gout
    d["grad:x"] = x.grad
    d["out:y"] = out
    d["attr:heads"] = heads
    if use_pe:
        d["attr:pe"] = enc.positional_embeddings
    _save(name, d)


def gen_mips():
    torch.manual_seed(5)
    m = BaselineMIPSModule(corpus_size=1001, embedding_dim=40)
    q = torch.randn(32, 40)
    idx, sc, emb = m(q, 10)
    d = {"randn:corpus": m.corpus, "randn:q": q, "randn:idx": idx, "randn:scores": sc, "randn:emb": emb}
    # exact-grid inputs: k/64 with |k| <= 127 -> every partial sum exact in fp32
    g = torch.Generator().manual_seed(6)
    m2 = BaselineMIPSModule(corpus_size=777, embedding_dim=48)
    m2.corpus = torch.randint(-127, 128, (777, 48), generator=g).float() / 64.0
    q2 = torch.randint(-127, 128, (24, 48), generator=g).float() / 64.0
    idx2, sc2, _ = m2(q2, 20)
    d.update({"grid:corpus": m2.corpus, "grid:q": q2, "grid:idx": idx2, "grid:scores": sc2})
    _save("mips.npz", d)


if __name__ == "__main__":
    # shapes of the reference's own unit test (tests/test_two_tower_base_retrieval.py:10-38)
    gen_base("base_reftest.npz", 0, B=32, DU=50, DI=40, IU=20, II=30, uhash=100, ihash=150, uvw=[0.1, 0.2, 0.3])
    # a scaled-down config-1 (d=64, F=64), single task
    gen_base("base_c1small.npz", 1, B=96, DU=64, DI=64, IU=64, II=64, uhash=300, ihash=300, uvw=[1.0])
    gen_history("hist_small.npz", 2, B=12, DU=24, DI=32, IU=16, II=20, H=10, uhash=50, ihash=60, uvw=[1.0, 0.5])
    for i, kind in enumerate(("position", "user", "both")):
        gen_debias(f"debias_{kind}.npz", kind, 10 + i, B=16, DU=24, DI=32, IU=16, II=20, H=8, uhash=50, ihash=60,
                   uvw=[1.0, 0.5])
    gen_encoder_kats()
    gen_encoder("encoder_l2.npz", 3, B=6, H=20, D=64, heads=4, L=2)
    gen_mips()
