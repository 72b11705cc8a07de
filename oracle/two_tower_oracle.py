"""CPU restatement of the reference two-tower hot path (TEST INFRASTRUCTURE ONLY).

This file is the *oracle*: a functional, module-free restatement of the arithmetic
that gauravchak/two_tower_models performs on its hot path.  It is never imported by
the product package; it exists so that the CUDA kernels can be checked on machines
where the reference checkout (`/root/reference`) does not exist (the GPU box).

Where the arithmetic really lives
---------------------------------
The reference contains no numerical code of its own: every FLOP is executed by a
third-party dependency that is *not* vendored under /root/reference, namely
**PyTorch** (unpinned by the reference: its setup.py:1-7 declares no requirements;
the version in this image is torch 2.11.0+cu128, CPU path = ATen + MKL/oneDNN).
The call sites restated here are

* towers      src/two_tower_base_retrieval.py:112-219   (nn.Embedding, nn.Linear, ReLU, cat)
* loss        src/two_tower_base_retrieval.py:279-347   (matmul, F.cross_entropy, sum/clamp/max/mean)
* encoder     src/user_history_encoder.py:69-121        (mean, PE, nn.MultiheadAttention xL, stack)
                └ torch/nn/functional.py multi_head_attention_forward, need_weights branch:
                  packed in-projection, q scaled by 1/sqrt(head_dim) before QK^T, softmax over keys,
                  PV, head concat, out-projection (published algorithm of nn.MultiheadAttention)
* history tower src/two_tower_with_user_history_encoder.py:85-122
* MIPS        src/baseline_mips_module.py:57-72         (matmul + topk + gather)

The restatement uses torch *CPU tensor algebra only* (matmul / exp / sum ...), no
nn.Module, no F.cross_entropy, no nn.MultiheadAttention, so that it is an independent
statement of the algorithm rather than a second call into the same library routine.
Backward passes are given both in closed form (`inbatch_ce_backward`) and through
autograd over the functional forward (`base_train_forward_with_grads`).

Pinning status: **pinned**.
  * the reference's only two known-answer vectors (tests/test_user_history_enc.py:48-124)
    are replayed against `history_encoder` in tests/test_oracle_golden.py;
  * everything else is pinned against outputs of the reference itself, produced in the
    build container by tests/golden/gen_golden.py (imports /root/reference read-only) and
    committed as tests/golden/*.npz.
"""

from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch

Tensor = torch.Tensor

__all__ = [
    "feature_mlp",
    "tower_forward",
    "inbatch_ce",
    "inbatch_ce_backward",
    "net_user_value",
    "weighted_loss",
    "base_user_embedding",
    "base_item_embedding",
    "base_train_forward",
    "base_train_forward_with_grads",
    "positional_encoding",
    "mha_self_attention",
    "history_encoder",
    "history_user_embedding",
    "history_train_forward",
    "history_train_forward_with_grads",
    "DEBIAS_HOOKS",
    "debias_position",
    "debias_user",
    "debias_both",
    "debias_train_forward",
    "debias_train_forward_with_grads",
    "mips_topk",
    "mips_forward",
]


# --------------------------------------------------------------------------------------
# Towers  (src/two_tower_base_retrieval.py:112-219)
# --------------------------------------------------------------------------------------
def feature_mlp(x: Tensor, w0: Tensor, b0: Tensor, w1: Tensor, b1: Tensor) -> Tensor:
    """Linear(F,256) -> ReLU -> Linear(256,D); reference :76-80 / :101-105."""
    h = torch.clamp_min(x @ w0.t() + b0, 0.0)
    return h @ w1.t() + b1


def tower_forward(
    ids: Tensor,
    feats: Tensor,
    table: Tensor,
    w0: Tensor,
    b0: Tensor,
    w1: Tensor,
    b1: Tensor,
    wt: Tensor,
    bt: Tensor,
    extra: Optional[Tensor] = None,
) -> Tensor:
    """[id_emb | feature_mlp(feats) | extra] @ Wt^T + bt.

    Reference: gather :126/:209, MLP :153/:211, cat :159-161/:214-216, tower Linear :190/:218
    (no activation, no normalisation).  `extra` is the history summary appended by
    src/two_tower_with_user_history_encoder.py:121.
    """
    parts = [table[ids], feature_mlp(feats, w0, b0, w1, b1)]
    if extra is not None:
        parts.append(extra)
    x = torch.cat(parts, dim=1)
    return x @ wt.t() + bt


def _p(params: Dict[str, Tensor], prefix: str) -> Tuple[Tensor, ...]:
    return (
        params[f"{prefix}_id_embedding_arch.weight"],
        params[f"{prefix}_features_arch.0.weight"],
        params[f"{prefix}_features_arch.0.bias"],
        params[f"{prefix}_features_arch.2.weight"],
        params[f"{prefix}_features_arch.2.bias"],
        params[f"{prefix}_tower_arch.weight"],
        params[f"{prefix}_tower_arch.bias"],
    )


def base_user_embedding(params: Dict[str, Tensor], user_id: Tensor, user_features: Tensor) -> Tensor:
    """compute_user_embedding of the base class (:164-191); user_history is unused there."""
    return tower_forward(user_id, user_features, *_p(params, "user"))


def base_item_embedding(params: Dict[str, Tensor], item_id: Tensor, item_features: Tensor) -> Tensor:
    """compute_item_embeddings (:193-219)."""
    return tower_forward(item_id, item_features, *_p(params, "item"))


# --------------------------------------------------------------------------------------
# In-batch sampled-softmax loss  (src/two_tower_base_retrieval.py:279-347)
# --------------------------------------------------------------------------------------
def inbatch_ce(u: Tensor, v: Tensor, target_offset: int = 0) -> Tuple[Tensor, Tensor]:
    """Row-wise cross entropy of S = U V^T against the (shifted) diagonal.

    Reference :287 (scores), :301 (target = arange), :310-312 (cross_entropy, reduction none).
    Returns (ce[B], lse[B]) with ce_i = logsumexp_j S_ij - S_{i, i+target_offset}.
    `target_offset` is the multi-GPU generalisation (rank r scores its local users against
    the all-gathered items, the positives sit at column row + r*B_local).
    """
    s = u @ v.t()
    m = s.max(dim=1, keepdim=True).values
    lse = (m + torch.log(torch.exp(s - m).sum(dim=1, keepdim=True))).squeeze(1)
    rows = torch.arange(s.shape[0])
    diag = s[rows, rows + target_offset]
    return lse - diag, lse


def inbatch_ce_backward(
    u: Tensor, v: Tensor, lse: Tensor, g: Tensor, target_offset: int = 0
) -> Tuple[Tensor, Tensor]:
    """Closed-form backward of `inbatch_ce` for upstream g_i = dL/dce_i  (SURVEY §3.3).

    dS_ij = g_i (softmax(S)_ij - [j == i+off]);  dU = dS V;  dV = dS^T U.
    """
    s = u @ v.t()
    ds = torch.exp(s - lse[:, None])
    rows = torch.arange(s.shape[0])
    ds[rows, rows + target_offset] -= 1.0
    ds = ds * g[:, None]
    return ds @ v, ds.t() @ u


def net_user_value(labels: Tensor, user_value_weights: Tensor) -> Tensor:
    """sum_t labels[:, t] * w_t   (reference :322)."""
    return torch.sum(labels * user_value_weights, dim=-1)


def weighted_loss(ce: Tensor, nuv: Tensor, global_max: Optional[Tensor] = None,
                  global_rows: Optional[int] = None) -> Tensor:
    """clamp(1e-6) -> / batch max -> mean(ce * w)   (reference :334-343).

    `global_max` / `global_rows` let the multi-GPU tests evaluate a shard of the batch
    with the batch-global max and row count.
    """
    w = torch.clamp(nuv, min=0.000001)
    mx = torch.max(w) if global_max is None else global_max
    w = w / mx
    n = ce.shape[0] if global_rows is None else global_rows
    return torch.sum(ce * w) / n


def base_train_forward(params: Dict[str, Tensor], user_value_weights: Tensor, batch: Dict[str, Tensor]) -> Tensor:
    """TwoTowerBaseRetrieval.train_forward (:349-394) with the identity debias hook (:251-277)."""
    u = base_user_embedding(params, batch["user_id"], batch["user_features"])
    v = base_item_embedding(params, batch["item_id"], batch["item_features"])
    ce, _ = inbatch_ce(u, v)
    return weighted_loss(ce, net_user_value(batch["labels"], user_value_weights))


def _with_grads(fn, params: Dict[str, Tensor], *args):
    leaf = {k: p.detach().clone().requires_grad_(True) for k, p in params.items()}
    loss = fn(leaf, *args)
    loss.backward()
    grads = {k: (p.grad if p.grad is not None else torch.zeros_like(p)) for k, p in leaf.items()}
    return loss.detach(), grads


def base_train_forward_with_grads(params, user_value_weights, batch):
    """Loss and d loss / d every parameter (the reference's backward is autograd, train/train.py:124)."""
    return _with_grads(base_train_forward, params, user_value_weights, batch)


# --------------------------------------------------------------------------------------
# History encoder  (src/user_history_encoder.py)
# --------------------------------------------------------------------------------------
def positional_encoding(seq_len: int, d_model: int, flipped: bool = True) -> Tensor:
    """The repository's own (non-textbook) sinusoid table, :69-78, flipped along positions (:54).

    PE[p, i]   = sin(p / 10000^(2 i / d))        for even i
    PE[p, i+1] = cos(p / 10000^(2 (i+1) / d))    (note: a different frequency from the sin)
    Row order is reversed afterwards because history index 0 is the newest item.
    Evaluated in float64 like the reference's python `math` calls and rounded to fp32 on store.
    """
    pe = torch.zeros(seq_len, d_model, dtype=torch.float32)
    for pos in range(seq_len):
        for i in range(0, d_model, 2):
            pe[pos, i] = math.sin(pos / (10000 ** ((2 * i) / d_model)))
            if i + 1 < d_model:
                pe[pos, i + 1] = math.cos(pos / (10000 ** ((2 * (i + 1)) / d_model)))
    return pe.flip([0]) if flipped else pe


def mha_self_attention(
    x: Tensor, in_w: Tensor, in_b: Tensor, out_w: Tensor, out_b: Tensor, heads: int
) -> Tensor:
    """One bare nn.MultiheadAttention self-attention layer on x[B, H, D] (no residual/LN/FFN).

    Published algorithm of torch.nn.functional.multi_head_attention_forward (need_weights branch):
    qkv = x W_in^T + b_in; q *= head_dim^-0.5; A = softmax(q k^T); o = A v; concat heads; out-proj.
    """
    b, h, d = x.shape
    hd = d // heads
    qkv = x @ in_w.t() + in_b
    q, k, v = qkv[..., :d], qkv[..., d : 2 * d], qkv[..., 2 * d :]

    def split(t):
        return t.reshape(b, h, heads, hd).permute(0, 2, 1, 3)  # [B, heads, H, hd]

    q, k, v = split(q) * (1.0 / math.sqrt(hd)), split(k), split(v)
    s = q @ k.transpose(-1, -2)
    s = s - s.max(dim=-1, keepdim=True).values
    a = torch.exp(s)
    a = a / a.sum(dim=-1, keepdim=True)
    o = (a @ v).permute(0, 2, 1, 3).reshape(b, h, d)
    return o @ out_w.t() + out_b


def history_encoder(
    x: Tensor,
    layers: Sequence[Tuple[Tensor, Tensor, Tensor, Tensor]],
    heads: int,
    pe: Optional[Tensor],
) -> Tensor:
    """UserHistoryEncoder.forward (:80-121): [B,H,D] -> [B,2,D] = stack(attn row 0, mean-pool).

    Mean pooling happens BEFORE the positional encoding is added (:89 vs :95).
    """
    mean_pooled = x.mean(dim=1)
    if pe is not None:
        x = x + pe.unsqueeze(0)
    for (in_w, in_b, out_w, out_b) in layers:
        x = mha_self_attention(x, in_w, in_b, out_w, out_b, heads)
    return torch.stack([x[:, 0, :], mean_pooled], dim=1)


def _encoder_layers(params: Dict[str, Tensor], prefix: str = "user_history_encoder.") -> List[Tuple[Tensor, ...]]:
    layers = []
    i = 0
    while f"{prefix}multihead_attn_layers.{i}.in_proj_weight" in params:
        base = f"{prefix}multihead_attn_layers.{i}."
        layers.append(
            (
                params[base + "in_proj_weight"],
                params[base + "in_proj_bias"],
                params[base + "out_proj.weight"],
                params[base + "out_proj.bias"],
            )
        )
        i += 1
    return layers


def history_user_embedding(
    params: Dict[str, Tensor], user_id: Tensor, user_features: Tensor, user_history: Tensor,
    heads: int, pe: Optional[Tensor]
) -> Tensor:
    """TwoTowerWithUserHistoryEncoder.process_user_features + tower (:85-122).

    History ids are looked up in the ITEM id table (:105); concat order is
    [id_emb, feat_emb, most_recent, mean_pool].
    """
    hist = params["item_id_embedding_arch.weight"][user_history]  # [B,H,DI]
    summary = history_encoder(hist, _encoder_layers(params), heads, pe)
    summary = summary.reshape(summary.shape[0], -1)
    return tower_forward(user_id, user_features, *_p(params, "user"), extra=summary)


def history_train_forward(params, user_value_weights, batch, heads: int, pe: Optional[Tensor]) -> Tensor:
    u = history_user_embedding(params, batch["user_id"], batch["user_features"], batch["user_history"], heads, pe)
    v = base_item_embedding(params, batch["item_id"], batch["item_features"])
    ce, _ = inbatch_ce(u, v)
    return weighted_loss(ce, net_user_value(batch["labels"], user_value_weights))


def history_train_forward_with_grads(params, user_value_weights, batch, heads, pe):
    return _with_grads(history_train_forward, params, user_value_weights, batch, heads, pe)


# --------------------------------------------------------------------------------------
# debias_net_user_value overrides  
# --------------------------------------------------------------------------------------
def debias_position(params, nuv: Tensor, position: Tensor, u: Tensor) -> Tuple[Tensor, Tensor]:
    """src/two_tower_with_position_debiased_weights.py:76-113."""
    est = params["position_bias_net_user_value.weight"][position][:, 0]
    return nuv / torch.clamp(est, min=1e-3), torch.sum((est - nuv) ** 2)


def debias_user(params, nuv: Tensor, position: Tensor, u: Tensor) -> Tuple[Tensor, Tensor]:
    """src/two_tower_with_user_debiased_weights.py:100-135 (clamp at 1e-1 before the squared error)."""
    w, b = params["user_debias_net_user_value.0.weight"], params["user_debias_net_user_value.0.bias"]
    est = torch.clamp((u @ w.t() + b)[:, 0], min=1e-1)
    return nuv / est, torch.sum((est - nuv) ** 2)


def debias_both(params, nuv: Tensor, position: Tensor, u: Tensor) -> Tuple[Tensor, Tensor]:
    """src/two_tower_with_debiasing.py:77-129; the position term is a sum over all B x B (estimate_i, target_j)
    pairs because the reference's mse_loss broadcasts [B,1] against [B] (:110)."""
    pos = params["position_bias_net_user_value.weight"][position]  # [B,1]
    w, b = params["user_debias_net_user_value.0.weight"], params["user_debias_net_user_value.0.bias"]
    est = (torch.cat([u, pos], dim=-1) @ w.t() + b)[:, 0]
    loss = torch.sum((est - nuv) ** 2) + torch.sum((pos - nuv[None, :]) ** 2)
    return nuv / torch.clamp(est, min=1e-3), loss


DEBIAS_HOOKS = {"position": debias_position, "user": debias_user, "both": debias_both}


def debias_train_forward(params, user_value_weights, batch, heads: int, pe: Optional[Tensor], kind: str) -> Tensor:
    """train_forward of the history model with a debias hook between the label weights and the clamp (:322-346)."""
    u = history_user_embedding(params, batch["user_id"], batch["user_features"], batch["user_history"], heads, pe)
    v = base_item_embedding(params, batch["item_id"], batch["item_features"])
    ce, _ = inbatch_ce(u, v)
    nuv, extra = DEBIAS_HOOKS[kind](params, net_user_value(batch["labels"], user_value_weights), batch["position"], u)
    return weighted_loss(ce, nuv) + extra


def debias_train_forward_with_grads(params, user_value_weights, batch, heads, pe, kind):
    return _with_grads(debias_train_forward, params, user_value_weights, batch, heads, pe, kind)


# --------------------------------------------------------------------------------------
# MIPS  (src/baseline_mips_module.py:57-72)
# --------------------------------------------------------------------------------------
def mips_topk(query: Tensor, corpus: Tensor, k: int) -> Tuple[Tensor, Tensor]:
    """top-k of query @ corpus^T per row, sorted by (score desc, index asc).

    torch.topk leaves the order of exact ties unspecified; the oracle fixes it with a
    stable sort so that bit-exact index parity is well defined.  Returns (indices i64, scores).
    """
    s = query @ corpus.t()
    order = torch.sort(s, dim=1, descending=True, stable=True).indices[:, :k]
    return order, torch.gather(s, 1, order)


def mips_forward(query: Tensor, corpus: Tensor, k: int) -> Tuple[Tensor, Tensor, Tensor]:
    """BaselineMIPSModule.forward: (indices, scores, embeddings[Q,k,D]) in that order (:72)."""
    idx, sc = mips_topk(query, corpus, k)
    return idx, sc, corpus[idx]
