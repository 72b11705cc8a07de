"""Host-side mirror of the reference's `BaselineMIPSModule` (src/baseline_mips_module.py:10-72).

Same constructor arguments, attributes (`corpus_size`, `embedding_dim`, assignable `corpus`) and
`forward(query_embedding, num_items) -> (indices, scores, embeddings)` contract.  The Q x C^T scoring
and the per-row top-k run in one persistent tcgen05 kernel (csrc/mips.cu) that never materialises the
[Q, C] score matrix; the winners are re-scored in fp32 against the fp32 corpus so that the returned
scores and their order follow the reference's fp32 arithmetic.
"""
from typing import Tuple

import torch
import torch.nn as nn

from . import ops


class BaselineMIPSModule(nn.Module):
    def __init__(self, corpus_size: int, embedding_dim: int) -> None:
        super().__init__()
        self.corpus_size = corpus_size
        self.embedding_dim = embedding_dim
        # [C, DI] random corpus as in the reference (:30).  A non-persistent buffer so that .to(device)
        # moves it (the reference keeps a plain CPU tensor, which breaks its CUDA path) while
        # state_dict() stays identical to the reference's (empty).  `module.corpus = t` still works.
        self.register_buffer("corpus", torch.randn(corpus_size, embedding_dim), persistent=False)
        self._packed = ops.PackedWeights()  # bf16 screening copy of the corpus

    def forward(
        self,
        query_embedding: torch.Tensor,  # [B, DI]
        num_items: int,  # NI
    ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """Top `num_items` corpus rows by inner product (reference :32-72).

        Returns (indices int64 [B, NI], scores fp32 [B, NI] sorted descending, embeddings fp32 [B, NI, DI]).
        Ties are ordered by ascending corpus index (torch.topk leaves tie order unspecified).
        """
        corpus = self.corpus
        if query_embedding.dim() != 2 or query_embedding.shape[1] != corpus.shape[1]:
            raise RuntimeError(
                f"query_embedding must be [B, {corpus.shape[1]}], got {tuple(query_embedding.shape)}"
            )
        if not 0 < num_items <= corpus.shape[0]:
            raise RuntimeError(f"selected index k out of range (k={num_items}, corpus size {corpus.shape[0]})")
        corpus16 = self._packed.get("corpus", corpus)
        indices, scores = ops.mips_topk(query_embedding, corpus, corpus16, num_items)
        embeddings = ops.gather_rows_new(corpus, indices.reshape(-1)).reshape(
            indices.shape[0], num_items, corpus.shape[1]
        )
        return indices, scores, embeddings
