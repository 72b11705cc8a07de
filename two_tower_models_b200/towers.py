"""Host-side mirror of the reference's `TwoTowerBaseRetrieval` (src/two_tower_base_retrieval.py:25-394).

Same constructor arguments, method names/signatures, attributes and state_dict keys
(`load_state_dict(reference.state_dict())` is the parity bridge).  The sub-modules below only HOST
the fp32 master parameters (so names and default initialisation match the reference); their
`forward` is never called - all arithmetic runs in libtt_b200.so through `ops`.
"""
from typing import List, Optional, Tuple

import torch
import torch.nn as nn

import os

from . import ops

_OVERLAP_TOWERS = os.environ.get("TT_B200_OVERLAP_TOWERS", "1") == "1"
_BATCH_TOWERS = os.environ.get("TT_B200_BATCH_TOWERS", "1") == "1"


class TwoTowerBaseRetrieval(nn.Module):
    """Two-tower candidate retrieval: user tower, item tower, in-batch sampled-softmax loss, MIPS inference."""

    def __init__(
        self,
        num_items: int,
        user_id_hash_size: int,
        user_id_embedding_dim: int,
        user_features_size: int,
        item_id_hash_size: int,
        item_id_embedding_dim: int,
        item_features_size: int,
        user_value_weights: List[float],
        mips_module: nn.Module,
    ) -> None:
        super().__init__()
        self.num_items = num_items
        # [T]; a (non-persistent) buffer so that .to(device) moves it - the reference keeps a plain CPU
        # tensor here (:62) which is why its CUDA path does not run as shipped.  Not part of state_dict.
        self.register_buffer("user_value_weights", torch.tensor(user_value_weights), persistent=False)
        self.mips_module = mips_module

        # parameter hosts; layout identical to the reference (:70-110)
        self.user_id_embedding_arch = nn.Embedding(user_id_hash_size, user_id_embedding_dim)
        self.user_features_arch = nn.Sequential(
            nn.Linear(user_features_size, 256), nn.ReLU(), nn.Linear(256, user_id_embedding_dim)
        )
        self.user_tower_arch = nn.Linear(2 * user_id_embedding_dim, item_id_embedding_dim)
        self.item_id_embedding_arch = nn.Embedding(item_id_hash_size, item_id_embedding_dim)
        self.item_features_arch = nn.Sequential(
            nn.Linear(item_features_size, 256), nn.ReLU(), nn.Linear(256, item_id_embedding_dim)
        )
        self.item_tower_arch = nn.Linear(2 * item_id_embedding_dim, item_id_embedding_dim)

        self._packed = ops.PackedWeights()  # bf16 operand copies of the weights
        self._dp = None  # optional data-parallel context, see distributed.enable_data_parallel
        self._streams = {}  # device -> side stream for the item tower

    def _row_exchange(self, table: torch.Tensor):
        """Data parallel only: callable that all-gathers the touched embedding rows' gradients (see
        distributed.DataParallelContext.row_exchange); None on a single GPU."""
        return None if self._dp is None else self._dp.row_exchange(self, table)

    def _side_stream(self, device) -> "torch.cuda.Stream":
        st = self._streams.get(device)
        if st is None:
            st = torch.cuda.Stream(device=device)
            self._streams[device] = st
        return st

    # ------------------------------------------------------------------ user tower
    def get_user_embedding(self, user_id: torch.Tensor, user_features: torch.Tensor) -> torch.Tensor:
        """Embedding-table lookup [B] -> [B, DU]; `user_features` is unused (reference :112-127)."""
        return ops.EmbeddingFunction.apply(user_id, self.user_id_embedding_arch.weight)

    def _user_feature_mlp(self, user_features: torch.Tensor) -> torch.Tensor:
        fa = self.user_features_arch
        return ops.FeatureMLPFunction.apply(
            user_features, fa[0].weight, fa[0].bias, fa[2].weight, fa[2].bias, self._packed, "user_features_arch"
        )

    def process_user_features(
        self, user_id: torch.Tensor, user_features: torch.Tensor, user_history: torch.Tensor
    ) -> torch.Tensor:
        """[B, 2*DU] = cat(id embedding, feature MLP); `user_history` unused here (reference :129-162)."""
        user_id_embedding = self.get_user_embedding(user_id=user_id, user_features=user_features)
        user_features_embedding = self._user_feature_mlp(user_features)
        return torch.cat([user_id_embedding, user_features_embedding], dim=1)

    def _user_tower_extra(self, user_history: torch.Tensor) -> Optional[torch.Tensor]:
        """Extra input block appended after [id_emb, feat_emb]; None in the base class."""
        return None

    def _fused_user_tower_ok(self) -> bool:
        """The fused tower kernel path is valid unless a subclass replaced the virtual pieces."""
        cls = type(self)
        known = getattr(cls, "_tt_fused_process_user_features", TwoTowerBaseRetrieval.process_user_features)
        return (
            cls.process_user_features is known
            and cls.get_user_embedding is TwoTowerBaseRetrieval.get_user_embedding
        )

    def compute_user_embedding(
        self, user_id: torch.Tensor, user_features: torch.Tensor, user_history: torch.Tensor
    ) -> torch.Tensor:
        """[B, DI] query embedding (reference :164-191; no activation or normalisation after the Linear)."""
        if self._fused_user_tower_ok():
            fa = self.user_features_arch
            return ops.TowerFunction.apply(
                user_id, user_features, self._user_tower_extra(user_history),
                self.user_id_embedding_arch.weight, fa[0].weight, fa[0].bias, fa[2].weight, fa[2].bias,
                self.user_tower_arch.weight, self.user_tower_arch.bias, self._packed, "user",
                self._row_exchange(self.user_id_embedding_arch.weight),
            )
        # a subclass overrode process_user_features / get_user_embedding: honour it, keep the Linear on-device
        user_tower_input = self.process_user_features(
            user_id=user_id, user_features=user_features, user_history=user_history
        )
        return ops.LinearFunction.apply(
            user_tower_input, self.user_tower_arch.weight, self.user_tower_arch.bias, self._packed, "user_tower_arch"
        )

    # ------------------------------------------------------------------ item tower
    def compute_item_embeddings(self, item_id: torch.Tensor, item_features: torch.Tensor) -> torch.Tensor:
        """[B, DI] item embeddings (reference :193-219)."""
        fa = self.item_features_arch
        return ops.TowerFunction.apply(
            item_id, item_features, None,
            self.item_id_embedding_arch.weight, fa[0].weight, fa[0].bias, fa[2].weight, fa[2].bias,
            self.item_tower_arch.weight, self.item_tower_arch.bias, self._packed, "item",
            self._row_exchange(self.item_id_embedding_arch.weight),
        )

    # ------------------------------------------------------------------ inference
    def forward(self, user_id: torch.Tensor, user_features: torch.Tensor, user_history: torch.Tensor) -> torch.Tensor:
        """Top `num_items` corpus indices per user, [B, num_items] int64 (reference :221-249)."""
        user_embedding = self.compute_user_embedding(user_id, user_features, user_history)
        top_items, _, _ = self.mips_module(query_embedding=user_embedding, num_items=self.num_items)
        return top_items

    # ------------------------------------------------------------------ training loss
    def debias_net_user_value(
        self, net_user_value: torch.Tensor, position: torch.Tensor, user_embedding: torch.Tensor
    ) -> Tuple[torch.Tensor, torch.Tensor]:
        """Virtual hook (reference :251-277): identity in the base class, additional loss 0."""
        return net_user_value, 0

    def compute_training_loss(
        self,
        user_embedding: torch.Tensor,  # [B, DI]
        item_embeddings: torch.Tensor,  # [B, DI]
        position: torch.Tensor,  # [B]
        labels: torch.Tensor,  # [B, T]
    ) -> torch.Tensor:
        """In-batch sampled-softmax loss weighted by net user value (reference :279-347).

        The B x B score matrix, its row softmax and the cross entropy against the diagonal run in one
        fused kernel (`ops.inbatch_cross_entropy`), per-row `ce[B]` comes back so that the [B]-sized
        weighting and the `debias_net_user_value` hook stay ordinary differentiable tensor code.
        """
        if self._dp is not None:
            return self._dp.compute_training_loss(self, user_embedding, item_embeddings, position, labels)
        if (
            type(self).debias_net_user_value is TwoTowerBaseRetrieval.debias_net_user_value
            and labels.dim() == 2
            and labels.shape[1] == self.user_value_weights.shape[0]
            and not labels.requires_grad
        ):
            # identity hook: label weights, batch max and the weighted mean ride in the CE's merge kernel
            return ops.InBatchWeightedLossFunction.apply(user_embedding, item_embeddings, labels, self.user_value_weights)
        loss = ops.inbatch_cross_entropy(user_embedding, item_embeddings)  # [B]
        net_user_value = torch.sum(labels * self.user_value_weights, dim=-1)  # [B]
        net_user_value, additional_loss = self.debias_net_user_value(
            net_user_value=net_user_value, position=position, user_embedding=user_embedding
        )
        net_user_value = torch.clamp(net_user_value, min=0.000001)
        net_user_value = net_user_value / torch.max(net_user_value)
        loss = torch.mean(loss * net_user_value)
        return loss + additional_loss

    def train_forward(
        self,
        user_id: torch.Tensor,  # [B]
        user_features: torch.Tensor,  # [B, IU]
        user_history: torch.Tensor,  # [B, H]
        item_id: torch.Tensor,  # [B]
        item_features: torch.Tensor,  # [B, II]
        position: torch.Tensor,  # [B]
        labels: torch.Tensor,  # [B, T]
    ) -> torch.Tensor:
        """Training loss, a 0-dim fp32 tensor with grad_fn (reference :349-394).

        The two towers are independent until the loss, so the item tower is enqueued on a second CUDA
        stream (forked from / joined to the caller's stream; autograd replays the same split in backward):
        each tower is a chain of small launch-latency-bound kernels that fills half of the SMs at most.
        """
        cls = type(self)
        base_towers = (cls.compute_user_embedding is TwoTowerBaseRetrieval.compute_user_embedding
                       and cls.compute_item_embeddings is TwoTowerBaseRetrieval.compute_item_embeddings)
        if _BATCH_TOWERS and item_id.is_cuda and self._fused_user_tower_ok() and base_towers:
            # (a subclass that overrides compute_user_embedding / compute_item_embeddings is dispatched through its
            # methods below, exactly as the reference's train_forward :380-386 does)
            # both towers advance in lock step: every stage is one launch covering the user and the item side
            fu, fi = self.user_features_arch, self.item_features_arch
            user_embedding, item_embeddings = ops.TowerSetFunction.apply(
                [("user", self._row_exchange(self.user_id_embedding_arch.weight)),
                 ("item", self._row_exchange(self.item_id_embedding_arch.weight))],
                self._packed,
                user_id, user_features, self._user_tower_extra(user_history), self.user_id_embedding_arch.weight,
                fu[0].weight, fu[0].bias, fu[2].weight, fu[2].bias, self.user_tower_arch.weight, self.user_tower_arch.bias,
                item_id, item_features, None, self.item_id_embedding_arch.weight,
                fi[0].weight, fi[0].bias, fi[2].weight, fi[2].bias, self.item_tower_arch.weight, self.item_tower_arch.bias,
            )
        elif _OVERLAP_TOWERS and item_id.is_cuda:
            cur = torch.cuda.current_stream(item_id.device)
            side = self._side_stream(item_id.device)
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                item_embeddings = self.compute_item_embeddings(item_id, item_features)
            user_embedding = self.compute_user_embedding(user_id, user_features, user_history)
            cur.wait_stream(side)
            item_embeddings.record_stream(cur)
            shadow = ops.shadow_of(item_embeddings)
            if shadow is not None:
                shadow.record_stream(cur)
        else:
            user_embedding = self.compute_user_embedding(user_id, user_features, user_history)
            item_embeddings = self.compute_item_embeddings(item_id, item_features)
        loss = self.compute_training_loss(
            user_embedding=user_embedding, item_embeddings=item_embeddings, position=position, labels=labels
        )
        ops.join_pending_fills()  # dense table-gradient zero fills ran beside the scoring kernels
        return loss


class TwoTowerWithUserHistoryEncoder(TwoTowerBaseRetrieval):
    """Base model + `UserHistoryEncoder` over the user's item history
    (reference src/two_tower_with_user_history_encoder.py:14-122).

    History ids are looked up in the ITEM id table (:105), summarised to [B, 2*DI] and appended to
    the user tower input, whose Linear therefore takes 2*DU + 2*DI inputs (:81-83).  The reference
    hard-codes 4 heads / 3 layers (:64-70); they are exposed here as optional trailing kwargs with
    those defaults (BASELINE config 3 uses 2 layers).
    """

    def __init__(
        self,
        num_items: int,
        user_id_hash_size: int,
        user_id_embedding_dim: int,
        user_features_size: int,
        user_history_seqlen: int,
        item_id_hash_size: int,
        item_id_embedding_dim: int,
        item_features_size: int,
        user_value_weights: List[float],
        mips_module: nn.Module,
        num_attention_heads: int = 4,
        num_attention_layers: int = 3,
    ) -> None:
        super().__init__(
            num_items=num_items,
            user_id_hash_size=user_id_hash_size,
            user_id_embedding_dim=user_id_embedding_dim,
            user_features_size=user_features_size,
            item_id_hash_size=item_id_hash_size,
            item_id_embedding_dim=item_id_embedding_dim,
            item_features_size=item_features_size,
            user_value_weights=user_value_weights,
            mips_module=mips_module,
        )
        from .history import UserHistoryEncoder

        self.user_history_encoder = UserHistoryEncoder(
            item_id_embedding_dim=item_id_embedding_dim,
            history_len=user_history_seqlen,
            num_attention_heads=num_attention_heads,
            num_attention_layers=num_attention_layers,
            use_positional_encoding=True,
        )
        self.user_tower_arch = nn.Linear(
            2 * user_id_embedding_dim + self.user_history_encoder.get_output_dim(), item_id_embedding_dim
        )

    def _user_tower_extra(self, user_history: torch.Tensor) -> Optional[torch.Tensor]:
        """[B, 2*DI] = [attention output of the newest item | mean-pooled history] (gather fused in)."""
        return self.user_history_encoder.encode_ids(self.item_id_embedding_arch.weight, user_history)

    def process_user_features(
        self, user_id: torch.Tensor, user_features: torch.Tensor, user_history: torch.Tensor
    ) -> torch.Tensor:
        """[B, 2*DU + 2*DI] = cat(id emb, feature MLP, most-recent attention row, mean-pool) (reference :85-122)."""
        user_tower_input = super().process_user_features(
            user_id=user_id, user_features=user_features, user_history=user_history
        )
        return torch.cat([user_tower_input, self._user_tower_extra(user_history)], dim=1)


TwoTowerWithUserHistoryEncoder._tt_fused_process_user_features = TwoTowerWithUserHistoryEncoder.process_user_features
