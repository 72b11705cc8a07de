"""Drop-ins for the reference's three `debias_net_user_value` subclasses:

    TwoTowerWithPositionDebiasedWeights   src/two_tower_with_position_debiased_weights.py:15-113
    TwoTowerWithUserDebiasedWeights       src/two_tower_with_user_debiased_weights.py:17-135
    TwoTowerWithDebiasing                 src/two_tower_with_debiasing.py:15-129

Same constructors and parameter names as the reference (`position_bias_net_user_value.weight`,
`user_debias_net_user_value.0.{weight,bias}`).  They only override the virtual hook: towers, history encoder and
the B x B loss still run in the sm_100a kernels (the fused loss op hands back per-row `ce[B]` and accepts the
gradient that flows back through the returned weights and through `user_embedding`).  The hook bodies are
[B]-sized tensor algebra and stay ordinary differentiable PyTorch code on the device, as the boundary contract
 requires.
"""
from typing import List, Tuple

import torch
import torch.nn as nn

from .towers import TwoTowerWithUserHistoryEncoder

_POSITION_VOCAB = 100  # hard-coded in the reference (position debias :72-74, combined :69-71)


def _sum_squared_error(estimate: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """F.mse_loss(input=estimate, target=target, reduction='sum') including its broadcasting."""
    return torch.sum((estimate - target) ** 2)


class TwoTowerWithPositionDebiasedWeights(TwoTowerWithUserHistoryEncoder):
    """net_user_value divided by a per-position estimate of itself (an `nn.Embedding(100, 1)` trained by a summed
    squared error against the undebiased value)."""

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
    ) -> None:
        super().__init__(
            num_items=num_items, user_id_hash_size=user_id_hash_size, user_id_embedding_dim=user_id_embedding_dim,
            user_features_size=user_features_size, user_history_seqlen=user_history_seqlen,
            item_id_hash_size=item_id_hash_size, item_id_embedding_dim=item_id_embedding_dim,
            item_features_size=item_features_size, user_value_weights=user_value_weights, mips_module=mips_module,
        )
        self.position_bias_net_user_value = nn.Embedding(num_embeddings=_POSITION_VOCAB, embedding_dim=1)

    def debias_net_user_value(
        self, net_user_value: torch.Tensor, position: torch.Tensor, user_embedding: torch.Tensor
    ) -> Tuple[torch.Tensor, torch.Tensor]:
        """(:76-113) estimate[b] = table[position[b]]; loss = sum (estimate - nuv)^2; nuv / clamp(estimate, 1e-3)."""
        estimate = self.position_bias_net_user_value(position).squeeze(1)  # [B]
        estimate_loss = _sum_squared_error(estimate, net_user_value)
        return net_user_value / torch.clamp(estimate, min=1e-3), estimate_loss


class TwoTowerWithUserDebiasedWeights(TwoTowerWithUserHistoryEncoder):
    """net_user_value divided by an estimate of itself from the user embedding (`Linear(DI, 1)`)."""

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
    ) -> None:
        super().__init__(
            num_items=num_items, user_id_hash_size=user_id_hash_size, user_id_embedding_dim=user_id_embedding_dim,
            user_features_size=user_features_size, user_history_seqlen=user_history_seqlen,
            item_id_hash_size=item_id_hash_size, item_id_embedding_dim=item_id_embedding_dim,
            item_features_size=item_features_size, user_value_weights=user_value_weights, mips_module=mips_module,
        )
        self.user_debias_net_user_value = nn.Sequential(nn.Linear(item_id_embedding_dim, 1))

    def debias_net_user_value(
        self, net_user_value: torch.Tensor, position: torch.Tensor, user_embedding: torch.Tensor
    ) -> Tuple[torch.Tensor, torch.Tensor]:
        """(:100-135) the clamp (1e-1) comes BEFORE the squared error here, unlike the position variant."""
        estimate = self.user_debias_net_user_value(user_embedding).squeeze(1)  # [B]
        estimate = torch.clamp(estimate, min=1e-1)
        estimate_loss = _sum_squared_error(estimate, net_user_value)
        return net_user_value / estimate, estimate_loss


class TwoTowerWithDebiasing(TwoTowerWithUserHistoryEncoder):
    """Position estimate fed, with the user embedding, into a `Linear(DI + 1, 1)` estimate of net_user_value."""

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
    ) -> None:
        super().__init__(
            num_items=num_items, user_id_hash_size=user_id_hash_size, user_id_embedding_dim=user_id_embedding_dim,
            user_features_size=user_features_size, user_history_seqlen=user_history_seqlen,
            item_id_hash_size=item_id_hash_size, item_id_embedding_dim=item_id_embedding_dim,
            item_features_size=item_features_size, user_value_weights=user_value_weights, mips_module=mips_module,
        )
        self.position_bias_net_user_value = nn.Embedding(num_embeddings=_POSITION_VOCAB, embedding_dim=1)
        self.user_debias_net_user_value = nn.Sequential(nn.Linear(item_id_embedding_dim + 1, 1))

    def debias_net_user_value(
        self, net_user_value: torch.Tensor, position: torch.Tensor, user_embedding: torch.Tensor
    ) -> Tuple[torch.Tensor, torch.Tensor]:
        """(:77-129) Note the reference compares the [B, 1] position estimate with the [B] target, which broadcasts
        to all B x B pairs (it emits a UserWarning at :110); that sum over pairs is reproduced here on purpose."""
        position_estimate = self.position_bias_net_user_value(position)  # [B, 1]
        user_estimate = self.user_debias_net_user_value(
            torch.cat([user_embedding, position_estimate], dim=-1)
        ).squeeze(1)  # [B]
        position_loss = _sum_squared_error(position_estimate, net_user_value)  # [B,1] - [B] -> [B,B]
        user_loss = _sum_squared_error(user_estimate, net_user_value)
        return net_user_value / torch.clamp(user_estimate, min=1e-3), user_loss + position_loss
