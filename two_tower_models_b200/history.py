"""Host-side mirror of the reference's `UserHistoryEncoder` (src/user_history_encoder.py:11-124).

Same constructor arguments, attributes, `forward([B,H,DI]) -> [B,2,DI]`, `positional_encoding`,
`get_output_dim` and parameter names (`multihead_attn_layers.{l}.in_proj_weight`, `.in_proj_bias`,
`.out_proj.weight`, `.out_proj.bias`).  The `nn.MultiheadAttention` sub-modules only host the fp32
master parameters (so default initialisation and the RNG stream match the reference, which its
known-answer tests depend on); their forward is never called - the mean-pool, positional add, packed
in-projection, per-head softmax(QK^T)V and out-projection run in libtt_b200.so through `ops`.
"""
import math

import torch
import torch.nn as nn

from . import ops


class UserHistoryEncoder(nn.Module):
    """[B, H, DI] history embeddings -> [B, 2, DI] = stack(attention output of the newest item, mean-pool)."""

    def __init__(
        self,
        item_id_embedding_dim: int,
        history_len: int,
        num_attention_heads: int,
        num_attention_layers: int,
        use_positional_encoding: bool,
    ) -> None:
        super().__init__()
        self.item_id_embedding_dim = item_id_embedding_dim
        self.history_len = history_len
        self.num_attention_heads = num_attention_heads
        self.num_attention_layers = num_attention_layers
        self.use_positional_encoding = use_positional_encoding
        if item_id_embedding_dim % num_attention_heads != 0:
            raise AssertionError("embed_dim must be divisible by num_heads")  # nn.MultiheadAttention's own check

        if self.use_positional_encoding:
            # fixed (non-learned) table, flipped because history index 0 is the newest item (reference :35-54).
            # Non-persistent buffer: follows .to(device), stays out of state_dict like the reference's attribute.
            pe = self.positional_encoding(seq_len=history_len, d_model=item_id_embedding_dim).flip([0])
            self.register_buffer("positional_embeddings", pe, persistent=False)

        self.multihead_attn_layers = nn.ModuleList(
            [
                nn.MultiheadAttention(embed_dim=item_id_embedding_dim, num_heads=self.num_attention_heads)
                for _ in range(self.num_attention_layers)
            ]
        )
        self._packed = ops.PackedWeights()

    def positional_encoding(self, seq_len: int, d_model: int) -> torch.Tensor:
        """The reference's own sinusoid table (:69-78): sin and cos columns use different frequencies."""
        # evaluated with python floats (libm doubles) exactly like the reference so the fp32 table is bit-identical
        wave = (math.sin, math.cos)
        rows = [
            [wave[c & 1](p / (10000 ** ((2 * c) / d_model))) for c in range(d_model)]
            for p in range(seq_len)
        ]
        return torch.tensor(rows, dtype=torch.float64).to(torch.float32).reshape(seq_len, d_model)

    def layer_parameters(self):
        out = []
        for layer in self.multihead_attn_layers:
            out += [layer.in_proj_weight, layer.in_proj_bias, layer.out_proj.weight, layer.out_proj.bias]
        return out

    def forward(self, user_history: torch.Tensor) -> torch.Tensor:
        """user_history [B, H, DI] (newest item first) -> summary [B, 2, DI] (reference :80-121)."""
        if user_history.dim() != 3 or user_history.shape[2] != self.item_id_embedding_dim:
            raise RuntimeError(
                f"user_history must be [B, H, {self.item_id_embedding_dim}], got {tuple(user_history.shape)}"
            )
        pe = self.positional_embeddings if self.use_positional_encoding else None
        if pe is not None and pe.shape[0] != user_history.shape[1]:
            raise RuntimeError(
                f"The size of tensor a ({user_history.shape[1]}) must match the size of tensor b ({pe.shape[0]}) "
                "at non-singleton dimension 1"
            )
        summary = ops.HistoryEncoderFunction.apply(
            None, user_history, pe, self.num_attention_heads, self._packed, "enc", *self.layer_parameters()
        )  # [B, 2*DI] = [most_recent | mean_pool]
        return summary.view(user_history.shape[0], 2, self.item_id_embedding_dim)

    def encode_ids(self, table: torch.Tensor, history_ids: torch.Tensor) -> torch.Tensor:
        """Fused path of TwoTowerWithUserHistoryEncoder: embedding lookup of history ids in `table`
        (reference src/two_tower_with_user_history_encoder.py:105) folded into the encoder; -> [B, 2*DI]."""
        pe = self.positional_embeddings if self.use_positional_encoding else None
        return ops.HistoryEncoderFunction.apply(
            history_ids, table, pe, self.num_attention_heads, self._packed, "enc", *self.layer_parameters()
        )

    def get_output_dim(self) -> int:
        return self.item_id_embedding_dim * 2
