"""GAT baseline on the B200 edge-softmax kernel; same constructor, attributes and state-dict
keys as etpgt/model/gat.py."""

from __future__ import annotations

import torch.nn as nn

from .. import ops
from ..nn import GATConv, batch_norm_rows
from .base import BaseRecommendationModel, SessionReadout


class GAT(BaseRecommendationModel):
    def __init__(self, num_items: int, embedding_dim: int = 256, hidden_dim: int = 256, num_layers: int = 3,
                 num_heads: int = 4, dropout: float = 0.1, readout_type: str = "mean", concat_heads: bool = False):
        super().__init__(num_items, embedding_dim, hidden_dim, num_layers, dropout)
        self.num_heads, self.readout_type, self.concat_heads = num_heads, readout_type, concat_heads
        # layer plan of gat.py:45-111: one input layer, num_layers-2 hidden layers, and (when
        # num_layers > 1) a head-averaging output layer; only non-final layers may concatenate
        plan = [concat_heads] * (1 + max(num_layers - 2, 0)) + ([False] if num_layers > 1 else [])
        self.convs, self.batch_norms = nn.ModuleList(), nn.ModuleList()
        width = embedding_dim
        for concat in plan:
            self.convs.append(GATConv(width, hidden_dim, heads=num_heads, dropout=dropout, concat=concat))
            width = hidden_dim * num_heads if concat else hidden_dim
            self.batch_norms.append(nn.BatchNorm1d(width))
        self.dropout_layer = nn.Dropout(dropout)
        self.readout = SessionReadout(hidden_dim, readout_type)

    def forward(self, batch):
        ids, index = self._graph(batch)
        x = ops.EmbedPE.apply(ids, self.item_embedding.weight, None, False, None, None,
                              self.item_embedding.padding_idx)
        last = len(self.convs) - 1
        for layer, (conv, bn) in enumerate(zip(self.convs, self.batch_norms)):
            # BN, then ReLU + dropout on every layer but the last (gat.py:136-141)
            drop_p = self.dropout_layer.p if (self.training and layer < last) else 0.0
            x = batch_norm_rows(bn, conv(x, index), relu=layer < last, group=self.bn_process_group, drop_p=drop_p)
        return self.readout(x, batch.batch, self._num_sessions(batch))


def create_gat(num_items: int, embedding_dim: int = 256, hidden_dim: int = 256, num_layers: int = 3,
               num_heads: int = 4, dropout: float = 0.1, readout_type: str = "mean",
               concat_heads: bool = False) -> GAT:
    return GAT(num_items, embedding_dim, hidden_dim, num_layers, num_heads, dropout, readout_type, concat_heads)
