"""GraphSAGE baseline on the B200 mean-aggregation kernel; same constructor, attributes and
state-dict keys as etpgt/model/graphsage.py."""

from __future__ import annotations

import torch.nn as nn

from .. import ops
from ..nn import SAGEConv, batch_norm_rows
from .base import BaseRecommendationModel, SessionReadout


class GraphSAGE(BaseRecommendationModel):
    def __init__(self, num_items: int, embedding_dim: int = 256, hidden_dim: int = 256, num_layers: int = 3,
                 dropout: float = 0.1, readout_type: str = "mean", aggregator: str = "mean"):
        super().__init__(num_items, embedding_dim, hidden_dim, num_layers, dropout)
        self.aggregator, self.readout_type = aggregator, readout_type
        widths = [embedding_dim] + [hidden_dim] * (num_layers - 1)
        self.convs = nn.ModuleList(SAGEConv(w, hidden_dim, aggr=aggregator) for w in widths)
        self.batch_norms = nn.ModuleList(nn.BatchNorm1d(hidden_dim) for _ in widths)
        self.dropout_layer = nn.Dropout(dropout)
        self.readout = SessionReadout(hidden_dim, readout_type)

    def forward(self, batch):
        ids, index = self._graph(batch)
        x = ops.EmbedPE.apply(ids, self.item_embedding.weight, None, False, None, None,
                              self.item_embedding.padding_idx)
        for conv, bn in zip(self.convs, self.batch_norms):
            # BN -> ReLU -> dropout on every layer (graphsage.py:74-78)
            x = batch_norm_rows(bn, conv(x, index), relu=True, group=self.bn_process_group,
                                drop_p=self.dropout_layer.p if self.training else 0.0)   # BN + ReLU + dropout, one kernel
        return self.readout(x, batch.batch, self._num_sessions(batch))


def create_graphsage(num_items: int, embedding_dim: int = 256, hidden_dim: int = 256, num_layers: int = 3,
                     dropout: float = 0.1, readout_type: str = "mean", aggregator: str = "mean") -> GraphSAGE:
    return GraphSAGE(num_items, embedding_dim, hidden_dim, num_layers, dropout, readout_type, aggregator)
