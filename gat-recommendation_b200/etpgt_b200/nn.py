"""Graph layers of the B200 path with the parameter names of the PyG layers they replace, so
reference checkpoints load unchanged (SURVEY.md §8b state-dict keys)."""

from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops


def _as_index(edge_index_or_index, num_nodes: int) -> ops.GraphIndex:
    if isinstance(edge_index_or_index, ops.GraphIndex):
        return edge_index_or_index
    return ops.GraphIndex(edge_index_or_index, num_nodes)


def _alpha_dropout_mask(num_edges: int, heads: int, p: float, training: bool, like: torch.Tensor):
    """Inverted-dropout mask over the attention weights, in original edge order."""
    if not training or p <= 0.0:
        return None
    return F.dropout(torch.ones(num_edges, heads, dtype=torch.float32, device=like.device), p=p, training=True)


class TransformerConv(nn.Module):
    """Drop-in for `torch_geometric.nn.TransformerConv(in, out, heads, dropout, concat=True,
    beta=...)` as constructed at etpgt/model/graph_transformer.py:73-98.  One dense projection
    produces query|key|value|skip for every node; everything per edge runs in the fused kernel."""

    def __init__(self, in_channels, out_channels, heads=1, concat=True, beta=False, dropout=0.0,
                 edge_dim=None, bias=True, root_weight=True):
        super().__init__()
        if not concat or edge_dim is not None or not root_weight or not bias:
            raise NotImplementedError("etpgt_b200.TransformerConv supports concat=True, edge_dim=None, "
                                      "root_weight=True, bias=True (the reference's configuration)")
        self.in_channels, self.out_channels, self.heads = in_channels, out_channels, heads
        self.concat, self.beta, self.dropout = concat, beta, dropout
        width = heads * out_channels
        self.lin_key = nn.Linear(in_channels, width)
        self.lin_query = nn.Linear(in_channels, width)
        self.lin_value = nn.Linear(in_channels, width)
        self.lin_skip = nn.Linear(in_channels, width)
        self.lin_beta = nn.Linear(3 * width, 1, bias=False) if beta else None

    def project(self, x: torch.Tensor) -> torch.Tensor:
        weight = torch.cat([self.lin_query.weight, self.lin_key.weight, self.lin_value.weight, self.lin_skip.weight])
        bias = torch.cat([self.lin_query.bias, self.lin_key.bias, self.lin_value.bias, self.lin_skip.bias])
        return F.linear(x, weight, bias)

    def forward(self, x, edge_index, alpha_mask=None):
        index = _as_index(edge_index, x.size(0))
        if alpha_mask is None:
            alpha_mask = _alpha_dropout_mask(index.num_edges, self.heads, self.dropout, self.training, x)
        qkvs = self.project(x)
        w_beta = None if self.lin_beta is None else self.lin_beta.weight
        return ops.TransformerConvFn.apply(qkvs, w_beta, alpha_mask, index, self.heads)


def batch_norm_rows(bn: nn.BatchNorm1d, x: torch.Tensor, residual: torch.Tensor | None = None,
                    relu: bool = False, group=None) -> torch.Tensor:
    """`bn(x) (+ residual) (-> relu)` with the statistics of an ordinary nn.BatchNorm1d module
    (its parameters / buffers are used and updated in place, so state dicts stay compatible)."""
    training = bn.training or not bn.track_running_stats
    if training and bn.track_running_stats and bn.num_batches_tracked is not None:
        bn.num_batches_tracked.add_(1)
    momentum = 0.1 if bn.momentum is None else bn.momentum
    return ops.BatchNormRows.apply(x, bn.weight, bn.bias, residual, bn.running_mean, bn.running_var, training,
                                   momentum, bn.eps, relu, group)
