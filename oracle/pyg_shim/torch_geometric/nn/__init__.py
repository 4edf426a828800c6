"""Oracle stand-in for `torch_geometric.nn.{TransformerConv,GATConv,SAGEConv}` (test
infrastructure only).  Thin `nn.Module` shells with upstream PyG >= 2.5 parameter names
(`lin_key/lin_query/lin_value/lin_skip/lin_beta`; `lin/att_src/att_dst/bias`;
`lin_l/lin_r`) around the restated math in `oracle/conv_ref.py`.  The parameter structure
is pinned by the reference's published parameter counts (docs/EXPERIMENTS.md:85-88).
"""

from __future__ import annotations

import sys
from pathlib import Path

import torch
import torch.nn as nn
import torch.nn.functional as F

_ORACLE_PARENT = str(Path(__file__).resolve().parents[4])
if _ORACLE_PARENT not in sys.path:
    sys.path.insert(0, _ORACLE_PARENT)

from oracle import conv_ref  # noqa: E402


def _dropout_mask(shape, p, training, like):
    if not training or p <= 0.0:
        return None
    return F.dropout(torch.ones(shape, dtype=like.dtype, device=like.device), p=p, training=True)


class TransformerConv(nn.Module):
    def __init__(self, in_channels, out_channels, heads=1, concat=True, beta=False,
                 dropout=0.0, edge_dim=None, bias=True, root_weight=True):
        super().__init__()
        if not concat or edge_dim is not None or not root_weight:
            raise NotImplementedError("oracle shim covers the reference's configuration only")
        self.in_channels, self.out_channels, self.heads = in_channels, out_channels, heads
        self.concat, self.beta, self.dropout = concat, beta, dropout
        width = heads * out_channels
        self.lin_key = nn.Linear(in_channels, width)
        self.lin_query = nn.Linear(in_channels, width)
        self.lin_value = nn.Linear(in_channels, width)
        self.lin_skip = nn.Linear(in_channels, width, bias=bias)
        self.lin_beta = nn.Linear(3 * width, 1, bias=False) if beta else None

    def forward(self, x, edge_index):
        mask = _dropout_mask((edge_index.size(1), self.heads), self.dropout, self.training, x)
        return conv_ref.transformer_conv(
            x, edge_index,
            self.lin_query.weight, self.lin_query.bias,
            self.lin_key.weight, self.lin_key.bias,
            self.lin_value.weight, self.lin_value.bias,
            self.lin_skip.weight, self.lin_skip.bias,
            None if self.lin_beta is None else self.lin_beta.weight,
            self.heads, mask,
        )


class GATConv(nn.Module):
    def __init__(self, in_channels, out_channels, heads=1, concat=True, negative_slope=0.2,
                 dropout=0.0, add_self_loops=True, bias=True):
        super().__init__()
        if not add_self_loops:
            raise NotImplementedError("oracle shim covers the reference's configuration only")
        self.in_channels, self.out_channels, self.heads = in_channels, out_channels, heads
        self.concat, self.negative_slope, self.dropout = concat, negative_slope, dropout
        self.lin = nn.Linear(in_channels, heads * out_channels, bias=False)
        self.att_src = nn.Parameter(torch.empty(1, heads, out_channels))
        self.att_dst = nn.Parameter(torch.empty(1, heads, out_channels))
        self.bias = nn.Parameter(torch.zeros(heads * out_channels if concat else out_channels)) if bias else None
        nn.init.xavier_uniform_(self.lin.weight)
        nn.init.xavier_uniform_(self.att_src)
        nn.init.xavier_uniform_(self.att_dst)

    def forward(self, x, edge_index):
        n = x.size(0)
        num_edges = int((edge_index[0] != edge_index[1]).sum()) + n
        mask = _dropout_mask((num_edges, self.heads), self.dropout, self.training, x)
        return conv_ref.gat_conv(x, edge_index, self.lin.weight, self.att_src, self.att_dst,
                                 self.bias, self.heads, self.concat, self.negative_slope, mask)


class SAGEConv(nn.Module):
    def __init__(self, in_channels, out_channels, aggr="mean", normalize=False,
                 root_weight=True, project=False, bias=True):
        super().__init__()
        if aggr != "mean" or normalize or project or not root_weight:
            raise NotImplementedError("oracle shim covers the reference's configuration only")
        self.in_channels, self.out_channels, self.aggr = in_channels, out_channels, aggr
        self.lin_l = nn.Linear(in_channels, out_channels, bias=bias)
        self.lin_r = nn.Linear(in_channels, out_channels, bias=False)

    def forward(self, x, edge_index):
        return conv_ref.sage_conv(x, edge_index, self.lin_l.weight, self.lin_l.bias, self.lin_r.weight)
