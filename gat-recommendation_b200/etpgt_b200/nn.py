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
    return ops.dropout_mask(num_edges * heads, p, like.device).view(num_edges, heads)


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

    def _fused(self, name: str):
        """query | key | value | skip `name` ("weight" / "bias") as ONE tensor without copying: the four
        parameters are re-homed (once per device move) as consecutive row blocks of a single buffer, and
        ops.FusedRows presents that buffer to autograd as a function of the four parameters, so state
        dicts, optimizers and checkpoints keep seeing four ordinary parameters."""
        store, parts = self.fused_store(name)
        return ops.FusedRows.apply(store, *parts)

    def fused_store(self, name: str):
        """(the single buffer holding query | key | value | skip `name`, the four parameters aliasing it)."""
        parts = [getattr(lin, name) for lin in (self.lin_query, self.lin_key, self.lin_value, self.lin_skip)]
        cache = self.__dict__.setdefault("_fused_store", {})
        store = cache.get(name)
        offset, aliased = 0, store is not None
        if aliased:
            for p in parts:
                aliased = aliased and p.data_ptr() == store.data_ptr() + offset * store.element_size() \
                    and p.device == store.device and p.dtype == store.dtype
                offset += p.numel()
        if not aliased:
            with torch.no_grad():
                store = torch.cat([p.detach() for p in parts]).contiguous()
                row = 0
                for p in parts:
                    p.data = store[row:row + p.size(0)]
                    row += p.size(0)
            cache[name] = store
        return store, parts

    def fused_parameters(self):
        if self.lin_query.weight.is_cuda and torch.is_grad_enabled():
            return self._fused("weight"), self._fused("bias")
        weight = torch.cat([self.lin_query.weight, self.lin_key.weight, self.lin_value.weight, self.lin_skip.weight])
        bias = torch.cat([self.lin_query.bias, self.lin_key.bias, self.lin_value.bias, self.lin_skip.bias])
        return weight, bias

    def project(self, x: torch.Tensor) -> torch.Tensor:
        weight, bias = self.fused_parameters()
        return ops.linear(x, weight, bias)

    def forward(self, x, edge_index, alpha_mask=None):
        index = _as_index(edge_index, x.size(0))
        if alpha_mask is None:
            alpha_mask = _alpha_dropout_mask(index.num_edges, self.heads, self.dropout, self.training, x)
        w_beta = None if self.lin_beta is None else self.lin_beta.weight
        if ops.fused_conv_supported(x, self.in_channels, 4 * self.heads * self.out_channels):
            weight, bias = self.fused_parameters()
            return ops.TransformerConvLayer.apply(x, weight, bias, w_beta, alpha_mask, index, self.heads)
        qkvs = self.project(x)
        return ops.TransformerConvFn.apply(qkvs, w_beta, alpha_mask, index, self.heads)


class GATConv(nn.Module):
    """Drop-in for `torch_geometric.nn.GATConv(in, out, heads, dropout, concat)` as constructed at
    etpgt/model/gat.py:49-109 (PyG >= 2.5 parameter names: lin, att_src, att_dst, bias)."""

    def __init__(self, in_channels, out_channels, heads=1, concat=True, negative_slope=0.2, dropout=0.0,
                 add_self_loops=True, bias=True):
        super().__init__()
        if not add_self_loops:
            raise NotImplementedError("etpgt_b200.GATConv implements add_self_loops=True (the reference's use)")
        self.in_channels, self.out_channels, self.heads = in_channels, out_channels, heads
        self.concat, self.negative_slope, self.dropout = concat, negative_slope, dropout
        self.lin = nn.Linear(in_channels, heads * out_channels, bias=False)
        self.att_src = nn.Parameter(torch.empty(1, heads, out_channels))
        self.att_dst = nn.Parameter(torch.empty(1, heads, out_channels))
        self.bias = nn.Parameter(torch.zeros(heads * out_channels if concat else out_channels)) if bias else None
        nn.init.xavier_uniform_(self.lin.weight)
        nn.init.xavier_uniform_(self.att_src)
        nn.init.xavier_uniform_(self.att_dst)

    def forward(self, x, edge_index, mask_edges=None, mask_self=None):
        index = _as_index(edge_index, x.size(0))
        n, heads, c = x.size(0), self.heads, self.out_channels
        if mask_edges is None and self.training and self.dropout > 0.0:
            masks = ops.dropout_mask((index.num_edges + n) * heads, self.dropout, x.device)   # one launch for both
            mask_edges = masks[: index.num_edges * heads].view(index.num_edges, heads)
            mask_self = masks[index.num_edges * heads:].view(n, heads)
        if not self.concat and ops.gat_layer_supported(x, self.in_channels, heads * c, heads):
            # projection, attention scalars from the input, edge softmax + aggregation + head mean: one autograd node
            return ops.GatLayerFn.apply(x, self.lin.weight, self.att_src, self.att_dst, self.bias, mask_edges,
                                        mask_self, index, heads, self.negative_slope)
        h = ops.linear(x, self.lin.weight, None)      # tcgen05 split-bf16 GEMM (fp32-grade)
        if not self.concat and c % 4 == 0 and h.size(1) <= 1024:
            # attention scalars, edge softmax + aggregation, head mean + bias: one autograd node
            return ops.GatConvFn.apply(h, self.att_src, self.att_dst, self.bias, mask_edges, mask_self, index, heads,
                                       self.negative_slope)
        hv = h.view(n, heads, c)
        a_src = (hv * self.att_src).sum(-1)
        a_dst = (hv * self.att_dst).sum(-1)
        agg = ops.GatAggregateFn.apply(h, a_src, a_dst, mask_edges, mask_self, index, heads, self.negative_slope)
        out = agg if self.concat else agg.view(n, heads, c).mean(dim=1)
        return out if self.bias is None else out + self.bias


class SAGEConv(nn.Module):
    """Drop-in for `torch_geometric.nn.SAGEConv(in, out, aggr="mean")` (etpgt/model/graphsage.py:43-48):
    lin_l(mean of in-neighbours) + lin_r(x), lin_l with bias, lin_r without."""

    def __init__(self, in_channels, out_channels, aggr="mean", normalize=False, root_weight=True, project=False,
                 bias=True):
        super().__init__()
        if aggr != "mean" or normalize or project or not root_weight:
            raise NotImplementedError("etpgt_b200.SAGEConv implements aggr='mean' (the only aggregator the "
                                      "reference uses, scripts/evaluate_local.py:43)")
        self.in_channels, self.out_channels, self.aggr = in_channels, out_channels, aggr
        self.lin_l = nn.Linear(in_channels, out_channels, bias=bias)
        self.lin_r = nn.Linear(in_channels, out_channels, bias=False)

    def forward(self, x, edge_index):
        index = _as_index(edge_index, x.size(0))
        if ops.sage_layer_supported(x, self.in_channels, self.out_channels):
            # one GEMM over [mean | x] against [W_l | W_r] (and its two backward GEMMs) instead of two of each
            return ops.SageLayerFn.apply(x, self.lin_l.weight, self.lin_l.bias, self.lin_r.weight, index)
        mean = ops.SageMeanFn.apply(x, index)
        return ops.linear(mean, self.lin_l.weight, self.lin_l.bias) + ops.linear(x, self.lin_r.weight, None)


def batch_norm_rows(bn: nn.BatchNorm1d, x: torch.Tensor, residual: torch.Tensor | None = None,
                    relu: bool = False, group=None, drop_p: float = 0.0) -> torch.Tensor:
    """`bn(x) (+ residual) (-> relu) (-> dropout_p)` with the statistics of an ordinary nn.BatchNorm1d
    module (its parameters / buffers are used and updated in place, so state dicts stay compatible).
    drop_p > 0 fuses the layer's nn.Dropout into the same kernel (Philox mask, seed from torch's CPU RNG)."""
    training = bn.training or not bn.track_running_stats
    if training and bn.track_running_stats and bn.num_batches_tracked is not None:
        bn.num_batches_tracked.add_(1)
    momentum = 0.1 if bn.momentum is None else bn.momentum
    seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if drop_p > 0.0 else 0
    return ops.BatchNormRows.apply(x, bn.weight, bn.bias, residual, bn.running_mean, bn.running_var, training,
                                   momentum, bn.eps, relu, group, float(drop_p), seed)


def transformer_layer(conv: TransformerConv, bn: nn.BatchNorm1d, x: torch.Tensor, x_split, index, drop_p: float,
                      group=None, want_split: bool = False):
    """dropout(bn(conv(x)) + x) as ONE fused autograd node (ops.TransformerLayer); returns (y, split of
    y or None).  Parameters / buffers of `conv` and `bn` are used and updated in place, so state dicts
    stay those of the reference."""
    training = bn.training or not bn.track_running_stats
    if training and bn.track_running_stats and bn.num_batches_tracked is not None:
        bn.num_batches_tracked.add_(1)
    momentum = 0.1 if bn.momentum is None else bn.momentum
    alpha_mask = _alpha_dropout_mask(index.num_edges, conv.heads, conv.dropout, conv.training, x)
    weight, bias = conv.fused_parameters()
    w_beta = None if conv.lin_beta is None else conv.lin_beta.weight
    seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if drop_p > 0.0 else 0   # torch's CPU generator
    y, y_hi, y_lo = ops.TransformerLayer.apply(
        x, x_split, weight, bias, w_beta, alpha_mask, index, conv.heads, bn.weight, bn.bias, bn.running_mean,
        bn.running_var, training, momentum, bn.eps, group, drop_p, seed, want_split)
    return y, ((y_hi, y_lo) if want_split else None)


def fused_layer_supported(conv: TransformerConv, bn: nn.BatchNorm1d, x: torch.Tensor) -> bool:
    width = 4 * conv.heads * conv.out_channels
    return ops.FUSED_LAYER and isinstance(conv, TransformerConv) and bn.affine and bn.track_running_stats and \
        x.size(1) == conv.heads * conv.out_channels and ops.fused_conv_supported(x, conv.in_channels, width)
