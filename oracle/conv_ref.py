"""TEST INFRASTRUCTURE ONLY — CPU oracle for the three graph convolutions on the hot path.

The reference delegates them to `torch_geometric` (absent from /root/reference and from this
image; `torch-geometric>=2.4.0`, requirements.txt:5), so the published algorithms of upstream
`nn/conv/transformer_conv.py`, `nn/conv/gat_conv.py`, `nn/conv/sage_conv.py` and
`utils/_softmax.py` are restated here as pure functions of (x, edge_index, weights).
Reference call sites: etpgt/model/graph_transformer.py:73-98,174; etpgt/model/gat.py:49-109,137;
etpgt/model/graphsage.py:43-48,75.  PARITY UNPINNED at this boundary (no numeric test in the
reference touches a conv output); everything is differentiable torch so fp64 runs give the
ground truth for gradients as well.

Convention everywhere: edge_index[0] = source j, edge_index[1] = target i, aggregation at i.
"""

from __future__ import annotations

import math
import sys
from pathlib import Path

import torch
import torch.nn.functional as F

_SHIM = str(Path(__file__).resolve().parent / "pyg_shim")
if _SHIM not in sys.path:
    sys.path.insert(0, _SHIM)

from torch_geometric.utils import add_self_loops, remove_self_loops, segment_softmax  # noqa: E402


def _scatter_rows(values: torch.Tensor, index: torch.Tensor, num_nodes: int) -> torch.Tensor:
    out = torch.zeros((num_nodes,) + tuple(values.shape[1:]), dtype=values.dtype, device=values.device)
    return out.index_add(0, index, values)


def transformer_conv(
    x: torch.Tensor,
    edge_index: torch.Tensor,
    w_query: torch.Tensor,
    b_query: torch.Tensor,
    w_key: torch.Tensor,
    b_key: torch.Tensor,
    w_value: torch.Tensor,
    b_value: torch.Tensor,
    w_skip: torch.Tensor,
    b_skip: torch.Tensor,
    w_beta: torch.Tensor | None,
    heads: int,
    edge_alpha_mask: torch.Tensor | None = None,
) -> torch.Tensor:
    """TransformerConv(concat=True, root_weight=True, edge_dim=None), optional beta gate.

    `edge_alpha_mask` [E, heads] multiplies the attention weights after the softmax; it is
    how a training-mode dropout mask (already scaled by 1/(1-p)) is injected so that the
    CUDA path can be compared under dropout."""
    n = x.size(0)
    c = w_query.size(0) // heads
    src, dst = edge_index[0], edge_index[1]
    q = F.linear(x, w_query, b_query).view(n, heads, c)
    k = F.linear(x, w_key, b_key).view(n, heads, c)
    v = F.linear(x, w_value, b_value).view(n, heads, c)
    logits = (q[dst] * k[src]).sum(-1) / math.sqrt(c)
    alpha = segment_softmax(logits, dst, n)
    if edge_alpha_mask is not None:
        alpha = alpha * edge_alpha_mask
    agg = _scatter_rows(v[src] * alpha.unsqueeze(-1), dst, n).reshape(n, heads * c)
    x_r = F.linear(x, w_skip, b_skip)
    if w_beta is None:
        return agg + x_r
    beta = torch.sigmoid(F.linear(torch.cat([agg, x_r, agg - x_r], dim=-1), w_beta))
    return beta * x_r + (1.0 - beta) * agg


def gat_conv(
    x: torch.Tensor,
    edge_index: torch.Tensor,
    w_lin: torch.Tensor,
    att_src: torch.Tensor,
    att_dst: torch.Tensor,
    bias: torch.Tensor | None,
    heads: int,
    concat: bool,
    negative_slope: float = 0.2,
    edge_alpha_mask: torch.Tensor | None = None,
) -> torch.Tensor:
    """GATConv(add_self_loops=True): existing self-loops are dropped and exactly one per
    node is appended after the real edges; scalar logit per (edge, head)."""
    n = x.size(0)
    c = w_lin.size(0) // heads
    h = F.linear(x, w_lin).view(n, heads, c)
    a_src = (h * att_src.view(1, heads, c)).sum(-1)
    a_dst = (h * att_dst.view(1, heads, c)).sum(-1)
    edge_index, _ = remove_self_loops(edge_index)
    edge_index, _ = add_self_loops(edge_index, num_nodes=n)
    src, dst = edge_index[0], edge_index[1]
    e = F.leaky_relu(a_src[src] + a_dst[dst], negative_slope)
    alpha = segment_softmax(e, dst, n)
    if edge_alpha_mask is not None:
        alpha = alpha * edge_alpha_mask
    out = _scatter_rows(h[src] * alpha.unsqueeze(-1), dst, n)
    out = out.reshape(n, heads * c) if concat else out.mean(dim=1)
    return out if bias is None else out + bias


def sage_conv(
    x: torch.Tensor,
    edge_index: torch.Tensor,
    w_l: torch.Tensor,
    b_l: torch.Tensor | None,
    w_r: torch.Tensor,
) -> torch.Tensor:
    """SAGEConv(aggr="mean", root_weight=True): lin_l(mean_j x_j) + lin_r(x_i); the mean over
    zero neighbours is 0."""
    n = x.size(0)
    src, dst = edge_index[0], edge_index[1]
    summed = _scatter_rows(x[src], dst, n)
    deg = torch.zeros(n, dtype=x.dtype, device=x.device).index_add(
        0, dst, torch.ones(dst.numel(), dtype=x.dtype, device=x.device)
    )
    mean = summed / deg.clamp(min=1.0).unsqueeze(-1)
    return F.linear(mean, w_l, b_l) + F.linear(x, w_r)
