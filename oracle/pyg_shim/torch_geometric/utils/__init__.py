"""Oracle stand-in for the `torch_geometric.utils` helpers on the reference path
(test infrastructure only).  Restated from upstream PyG >= 2.4:
`utils/_softmax.py`, `utils/laplacian.py`, `utils/loop.py`, `utils/convert.py`.
Call sites in the reference: etpgt/encodings/laplacian_pe.py:40-47.
"""

from __future__ import annotations

import torch


def segment_softmax(src: torch.Tensor, index: torch.Tensor, num_nodes: int) -> torch.Tensor:
    """PyG `utils.softmax(src, index, num_nodes=N)`: subtract the (detached) per-target
    maximum, exponentiate, divide by the per-target sum plus 1e-16."""
    shape = (num_nodes,) + tuple(src.shape[1:])
    idx = index.view(-1, *([1] * (src.dim() - 1))).expand_as(src)
    seg_max = torch.full(shape, float("-inf"), dtype=src.dtype, device=src.device)
    seg_max = seg_max.scatter_reduce(0, idx, src.detach(), reduce="amax", include_self=True)
    p = (src - seg_max.gather(0, idx)).exp()
    seg_sum = torch.zeros(shape, dtype=src.dtype, device=src.device).scatter_add(0, idx, p)
    return p / (seg_sum.gather(0, idx) + 1e-16)


softmax = segment_softmax


def remove_self_loops(edge_index, edge_attr=None):
    keep = edge_index[0] != edge_index[1]
    return edge_index[:, keep], (None if edge_attr is None else edge_attr[keep])


def add_self_loops(edge_index, edge_attr=None, fill_value=1.0, num_nodes=None):
    n = int(num_nodes)
    loops = torch.arange(n, dtype=edge_index.dtype, device=edge_index.device)
    edge_index = torch.cat([edge_index, torch.stack([loops, loops])], dim=1)
    if edge_attr is not None:
        fill = torch.full((n,), float(fill_value), dtype=edge_attr.dtype, device=edge_attr.device)
        edge_attr = torch.cat([edge_attr, fill])
    return edge_index, edge_attr


def get_laplacian(edge_index, edge_weight=None, normalization=None, dtype=None, num_nodes=None):
    """Unit weights, self-loops dropped, `deg = scatter_add(w, row)`; "sym":
    `-deg^-1/2[row] * w * deg^-1/2[col]` (inf -> 0) plus N unit self-loops.  Does not
    symmetrise its input."""
    edge_index, edge_weight = remove_self_loops(edge_index, edge_weight)
    if edge_weight is None:
        edge_weight = torch.ones(edge_index.size(1), dtype=dtype or torch.float32)
    n = int(num_nodes) if num_nodes is not None else int(edge_index.max()) + 1
    row, col = edge_index[0], edge_index[1]
    deg = torch.zeros(n, dtype=edge_weight.dtype).scatter_add(0, row, edge_weight)
    if normalization is None:
        edge_index = torch.cat([edge_index, torch.arange(n).repeat(2, 1)], dim=1)
        edge_weight = torch.cat([-edge_weight, deg])
    elif normalization == "sym":
        dis = deg.pow(-0.5)
        dis[torch.isinf(dis)] = 0.0
        edge_weight = dis[row] * edge_weight * dis[col]
        edge_index, edge_weight = add_self_loops(edge_index, -edge_weight, 1.0, n)
    else:  # "rw"
        dinv = 1.0 / deg
        dinv[torch.isinf(dinv)] = 0.0
        edge_weight = dinv[row] * edge_weight
        edge_index, edge_weight = add_self_loops(edge_index, -edge_weight, 1.0, n)
    return edge_index, edge_weight


def to_scipy_sparse_matrix(edge_index, edge_attr=None, num_nodes=None):
    import scipy.sparse

    row, col = edge_index.cpu().numpy()
    if edge_attr is None:
        edge_attr = torch.ones(row.shape[0])
    n = int(num_nodes) if num_nodes is not None else int(edge_index.max()) + 1
    return scipy.sparse.coo_matrix((edge_attr.cpu().numpy(), (row, col)), (n, n))
