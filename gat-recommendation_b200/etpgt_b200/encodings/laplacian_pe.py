"""Laplacian positional encoding: cached [num_items, k] table + Linear(k -> D) projection
(etpgt/encodings/laplacian_pe.py).  The gather + projection is fused into the embedding kernel
(ops.EmbedPE); the one-off eigendecomposition stays a host precompute."""

from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn


def compute_laplacian_pe(edge_index: torch.Tensor, num_nodes: int, k: int = 16, normalization: str = "sym"):
    """|eigenvectors 1..k| of the (sym-normalised, not symmetrised) graph Laplacian; one-off CPU
    setup that feeds `_cached_pe` — laplacian_pe.py:19-66."""
    try:
        import scipy.sparse as sp
        from scipy.sparse.linalg import eigsh
    except ImportError as exc:  # pragma: no cover
        raise ImportError("scipy is required for Laplacian PE computation") from exc
    row, col = edge_index.detach().cpu().numpy()
    off = row != col
    row, col = row[off], col[off]
    w = np.ones(row.shape[0], dtype=np.float32)
    deg = np.bincount(row, weights=w, minlength=num_nodes).astype(np.float32)
    with np.errstate(divide="ignore"):
        scale = np.where(deg > 0, deg ** (-0.5 if normalization == "sym" else -1.0), 0.0).astype(np.float32)
    vals = -(scale[row] * w * scale[col]) if normalization == "sym" else -(scale[row] * w)
    diag = np.arange(num_nodes)
    lap = sp.coo_matrix((np.concatenate([vals, np.ones(num_nodes, dtype=np.float32)]),
                         (np.concatenate([row, diag]), np.concatenate([col, diag]))), (num_nodes, num_nodes))
    try:
        _, vecs = eigsh(lap, k=k + 1, which="SM", return_eigenvectors=True)
    except Exception:
        _, vecs_t = torch.linalg.eigh(torch.from_numpy(lap.toarray()).float())
        vecs = vecs_t.numpy()
    return torch.from_numpy(np.ascontiguousarray(vecs[:, 1:k + 1])).float().abs()


class LaplacianPECached(nn.Module):
    def __init__(self, k: int = 16, embedding_dim: int = 256, normalization: str = "sym"):
        super().__init__()
        self.k, self.embedding_dim, self.normalization = k, embedding_dim, normalization
        self.projection = nn.Linear(k, embedding_dim)
        nn.init.xavier_uniform_(self.projection.weight)
        nn.init.zeros_(self.projection.bias)
        self.register_buffer("_cached_pe", None)

    def precompute(self, data) -> None:
        pe = compute_laplacian_pe(data.edge_index, data.num_nodes, k=self.k, normalization=self.normalization)
        self._cached_pe = pe.to(self.projection.weight.device)

    def cached(self) -> torch.Tensor:
        if self._cached_pe is None:
            raise RuntimeError("Laplacian PE not precomputed. Call precompute() first.")
        return self._cached_pe

    def project(self, pe: torch.Tensor) -> torch.Tensor:
        return self.projection(pe)

    def forward(self, node_indices: torch.Tensor) -> torch.Tensor:
        return self.projection(self.cached()[node_indices])
