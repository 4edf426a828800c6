"""Laplacian positional encoding: cached [num_items, k] table + Linear(k -> D) projection
(etpgt/encodings/laplacian_pe.py).  The gather + projection is fused into the embedding kernel
(ops.EmbedPE).  The one-off eigendecomposition has two paths: `method="scipy"` (default) is the
reference's own call sequence on the host — the only way to reproduce what the reference stores in
`_cached_pe`, see the note at `compute_laplacian_pe` — and `method="device"` is a B200 eigen solver
(Chebyshev-filtered subspace iteration over etpgt_lap_sym_block) for graphs where the host solver takes
many minutes (the 1M-node configuration)."""

from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn


def compute_laplacian_pe(edge_index: torch.Tensor, num_nodes: int, k: int = 16, normalization: str = "sym",
                         method: str = "scipy", **solver_args):
    """|eigenvectors 1..k| of the sym-normalised graph Laplacian — laplacian_pe.py:19-66.

    method="scipy": the reference's sequence (PyG `get_laplacian` of the edge list AS GIVEN, scipy
    `eigsh(k+1, which="SM")`, drop the first vector, abs).  `train_baseline.py:234-243` passes the stored edge list,
    one direction per co-occurrence pair (item_i <= item_j): that Laplacian is upper triangular, not symmetric, and
    what the symmetric Lanczos solver returns for it is not an eigen decomposition of anything (measured on RR-shaped
    data: ||L v - lambda v|| = 0.3-0.7 for unit v) — an ARPACK artefact that only ARPACK itself reproduces.  This
    path keeps it bit-for-bit because checkpoints store it.
    method="device": eigenvectors of the UNDIRECTED graph's Laplacian (the edge list is symmetrised first), computed
    on the GPU; agrees with scipy on symmetric inputs (tests/test_gpu_laplacian.py)."""
    if method == "device":
        return compute_laplacian_pe_device(edge_index, num_nodes, k=k, normalization=normalization, **solver_args)
    if method != "scipy":
        raise ValueError(f"Unknown Laplacian PE method: {method}")
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


class _SymLaplacian:
    """L = I - D^-1/2 A D^-1/2 of the undirected graph as a device operator on blocks of fp64 vectors.
    Setup (once): symmetrise, drop self loops (PyG get_laplacian does), sort into CSR rows, deg^-1/2 with 0 for
    isolated nodes.  apply(): etpgt_lap_sym_block — one pass per Chebyshev step."""

    def __init__(self, edge_index: torch.Tensor, num_nodes: int, device):
        ei = edge_index.detach().to(device=device, dtype=torch.int64)
        row = torch.cat([ei[0], ei[1]])
        col = torch.cat([ei[1], ei[0]])
        keep = row != col
        row, col = row[keep], col[keep]
        # an undirected pair listed in both directions (or several times) must count once per direction
        key = torch.unique(row * num_nodes + col)
        row, col = key // num_nodes, key % num_nodes
        counts = torch.bincount(row, minlength=num_nodes)
        self.rowptr = torch.zeros(num_nodes + 1, dtype=torch.int64, device=device)
        self.rowptr[1:] = torch.cumsum(counts, 0)
        self.col = col.to(torch.int32).contiguous()          # `key` is sorted: rows ascending, columns ascending
        deg = counts.to(torch.float64)
        self.scale = torch.where(deg > 0, deg.clamp_min(1.0).rsqrt(), torch.zeros_like(deg))
        self.n, self.device = num_nodes, device

    def apply(self, x: torch.Tensor, alpha: float = 1.0, beta: float = 0.0, z: torch.Tensor | None = None,
              gamma: float = 0.0) -> torch.Tensor:
        """alpha * (L x) + beta * x + gamma * z for x, z [n, b] fp64 contiguous, b <= 32."""
        from .._lib import call, ptr, stream

        x = x.contiguous()                       # torch.linalg.qr hands back column-major factors
        z = None if z is None else z.contiguous()
        y = torch.empty_like(x)
        call("etpgt_lap_sym_block", ptr(self.rowptr), ptr(self.col), ptr(self.scale), self.n, x.size(1), ptr(x),
             ptr(z), float(alpha), float(beta), float(gamma), ptr(y), stream())
        return y


@torch.no_grad()
def smallest_eigenpairs_device(op: _SymLaplacian, count: int, tol: float = 1e-7, guard: int | None = None,
                               degree: int = 40, max_outer: int = 200, seed: int = 0):
    """The `count` smallest eigenpairs of the symmetric operator `op` (spectrum inside [0, 2]) by Chebyshev-filtered
    subspace iteration (Zhou & Saad): a block of count + guard vectors is repeatedly passed through a degree-`degree`
    Chebyshev polynomial of L that damps [cut, 2] — `cut` = the block's largest Ritz value — and amplifies what lies
    below it, re-orthonormalised (QR) and rotated to Ritz vectors (a (count+guard)^2 symmetric eigenproblem), until the
    wanted pairs have residuals ||L v - lambda v|| <= tol.  A block method: eigenvalues of multiplicity up to the block
    width (one zero per connected component) come out with their multiplicity, which single-vector Lanczos cannot do.
    Returns (values [count] ascending, vectors [n, count], residuals [count], iterations)."""
    n = op.n
    if guard is None:
        guard = max(8, count // 2)
    b = min(count + guard, 32, n)
    if count > b:
        raise ValueError(f"at most {b} eigenpairs per call (block width <= 32)")
    gen = torch.Generator(device=op.device).manual_seed(seed)
    x = torch.randn(n, b, dtype=torch.float64, device=op.device, generator=gen)
    x, _ = torch.linalg.qr(x)
    upper = 2.0
    res = vals = None
    for it in range(max_outer):
        ax = op.apply(x)
        vals, q = torch.linalg.eigh(x.t() @ ax)          # Rayleigh-Ritz: ascending
        x, ax = x @ q, ax @ q
        res = (ax - x * vals).norm(dim=0)
        if float(res[:count].max()) <= tol:
            break
        lo, cut = float(vals[0]), float(vals[-1])
        cut = min(max(cut, lo + 1e-8), upper - 1e-6)
        # scaled Chebyshev filter of degree `degree` that maps [cut, upper] to [-1, 1]
        e, c = (upper - cut) / 2.0, (upper + cut) / 2.0
        sigma = e / (lo - c)
        sigma1 = sigma
        y = op.apply(x, alpha=sigma1 / e, beta=-c * sigma1 / e)
        for _ in range(2, degree + 1):
            sigma2 = 1.0 / (2.0 / sigma1 - sigma)
            y_new = op.apply(y, alpha=2.0 * sigma2 / e, beta=-2.0 * c * sigma2 / e, z=x, gamma=-sigma * sigma2)
            x, y, sigma = y, y_new, sigma2
        x, _ = torch.linalg.qr(y)
    return vals[:count], x[:, :count], res[:count], it + 1


def compute_laplacian_pe_device(edge_index: torch.Tensor, num_nodes: int, k: int = 16, normalization: str = "sym",
                                device=None, tol: float = 1e-7, return_info: bool = False, **solver_args):
    """Device counterpart of `compute_laplacian_pe` for the undirected graph: the k+1 smallest eigenpairs of the
    sym-normalised Laplacian, first one dropped, absolute values, fp32 [num_nodes, k] on the device."""
    if normalization != "sym":
        raise NotImplementedError("the device solver handles the symmetric normalisation (the reference's default; "
                                  "the random-walk Laplacian is not symmetric)")
    if device is None:
        device = edge_index.device if edge_index.is_cuda else torch.device("cuda")
    op = _SymLaplacian(edge_index, num_nodes, device)
    vals, vecs, res, iters = smallest_eigenpairs_device(op, k + 1, tol=tol, **solver_args)
    pe = vecs[:, 1:k + 1].abs().float().contiguous()
    if return_info:
        return pe, {"eigenvalues": vals, "residuals": res, "iterations": iters}
    return pe


class LaplacianPECached(nn.Module):
    def __init__(self, k: int = 16, embedding_dim: int = 256, normalization: str = "sym"):
        super().__init__()
        self.k, self.embedding_dim, self.normalization = k, embedding_dim, normalization
        self.projection = nn.Linear(k, embedding_dim)
        nn.init.xavier_uniform_(self.projection.weight)
        nn.init.zeros_(self.projection.bias)
        self.register_buffer("_cached_pe", None)

    def precompute(self, data, method: str = "scipy") -> None:
        """`method="device"`: the B200 eigen solver on the undirected graph (see compute_laplacian_pe)."""
        pe = compute_laplacian_pe(data.edge_index, data.num_nodes, k=self.k, normalization=self.normalization,
                                  method=method)
        self._cached_pe = pe.to(self.projection.weight.device)

    def cached(self) -> torch.Tensor:
        if self._cached_pe is None:
            raise RuntimeError("Laplacian PE not precomputed. Call precompute() first.")
        return self._cached_pe

    def project(self, pe: torch.Tensor) -> torch.Tensor:
        return self.projection(pe)

    def forward(self, node_indices: torch.Tensor) -> torch.Tensor:
        return self.projection(self.cached()[node_indices])
