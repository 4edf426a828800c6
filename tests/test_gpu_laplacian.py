"""GPU tests of the device Laplacian-PE eigen solver (etpgt/encodings/laplacian_pe.py:19-66).
Oracle: scipy's eigsh on the same symmetric Laplacian — the call the reference makes.  (On the reference's own
NON-symmetric input eigsh's output is not an eigen decomposition, see compute_laplacian_pe; nothing to compare.)"""

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _scipy_laplacian(edge_index: np.ndarray, n: int):
    """PyG get_laplacian(normalization="sym") of the symmetrised, de-duplicated edge list as a scipy matrix."""
    import scipy.sparse as sp

    row = np.concatenate([edge_index[0], edge_index[1]])
    col = np.concatenate([edge_index[1], edge_index[0]])
    keep = row != col
    key = np.unique(row[keep].astype(np.int64) * n + col[keep])
    row, col = key // n, key % n
    deg = np.bincount(row, minlength=n).astype(np.float64)
    scale = np.where(deg > 0, 1.0 / np.sqrt(np.maximum(deg, 1.0)), 0.0)
    vals = -(scale[row] * scale[col])
    diag = np.arange(n)
    return sp.coo_matrix((np.concatenate([vals, np.ones(n)]), (np.concatenate([row, diag]), np.concatenate([col, diag]))),
                         (n, n)).tocsr()


def _connected_graph(n: int, extra: int, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    path = np.arange(n - 1)
    src, dst = rng.integers(0, n, extra), rng.integers(0, n, extra)
    return np.stack([np.concatenate([src, path]), np.concatenate([dst, path + 1])])


def test_laplacian_block_operator_matches_scipy():
    from etpgt_b200.encodings.laplacian_pe import _SymLaplacian

    n = 3000
    ei = _connected_graph(n, 9000, 3)
    ei[:, :50] = ei[:, 50:100]            # parallel edges count once per direction
    ei[1, 100:120] = ei[0, 100:120]       # self loops are dropped
    lap = _scipy_laplacian(ei, n)
    op = _SymLaplacian(torch.from_numpy(ei), n, torch.device("cuda"))
    rng = np.random.default_rng(0)
    for b in (1, 17, 32):
        x, z = rng.standard_normal((n, b)), rng.standard_normal((n, b))
        got = op.apply(torch.from_numpy(x).cuda(), alpha=0.7, beta=-1.3, z=torch.from_numpy(z).cuda(), gamma=0.25)
        want = 0.7 * (lap @ x) - 1.3 * x + 0.25 * z
        assert np.abs(got.cpu().numpy() - want).max() < 1e-12
        plain = op.apply(torch.from_numpy(x).cuda())
        assert np.abs(plain.cpu().numpy() - lap @ x).max() < 1e-12


@pytest.mark.parametrize("n,extra,k", [(600, 2400, 16), (2500, 5000, 8)])
def test_device_laplacian_pe_matches_scipy_on_a_connected_graph(n, extra, k):
    """Distinct eigenvalues: |eigenvectors| are unique, so the device result must equal the reference's formula
    (eigsh(k+1, "SM"), drop the first, abs) on the same symmetric Laplacian."""
    from scipy.sparse.linalg import eigsh

    from etpgt_b200.encodings.laplacian_pe import compute_laplacian_pe, compute_laplacian_pe_device

    ei = _connected_graph(n, extra, n)
    w, v = eigsh(_scipy_laplacian(ei, n), k=k + 1, which="SM")
    order = np.argsort(w)
    w, v = w[order], v[:, order]
    pe, info = compute_laplacian_pe_device(torch.from_numpy(ei), n, k=k, return_info=True)
    assert pe.shape == (n, k) and pe.dtype == torch.float32 and pe.is_cuda
    assert np.abs(info["eigenvalues"].cpu().numpy() - w).max() < 1e-9
    assert float(info["residuals"].max()) < 1e-6
    assert np.abs(pe.cpu().numpy() - np.abs(v[:, 1:k + 1])).max() < 1e-5
    # symmetric input listed in both directions: the reference-faithful host path gives the same table
    both = np.concatenate([ei, ei[::-1]], axis=1)
    both = both[:, both[0] != both[1]]
    both = np.unique(both, axis=1)
    host = compute_laplacian_pe(torch.from_numpy(both), n, k=k)
    assert np.abs(host.numpy() - pe.cpu().numpy()).max() < 2e-4      # the host path runs eigsh in fp32


def test_device_laplacian_pe_on_the_co_occurrence_graph():
    """RR-shaped co-occurrence graph (several connected components, clustered small eigenvalues): every pair is an
    eigenpair (residual), the block is orthonormal, the zero eigenvalue comes with its multiplicity (one per
    connected component of two or more nodes — a block method finds them, single-vector Lanczos reports one), and
    every eigenvalue scipy finds is in the list."""
    import scipy.sparse.csgraph as csgraph
    from scipy.sparse.linalg import eigsh

    from etpgt_b200 import synth
    from etpgt_b200.encodings.laplacian_pe import _SymLaplacian, smallest_eigenpairs_device
    from etpgt_b200.model import create_graph_transformer_optimized

    data = synth.generate(num_sessions=20000, graph_sessions=15000, num_items=8000, clusters=160)
    ei = np.stack([data.item_i, data.item_j])
    n, k = data.num_items, 16
    lap = _scipy_laplacian(ei, n)
    op = _SymLaplacian(torch.from_numpy(ei), n, torch.device("cuda"))
    vals, vecs, res, _ = smallest_eigenpairs_device(op, k + 1)
    vals_h, vecs_h = vals.cpu().numpy(), vecs.cpu().numpy()
    assert float(res.max()) < 1e-6
    assert np.abs(lap @ vecs_h - vecs_h * vals_h).max() < 1e-6
    assert np.abs(vecs_h.T @ vecs_h - np.eye(k + 1)).max() < 1e-9
    assert np.all(np.diff(vals_h) >= -1e-12)
    adj = (lap != 0).astype(np.int8)
    _, labels = csgraph.connected_components(adj, directed=False)
    sizes = np.bincount(labels)
    zeros_expected = min(int((sizes >= 2).sum()), k + 1)
    assert int((np.abs(vals_h) < 1e-9).sum()) == zeros_expected
    w = np.sort(eigsh(lap, k=k + 1, which="SM")[0])
    for value in w[: k + 1 - zeros_expected]:
        assert np.abs(vals_h - value).min() < 1e-7
    # through the module API (laplacian_pe.py:156-168)
    model = create_graph_transformer_optimized(n, 64, 64).cuda()

    class Graph:
        edge_index, num_nodes = torch.from_numpy(ei), n

    model.laplacian_pe.precompute(Graph, method="device")
    pe = model.laplacian_pe._cached_pe
    assert pe.shape == (n, 16) and pe.is_cuda and bool((pe >= 0).all())
