"""CSR-native graph utilities (graph-neural-networks_b200/graphtools_sparse.py) against the UNMODIFIED reference's dense
`alegnn.utils.graphTools` (live, CPU; skipped where /root/reference does not exist), plus reference-free properties."""
import numpy as np
import pytest
import scipy.sparse as sp

import ref_import

needs_ref = pytest.mark.skipif(not ref_import.reference_available(), reason="reference tree not present")


def _graph(seed, N=24, p=0.18, directed=False, weights=True):
    rng = np.random.default_rng(seed)
    A = (rng.random((N, N)) < p).astype(float)
    np.fill_diagonal(A, 0.0)
    if weights:
        A = A * rng.uniform(0.2, 1.5, (N, N))
    if not directed:
        A = np.triu(A, 1)
        A = A + A.T
    ring = np.roll(np.eye(N), 1, axis=1) * 0.7       # a ring keeps every graph connected and every degree positive
    return A + (ring if directed else ring + ring.T)


@pytest.fixture(scope="module")
def gt():
    ref_import.import_reference()
    import alegnn.utils.graphTools as graphTools
    return graphTools


@pytest.fixture(scope="module")
def gs():
    import gnn_b200  # noqa: F401
    import gnn_b200.graphtools_sparse as g
    return g


@needs_ref
@pytest.mark.parametrize("seed,directed", [(0, False), (1, False), (2, True), (3, True)])
def test_normalisations_and_spectrum(gt, gs, seed, directed):
    W = _graph(seed, directed=directed)
    Ws = sp.csr_matrix(W)
    assert np.allclose(gs.adjacency_to_laplacian(Ws).toarray(), gt.adjacencyToLaplacian(W), atol=1e-13)
    assert np.allclose(gs.normalize_adjacency(Ws).toarray(), gt.normalizeAdjacency(W), atol=1e-13)
    L = gt.adjacencyToLaplacian(W)
    assert np.allclose(gs.normalize_laplacian(sp.csr_matrix(L)).toarray(), gt.normalizeLaplacian(L), atol=1e-13)
    E, _ = gt.computeGFT(W)                                       # what the examples divide by (sourceLocGNN.py:752)
    lam = np.max(np.real(np.diag(E)))
    assert abs(gs.largest_real_eigenvalue(Ws) - lam) < 1e-8 * abs(lam)
    assert np.allclose(gs.spectral_normalize(Ws).toarray(), W / lam, atol=1e-8)


@needs_ref
def test_connectivity(gt, gs):
    for seed in range(4):
        W = _graph(seed, directed=seed % 2 == 1)
        assert gs.is_connected(sp.csr_matrix(W)) == gt.isConnected(W) is True
    two = np.zeros((10, 10))
    two[:5, :5] = _graph(7, N=5)
    two[5:, 5:] = _graph(8, N=5)
    assert gs.is_connected(sp.csr_matrix(two)) == gt.isConnected(two) is False
    one_way = np.diag(np.ones(5), 1)                             # a directed path counts as connected (:570-574)
    assert gs.is_connected(sp.csr_matrix(one_way)) == gt.isConnected(one_way) is True


@needs_ref
@pytest.mark.parametrize("K", [0, 1, 2, 3])
@pytest.mark.parametrize("seed,directed", [(0, False), (5, True)])
def test_neighbourhoods(gt, gs, K, seed, directed):
    W = _graph(seed, N=20, p=0.1, directed=directed)
    ref = gt.computeNeighborhood(W, K)
    got = gs.compute_neighborhood(sp.csr_matrix(W), K)
    assert [sorted(int(j) for j in r) for r in ref] == got
    # first N nodes only, neighbours restricted to nodes < nb, matrix output padded with the node's own index
    ref_m = gt.computeNeighborhood(W, K, N=7, nb=15, outputType="matrix")
    got_m = gs.compute_neighborhood(sp.csr_matrix(W), K, N=7, nb=15, outputType="matrix")
    assert ref_m.shape == got_m.shape
    assert [sorted(set(r.tolist())) for r in ref_m] == [sorted(set(r.tolist())) for r in got_m]
    # edge-feature GSO: an edge exists where any S_e is non-zero (:424-432)
    S3 = np.stack([W * (np.arange(20)[:, None] % 2 == 0), W * (np.arange(20)[:, None] % 2 == 1)])
    ref3 = gt.computeNeighborhood(S3, K)
    got3 = gs.compute_neighborhood([sp.csr_matrix(S3[0]), sp.csr_matrix(S3[1])], K)
    assert [sorted(int(j) for j in r) for r in ref3] == got3


@needs_ref
def test_perm_degree(gt, gs):
    W = _graph(11, directed=True)
    refS, refOrder = gt.permDegree(W)
    gotS, gotOrder = gs.perm_degree(sp.csr_matrix(W))
    assert refOrder == gotOrder and np.array_equal(gotS.toarray(), refS)
    S3 = np.stack([W, W.T * 0.5])
    ref3, order3 = gt.permDegree(S3)
    got3, gorder3 = gs.perm_degree([sp.csr_matrix(S3[0]), sp.csr_matrix(S3[1])])
    assert order3 == gorder3 and np.array_equal(np.stack([m.toarray() for m in got3]), ref3)


@needs_ref
@pytest.mark.parametrize("directed", [False, True])
def test_edge_fail_sampling_matches_reference_rng(gt, gs, directed):
    W = _graph(13, directed=directed)
    np.random.seed(5)
    ref = gt.edgeFailSampling(W, 0.3)
    np.random.seed(5)
    got = gs.edge_fail_sampling(sp.csr_matrix(W), 0.3, dense_rng_compat=True)
    assert np.array_equal(got.toarray(), ref)


def test_edge_fail_sampling_scalable_mode(gs):
    W = sp.csr_matrix(_graph(17, N=200, p=0.05))
    out = gs.edge_fail_sampling(W, 0.4, rng=np.random.default_rng(1))
    assert abs(out - out.T).max() == 0                            # stays undirected
    assert (out != 0).multiply(W == 0).nnz == 0                   # never creates an edge
    frac = out.nnz / W.nnz
    assert 0.5 < frac < 0.7                                       # ~60 % of the edges survive
    assert gs.edge_fail_sampling(W, 0.0, rng=np.random.default_rng(1)).nnz == W.nnz


@needs_ref
@pytest.mark.parametrize("kind,p", [("threshold", 0.6), ("threshold", 1.2), ("NN", 3), ("NN", 1)])
@pytest.mark.parametrize("directed", [False, True])
def test_sparsify(gt, gs, kind, p, directed):
    W = _graph(19, N=18, p=0.35, directed=directed)
    ref = gt.sparsifyGraph(W, kind, p)
    got = gs.sparsify_graph(sp.csr_matrix(W), kind, p)
    assert np.allclose(got.toarray(), ref, atol=1e-13)


def test_large_graph_pipeline_feeds_the_filter(gs):
    """The sparse pipeline at a size the dense reference cannot hold: build, check, normalise, wrap as a SparseGSO."""
    import gnn_b200
    from gnn_b200 import graphs
    g = graphs.er_gso(200_000, 8, seed=3)
    r, c, v = g.csr[0]
    A = sp.csr_matrix((np.ones_like(v, dtype=np.float64), c, r), shape=(g.N, g.N))
    assert gs.is_connected(A) in (True, False)
    S = gs.spectral_normalize(A)
    assert abs(gs.largest_real_eigenvalue(S) - 1.0) < 1e-6
    nb = gs.compute_neighborhood(A, 2, N=5)
    assert all(i in nb[i] for i in range(5))
    gso = gs.to_sparse_gso(S, dtype="torch.float32")
    assert gso.shape == (1, 200_000, 200_000) and gso.nnz() == A.nnz
