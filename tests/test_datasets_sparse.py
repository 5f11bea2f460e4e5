"""Sparse source-localization data set (gnn_b200.datasets_sparse; reference: alegnn/utils/dataTools.py:472-592).
Equality with the reference class itself is checked live in tests/test_oracle_vs_reference.py; here: the definition
x = (W / lambda_max)^t delta_source on a small graph, and a graph far beyond what dense matrix powers could hold."""
import numpy as np
import scipy.sparse as sp
import torch


def test_signals_follow_the_definition():
    from gnn_b200 import datasets_sparse
    from gnn_b200.graphtools_sparse import largest_real_eigenvalue
    rng = np.random.default_rng(0)
    N = 40
    A = (rng.random((N, N)) < 0.15).astype(float)
    W = np.triu(A, 1)
    W = sp.csr_matrix(W + W.T)
    Wn = (W / largest_real_eigenvalue(W)).toarray()
    sources = np.array([3, 17, 3, 29, 17, 3])
    times = np.array([0, 4, 2, 7, 4, 0])
    x = datasets_sparse.diffusion_signals(W, sources, times)
    for i, (s, t) in enumerate(zip(sources, times)):
        ref = np.linalg.matrix_power(Wn, int(t))[:, s]
        assert np.abs(x[i] - ref).max() < 1e-13
    assert np.array_equal(x[0], np.eye(N)[3])                      # t = 0: the delta itself


def test_large_graph_dataset_and_api():
    from gnn_b200 import datasets_sparse, graphs
    gso = graphs.er_gso(200_000, 8, seed=2, dtype=torch.float64)
    r, c, v = gso.csr[0]
    W = sp.csr_matrix((np.ones_like(v), c, r), shape=(gso.N, gso.N))     # unweighted adjacency, 200k x 200k
    np.random.seed(1)
    sources = [5, 1000, 150_000]
    data = datasets_sparse.SourceLocalization(W, 12, 4, 4, sources, tMax=6, dataType=torch.float32)
    assert data.samples["train"]["signals"].shape == (12, 200_000) and data.samples["train"]["signals"].dtype == torch.float32
    assert data.samples["test"]["targets"].dtype == torch.int64
    assert set(data.samples["train"]["targets"].tolist()) <= {0, 1, 2}
    data.expandDims()
    x, y = data.getSamples("train", 3)
    assert tuple(x.shape) == (3, 1, 200_000) and tuple(y.shape) == (3,)
    x1, _ = data.getSamples("valid", np.int64(2))                   # a single (non-int-typed) index keeps the sample axis
    assert tuple(x1.shape) == (1, 1, 200_000)
    # every signal is a non-negative diffusion whose mass starts on its source node
    s = data.samples["train"]["signals"]
    assert float(s.min()) >= 0.0 and bool((s.sum(dim=(1, 2)) > 0).all())
    scores = torch.zeros(4, 3)
    scores[torch.arange(4), data.samples["test"]["targets"]] = 1.0
    assert float(data.evaluate(scores, data.samples["test"]["targets"])) == 0.0
