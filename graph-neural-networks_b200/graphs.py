"""Sparse synthetic graphs for the measurement configs.

The reference generates graphs densely (`graphTools.createGraph`, alegnn/utils/graphTools.py:682-986: an N x N
`np.random.rand` plus a dense eigendecomposition for connectivity) which cannot run at N >= 1e5.  These generators
produce the same graph families directly in CSR:

  er_gso(N, avg_deg)            Erdős–Rényi G(N, M): N*avg_deg/2 undirected pairs drawn uniformly, self-loops and
                                duplicates dropped, symmetrised (the reference has no 'ER' type; ER is SBM with
                                probIntra == probInter, graphTools.py:747-800).
  sbm_gso(N, C, p_in, p_out)    stochastic block model with C equal communities (graphTools.py:761-798 semantics,
                                without the eig-based connectivity retry loop).

GSO normalisation: S = D^-1/2 A D^-1/2 (semantics of graphTools.normalizeAdjacency, graphTools.py:224-245), which
keeps the spectral radius at 1 like the examples' S = W / lambda_max (examples/sourceLocGNN.py:752) without a dense
eigendecomposition.  Randomness comes from a seeded numpy PCG64 generator on the host (device independent);
sorting / de-duplication runs in torch on the GPU when there is one.
"""
import numpy as np
import torch

from .gso import SparseGSO


def _dev():
    return torch.device("cuda") if torch.cuda.is_available() else torch.device("cpu")


def _unique_undirected(i, j, N):
    """int64 arrays of endpoints -> unique undirected pairs (lo < hi) as torch tensors on the work device."""
    dev = _dev()
    i = torch.from_numpy(i).to(dev)
    j = torch.from_numpy(j).to(dev)
    lo = torch.minimum(i, j)
    hi = torch.maximum(i, j)
    keep = lo != hi
    key = torch.unique(lo[keep] * N + hi[keep])
    return key // N, key % N


def _symmetric_csr(lo, hi, N):
    """unique undirected pairs -> CSR pattern (rowptr, col) of the symmetric adjacency, columns ascending."""
    rows = torch.cat([lo, hi])
    cols = torch.cat([hi, lo])
    order = torch.argsort(rows * N + cols)
    rows, cols = rows[order], cols[order]
    counts = torch.bincount(rows, minlength=N)
    rowptr = torch.zeros(N + 1, dtype=torch.int64, device=rows.device)
    rowptr[1:] = torch.cumsum(counts, 0)
    return rowptr, rows, cols, counts


def _finish(rowptr, rows, cols, counts, N, E, rng, dtype):
    deg = counts.to(torch.float64).clamp_(min=1.0)
    dinv = deg.rsqrt()
    base = dinv[rows] * dinv[cols]
    npd = np.float32 if dtype == torch.float32 else np.float64
    rp = rowptr.cpu().numpy()
    ci = cols.to(torch.int32).cpu().numpy()
    csr = []
    for e in range(E):
        if e == 0 and E == 1:
            val = base
        else:
            # tensor GSO (cfg4): same pattern, independent U(0,1) edge weights, then degree-normalised
            w = torch.from_numpy(rng.random(rows.numel())).to(base.device)
            val = base * w * 2.0
        csr.append((rp, ci, val.cpu().numpy().astype(npd)))
    return SparseGSO(csr, N)


def er_gso(N, avg_deg, seed=0, E=1, dtype=torch.float32):
    rng = np.random.Generator(np.random.PCG64(seed))
    M = int(N * avg_deg // 2)
    i = rng.integers(0, N, size=M, dtype=np.int64)
    j = rng.integers(0, N, size=M, dtype=np.int64)
    lo, hi = _unique_undirected(i, j, N)
    return _finish(*_symmetric_csr(lo, hi, N), N, E, rng, dtype)


def sbm_gso(N, n_communities, p_intra, p_inter, seed=0, E=1, dtype=torch.float32):
    """Stochastic block model with (near-)equal communities of consecutive node ids.  Vectorised sampling: per
    community a Binomial(n(n-1)/2, p_intra) number of uniformly drawn inside pairs; a Binomial(#outside pairs, p_inter)
    number of uniformly drawn pairs with endpoints in different communities (duplicates / self-loops dropped)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    bounds = np.linspace(0, N, n_communities + 1).astype(np.int64)
    sizes = np.diff(bounds)
    comm_of = np.repeat(np.arange(n_communities), sizes)
    m_in = rng.binomial(sizes * (sizes - 1) // 2, min(1.0, p_intra))
    cid = np.repeat(np.arange(n_communities), m_in)
    M_in = int(m_in.sum())
    i_in = bounds[cid] + (rng.random(M_in) * sizes[cid]).astype(np.int64)
    j_in = bounds[cid] + (rng.random(M_in) * sizes[cid]).astype(np.int64)
    out_pairs = (N * (N - 1) // 2) - int((sizes * (sizes - 1) // 2).sum())
    M_out = int(rng.binomial(out_pairs, min(1.0, p_inter))) if out_pairs > 0 else 0
    # draw a little more than needed, keep the pairs that cross communities
    draw = int(M_out * (1.0 + 2.0 / max(n_communities, 2))) + 16
    io = rng.integers(0, N, size=draw, dtype=np.int64)
    jo = rng.integers(0, N, size=draw, dtype=np.int64)
    keep = comm_of[io] != comm_of[jo]
    io, jo = io[keep][:M_out], jo[keep][:M_out]
    i = np.concatenate([i_in, io])
    j = np.concatenate([j_in, jo])
    lo, hi = _unique_undirected(i, j, N)
    return _finish(*_symmetric_csr(lo, hi, N), N, E, rng, dtype)


def knn_like_gso(N, k, seed=0, dtype=torch.float32):
    """MovieLens-shaped (cfg3): every node picks k random neighbours, then symmetrised (degree ~ k .. 2k),
    standing in for the k-NN similarity graph of examples/movieGNN.py:155."""
    rng = np.random.Generator(np.random.PCG64(seed))
    i = np.repeat(np.arange(N, dtype=np.int64), k)
    j = rng.integers(0, N, size=N * k, dtype=np.int64)
    lo, hi = _unique_undirected(i, j, N)
    return _finish(*_symmetric_csr(lo, hi, N), N, 1, rng, dtype)


def degrees(gso, e=0):
    return np.diff(gso.csr[e][0])
