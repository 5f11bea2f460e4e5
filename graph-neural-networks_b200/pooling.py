"""Local max-pooling over K-hop neighbourhoods for graphs the reference cannot handle (host-side part).

    MaxPoolLocal(in_dim, out_dim, number_hops)          <- alegnn/utils/graphML.py:1850-2028

Same surface and results as the reference layer.  Differences are internal: `addGSO` also accepts a `SparseGSO`, the
neighbourhoods come from the CSR routine (`graphtools_sparse.compute_neighborhood`) instead of a pure-python search over a
dense N x N matrix (graphML.py:1953-1957 -> graphTools.py:459-500, O(N deg^K) python), and `forward` gathers only the
neighbours (`index_select`) instead of first repeating the whole signal `maxNeighborhoodSize` times (graphML.py:1990-1996).
Pure PyTorch: runs wherever its input lives, including on the permuted node-major views LSIGF returns.
"""
import numpy as np
import scipy.sparse as sp
import torch
import torch.nn as nn

from .graphtools_sparse import compute_neighborhood
from .gso import SparseGSO


class MaxPoolLocal(nn.Module):
    def __init__(self, nInputNodes, nOutputNodes, nHops):
        super().__init__()
        self.nInputNodes = nInputNodes
        self.nOutputNodes = nOutputNodes
        self.nHops = nHops
        self.neighborhood = None

    def addGSO(self, S):
        assert len(S.shape) == 3                     # graphML.py:1944
        self.N = S.shape[1]
        assert S.shape[2] == self.N
        if isinstance(S, SparseGSO):
            device = torch.device("cpu")
            mats = [sp.csr_matrix((v, c, r), shape=(S.N, S.N)) for (r, c, v) in S.csr]
        else:
            device = S.device
            Sc = S.detach().cpu()
            mats = []
            for e in range(Sc.shape[0]):
                nz = Sc[e].nonzero(as_tuple=False).numpy()
                vals = Sc[e][nz[:, 0], nz[:, 1]].numpy().astype(np.float64)
                mats.append(sp.csr_matrix((vals, (nz[:, 0], nz[:, 1])), shape=(self.N, self.N)))
        nb = compute_neighborhood(mats if len(mats) > 1 else mats[0], self.nHops, self.nOutputNodes,
                                  self.nInputNodes, "matrix")
        neighborhood = torch.tensor(nb, dtype=torch.int64, device=device)
        assert neighborhood.shape[0] == self.nOutputNodes
        assert neighborhood.numel() == 0 or neighborhood.max() <= self.nInputNodes
        self.maxNeighborhoodSize = neighborhood.shape[1]
        self.neighborhood = neighborhood

    def forward(self, x):
        B, F, Nin = x.shape
        assert Nin == self.nInputNodes               # graphML.py:1977
        assert Nin >= self.nOutputNodes
        nb = self.neighborhood
        if nb.device != x.device:
            nb = self.neighborhood = nb.to(x.device)
        xn = x.index_select(2, nb.reshape(-1)).reshape(B, F, self.nOutputNodes, self.maxNeighborhoodSize)
        v, _ = torch.max(xn, dim=3)
        return v

    def extra_repr(self):
        s = "in_dim=%d, out_dim=%d, number_hops = %d, " % (self.nInputNodes, self.nOutputNodes, self.nHops)
        return s + ("neighborhood stored" if self.neighborhood is not None else "NO neighborhood stored")
