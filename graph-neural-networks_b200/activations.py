"""Localized activation functions on K-hop neighbourhoods, built on the CSR neighbourhood routine (SURVEY.md §2.1 row 6,
a "next" candidate of §8f).

    MaxLocalActivation(K)        <- alegnn/utils/graphML.py:1535-1660
    MedianLocalActivation(K)     <- alegnn/utils/graphML.py:1662-1810

Same surface (parameter `weight [1, K+1]`, `addGSO`, `forward`, `reset_parameters`) and the same results as the reference
layers: out = sum_{k=0..K} w_k * agg_k(x), agg_0 = x, agg_k = max / median over the k-hop neighbourhood of each node.
Internal differences: neighbourhoods come from `graphtools_sparse.compute_neighborhood` (accepts a SparseGSO, no dense
N x N matrix, no pure-python search); the max gathers only the neighbours instead of repeating the signal; the median
is one masked sort per hop instead of a python loop over the N nodes with a `torch.cat` per node (graphML.py:1778-1795).
Pure PyTorch (host code around the filter path).
"""
import math

import numpy as np
import scipy.sparse as sp
import torch
import torch.nn as nn

from .graphtools_sparse import compute_neighborhood
from .gso import SparseGSO


def _scipy_mats(S):
    if isinstance(S, SparseGSO):
        return [sp.csr_matrix((v, c, r), shape=(S.N, S.N)) for (r, c, v) in S.csr], torch.device("cpu")
    Sc = S.detach().cpu()
    mats = []
    for e in range(Sc.shape[0]):
        nz = Sc[e].nonzero(as_tuple=False).numpy()
        vals = Sc[e][nz[:, 0], nz[:, 1]].numpy().astype(np.float64)
        mats.append(sp.csr_matrix((vals, (nz[:, 0], nz[:, 1])), shape=(Sc.shape[1], Sc.shape[1])))
    return mats, S.device


class _LocalActivation(nn.Module):
    def __init__(self, K):
        super().__init__()
        assert K > 0                                   # graphML.py:1582
        self.K = K
        self.S = None
        self.N = None
        self.neighborhood = "None"
        self.weight = nn.parameter.Parameter(torch.Tensor(1, self.K + 1))
        self.reset_parameters()

    def reset_parameters(self):
        stdv = 1. / math.sqrt(self.K)                  # graphML.py:1651
        self.weight.data.uniform_(-stdv, stdv)

    def _neighbourhood_matrices(self, S):
        """per hop k = 1..K: (index matrix [N, M_k] padded with the node itself, lengths [N])"""
        mats, device = _scipy_mats(S)
        arg = mats if len(mats) > 1 else mats[0]
        out = []
        for k in range(1, self.K + 1):
            lists = compute_neighborhood(arg, k, outputType="list")
            lens = torch.tensor([len(x) for x in lists], dtype=torch.int64)
            width = int(lens.max()) if len(lists) else 0
            idx = torch.arange(len(lists), dtype=torch.int64).unsqueeze(1).repeat(1, max(width, 1))[:, :width].clone()
            for i, x in enumerate(lists):
                idx[i, :len(x)] = torch.tensor(x, dtype=torch.int64)
            out.append((idx.to(device), lens.to(device)))
        return out

    def addGSO(self, S):
        assert len(S.shape) == 3                       # graphML.py:1594
        self.N = S.shape[1]
        assert S.shape[2] == self.N
        self.S = S
        self.neighborhood = self._neighbourhood_matrices(S)
        self.maxNeighborhoodSizes = [m.shape[1] for (m, _) in self.neighborhood]

    def _combine(self, aggs):
        xK = torch.stack(aggs, dim=3)                                   # [B, F, N, K+1]
        out = torch.matmul(xK, self.weight.unsqueeze(2))                # graphML.py:1644
        return out.reshape(xK.shape[0], xK.shape[1], self.N)

    def extra_repr(self):
        return "neighborhood stored" if self.neighborhood is not None else "NO neighborhood stored"


class MaxLocalActivation(_LocalActivation):
    def forward(self, x):
        B, F, N = x.shape
        assert N == self.N                              # graphML.py:1617
        aggs = [x]
        for (idx, _) in self.neighborhood:
            idx = idx.to(x.device)
            v, _ = torch.max(x.index_select(2, idx.reshape(-1)).reshape(B, F, N, idx.shape[1]), dim=3)
            aggs.append(v)
        return self._combine(aggs)


class MedianLocalActivation(_LocalActivation):
    def forward(self, x):
        B, F, N = x.shape
        assert N == self.N                              # graphML.py:1768
        aggs = [x]
        for (idx, lens) in self.neighborhood:
            idx, lens = idx.to(x.device), lens.to(x.device)
            M = idx.shape[1]
            xn = x.index_select(2, idx.reshape(-1)).reshape(B, F, N, M)
            pad = torch.arange(M, device=x.device).unsqueeze(0) >= lens.unsqueeze(1)          # [N, M] padded slots
            xs, _ = torch.sort(xn.masked_fill(pad.reshape(1, 1, N, M), float("inf")), dim=3)
            pick = ((lens - 1) // 2).reshape(1, 1, N, 1).expand(B, F, N, 1)                  # torch.median: lower median
            aggs.append(torch.gather(xs, 3, pick).squeeze(3))
        return self._combine(aggs)


class NoActivation(nn.Module):
    """graphML.py:1812-1848"""

    def forward(self, x):
        return x

    def extra_repr(self):
        return "No Activation Function"
