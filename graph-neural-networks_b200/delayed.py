"""Batch- and time-varying graph filter on the B200 LSIGF path.

    LSIGF_DB(h, S, x, b=None)                   <- alegnn/utils/graphML.py:977-1094
    GraphFilter_DB(G, F, K, E=1, bias=True)     <- graphML.py:3278-3393

    y_f(b, t) = sum_e sum_k sum_g h[f,e,k,g]  x_g(b, t-k) S_e(b, t-k+1) ... S_e(b, t)  +  bias_f
    (row-vector convention, signals before t = 0 are zero: graphML.py:1060-1075)

The reference runs B*T small dense N x N products per tap (N = 50 agents in the flocking examples).  Here the whole
batch is ONE sparse graph filter: stack the B*T copies of the node set into M = B*T*N "space-time" nodes and put
S_e(b, t) on the block that links copy (b, t-1) to copy (b, t):

        S_big_e[(b, t-1, i), (b, t, j)] = S_e(b, t)[i, j]          (zero elsewhere; nothing enters t = 0)

Then (x S_big^k)[(b,t,:)] = x(b, t-k) S(b, t-k+1) ... S(b, t) is exactly the delayed shift, the unit delay and the
zero history included, and LSIGF_DB(h, S, x, b) = LSIGF(h, S_big, x_big, b_big) with a single "sample" of M nodes.
One plan per GSO batch, K-1 launches of the same CSR hop kernel over all B*T graphs at once, one contraction — the
same CUDA path (and C ABI) as the static filter, instead of (K-1) batched GEMM launches plus permute copies.
"""
import math
import weakref

import torch
import torch.nn as nn

from . import graphML as _gml
from .gso import Plan

_CACHE = {}
_CACHE_MAX = 4


def block_delay_csr(S):
    """S [B, T, E, N, N] (pattern = S != 0) -> ([(rowptr int64 [M+1], col int32 [nnz], val [nnz]) for e], M = B*T*N):
    CSR (row i lists S_big_e[i, j], columns ascending) of the space-time operator described in the module docstring."""
    assert S.dim() == 5 and S.shape[3] == S.shape[4]
    B, T, E, N, _ = S.shape
    M = B * T * N
    assert M < 2 ** 31, "b200gf: B*T*N must fit the int32 column index"
    out = []
    for e in range(E):
        rowptr = torch.zeros(M + 1, dtype=torch.int64, device=S.device)
        if T > 1:
            Se = S[:, 1:, e]                                       # S(b, t) for t = 1 .. T-1  [B, T-1, N, N]
            nz = (Se != 0).nonzero(as_tuple=False)                 # (b, t-1, i, j), row-major = sorted by (row, col)
            bb, tt, ii, jj = nz.unbind(1)
            rows = (bb * T + tt) * N + ii                          # source copy (b, t-1)
            cols = (bb * T + tt + 1) * N + jj                      # destination copy (b, t)
            vals = Se[bb, tt, ii, jj].contiguous()
            rowptr[1:] = torch.cumsum(torch.bincount(rows, minlength=M), 0)
        else:
            cols = torch.zeros(0, dtype=torch.int64, device=S.device)
            vals = torch.zeros(0, dtype=S.dtype, device=S.device)
        out.append((rowptr, cols.to(torch.int32).contiguous(), vals))
    return out, M


def transpose_csr_device(csr, M):
    """CSR (rowptr, col, val) of an M x M operator -> CSR of its transpose, on the same device: one stable sort by column
    (entries of an output row keep ascending source-row order, like the host builder's counting sort)."""
    rowptr, col, val = csr
    rows = torch.repeat_interleave(torch.arange(M, device=rowptr.device), rowptr[1:] - rowptr[:-1])
    order = torch.sort(col.to(torch.int64), stable=True)[1]
    t_rowptr = torch.zeros(M + 1, dtype=torch.int64, device=rowptr.device)
    t_rowptr[1:] = torch.cumsum(torch.bincount(col.to(torch.int64), minlength=M), 0)
    return t_rowptr, rows[order].to(torch.int32).contiguous(), val[order].contiguous()


def _plan_for_batch(S):
    """Plan of the space-time operator, cached per (storage, version): GraphFilter_DB.addGSO is called once per batch
    (architecturesTime.py) and every layer of the network shares that GSO tensor."""
    if S.requires_grad:
        raise NotImplementedError("b200gf: gradients w.r.t. the GSO are not part of the LSIGF path")
    if S.device.type != "cuda":
        raise RuntimeError("b200gf: LSIGF_DB needs CUDA tensors (there is no CPU fallback); got GSO on %s" % S.device)
    key = (S.data_ptr(), S._version, tuple(S.shape), tuple(S.stride()), S.dtype, str(S.device))
    for k in [k for k, h in _CACHE.items() if h[0]() is None]:     # GSO batch already collected: free its plan now
        del _CACHE[k]
    hit = _CACHE.get(key)
    if hit is not None and hit[0]() is S:
        return hit[1]
    csr, M = block_delay_csr(S)
    # both operators are built on the device (one sort for the transpose) and adopted device to device: no host round trip
    plan = Plan.from_device_ops([transpose_csr_device(c, M) for c in csr], csr, M, S.dtype, S.device)
    if len(_CACHE) >= _CACHE_MAX:
        _CACHE.pop(next(iter(_CACHE)))
    _CACHE[key] = (weakref.ref(S), plan)
    return plan


def _filter_on_space_time_graph(h, S, x_big, b_big):
    """LSIGF over the M-node operator of `S`; x_big [1, G, M] (node-major view), b_big None / [F, 1] / [F, M]."""
    if x_big.device.type != "cuda":
        raise RuntimeError("b200gf: LSIGF_DB needs CUDA tensors (there is no CPU fallback); got x on %s" % x_big.device)
    if x_big.dtype not in (torch.float32, torch.float64):
        raise RuntimeError("b200gf: LSIGF_DB supports float32 and float64, got %s" % x_big.dtype)
    if S.device != x_big.device:
        raise RuntimeError("b200gf: GSO on %s but x on %s" % (S.device, x_big.device))
    return _gml._LSIGFFunction.apply(h, x_big, b_big, _plan_for_batch(S))


# tests swap this for the CPU oracle applied to block_delay_csr(S) to check the host logic without a GPU
_apply = _filter_on_space_time_graph


def LSIGF_DB(h, S, x, b=None):
    """LSIGF_DB(filter_taps, GSO, input, bias=None)   (graphML.py:977-1094)

    h [F, E, K, G]; S [B, T, E, N, N]; x [B, T, G, N]; b [F, 1] or [F, N] or None  ->  y [B, T, F, N]."""
    assert len(h.shape) == 4
    F = h.shape[0]
    E = h.shape[1]
    G = h.shape[3]
    assert len(S.shape) == 5
    B = S.shape[0]
    T = S.shape[1]
    assert S.shape[2] == E
    N = S.shape[3]
    assert S.shape[4] == N
    assert len(x.shape) == 4
    assert x.shape[0] == B
    assert x.shape[1] == T
    assert x.shape[2] == G
    assert x.shape[3] == N
    if h.dtype != x.dtype or S.dtype != x.dtype or (b is not None and b.dtype != x.dtype):
        raise RuntimeError("b200gf: LSIGF_DB expects h, S, x, b of one dtype, got h=%s S=%s x=%s" % (h.dtype, S.dtype, x.dtype))
    M = B * T * N
    # space-time node-major input [M, G] (node index (b, t, n)), handed to the filter as a [1, G, M] view of it
    x_big = x.permute(0, 1, 3, 2).reshape(M, G).t().unsqueeze(0)
    b_big = b
    if b is not None:
        assert b.dim() == 2 and b.shape[0] == F and b.shape[1] in (1, N)
        if b.shape[1] == N and M != N:
            b_big = b.repeat(1, B * T)                              # per-node bias, the same for every (b, t)
    y_big = _apply(h, S, x_big, b_big)                              # [1, F, M]
    return y_big[0].reshape(F, B, T, N).permute(1, 2, 0, 3)         # [B, T, F, N], still a view of the node-major buffer


class GraphFilter_DB(nn.Module):
    """GraphFilter_DB(in_features, out_features, filter_taps, edge_features=1, bias=True)

    Same surface as graphML.py:3278-3393: parameters `weight` [F, E, K, G], `bias` [F, 1] or None;
    addGSO(S [B, T, E, N, N]); forward(x [B, T, G, N]) -> [B, T, F, N]."""

    def __init__(self, G, F, K, E=1, bias=True):
        super().__init__()
        self.G = G
        self.F = F
        self.K = K
        self.E = E
        self.S = None
        self.weight = nn.parameter.Parameter(torch.Tensor(F, E, K, G))
        if bias:
            self.bias = nn.parameter.Parameter(torch.Tensor(F, 1))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self):
        stdv = 1. / math.sqrt(self.G * self.K)      # graphML.py:3353-3358
        self.weight.data.uniform_(-stdv, stdv)
        if self.bias is not None:
            self.bias.data.uniform_(-stdv, stdv)

    def addGSO(self, S):
        assert len(S.shape) == 5                    # graphML.py:3362
        assert S.shape[2] == self.E
        self.N = S.shape[3]
        assert S.shape[4] == self.N
        self.S = S

    def forward(self, x):
        assert len(x.shape) == 4
        B = x.shape[0]
        assert self.S.shape[0] == B
        T = x.shape[1]
        assert self.S.shape[1] == T
        assert x.shape[3] == self.N
        return LSIGF_DB(self.weight, self.S, x, self.bias)

    def extra_repr(self):
        reprString = "in_features=%d, out_features=%d, " % (self.G, self.F) + "filter_taps=%d, " % (self.K) + \
                     "edge_features=%d, " % (self.E) + "bias=%s, " % (self.bias is not None)
        if self.S is not None:
            reprString += "GSO stored"
        else:
            reprString += "no GSO stored"
        return reprString
