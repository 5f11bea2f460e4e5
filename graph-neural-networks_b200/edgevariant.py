"""Edge-variant graph filter (the variant row of the path), executed sparsely on the GPU.

    EVGF(S, x, b=None)                                   <- alegnn/utils/graphML.py:389-488
    EdgeVariantGF(G, F, K, M, N, E=1, bias=True)         <- alegnn/utils/graphML.py:2511-2712

The reference multiplies dense N x N filter matrices Phi^(k)_{f e g}; only the entries on the masked sparsity pattern
of |S_e| + I are ever non-zero (`sparsityPatternFull`, graphML.py:2620-2663).  Here those entries are gathered once per
call into per-non-zero weights on a compact node set and the K-step chains run in `b200gf_ev_forward/backward`
(csrc/ev.cu).  Parameter names and shapes (`weightEV [F,E,K,G,N,N]`, `weightLSI [F,E,K,G]`, `bias [F,1]`) are the
reference's, so its checkpoints load; the hybrid layer's LSI part goes through this package's LSIGF.
Reference quirks kept on purpose: the hybrid layer adds the bias twice (once in EVGF :2682, once in LSIGF :2686).
"""
import math

import torch
import torch.nn as nn

from . import _cabi
from .graphML import LSIGF

zeroTolerance = 1e-9  # graphML.py:72
_ENUM = {torch.float32: _cabi.F32, torch.float64: _cabi.F64}


class EVStructure:
    """Compact CSR patterns (one per edge feature) for a masked sparsity pattern [E, N, N] (bool)."""

    def __init__(self, pattern):
        assert pattern.dim() == 3 and pattern.shape[1] == pattern.shape[2]
        dev = pattern.device
        E, N, _ = pattern.shape
        self.E, self.N = E, N
        nz = [pattern[e].nonzero(as_tuple=False) for e in range(E)]           # row-major order
        touched = torch.zeros(N, dtype=torch.bool, device=dev)
        for z in nz:
            touched[z[:, 0]] = True
            touched[z[:, 1]] = True
        self.A = touched.nonzero(as_tuple=False).flatten()                     # compact set, ascending global ids
        self.NA = int(self.A.numel())
        inv = torch.full((N,), -1, dtype=torch.int64, device=dev)
        inv[self.A] = torch.arange(self.NA, device=dev)
        self.per_e = []
        for z in nz:
            ci, cj = inv[z[:, 0]], inv[z[:, 1]]
            nnz = int(ci.numel())
            rowptr = torch.zeros(self.NA + 1, dtype=torch.int64, device=dev)
            rowptr[1:] = torch.cumsum(torch.bincount(ci, minlength=self.NA), 0)
            permT = torch.argsort(cj * max(self.NA, 1) + ci)                   # entries ordered by (column, row)
            rowptrT = torch.zeros(self.NA + 1, dtype=torch.int64, device=dev)
            rowptrT[1:] = torch.cumsum(torch.bincount(cj, minlength=self.NA), 0)
            self.per_e.append(dict(
                nnz=nnz, rowptr=rowptr, col=cj.to(torch.int32).contiguous(), rowidx=ci.to(torch.int32).contiguous(),
                rowptrT=rowptrT, colT=ci[permT].to(torch.int32).contiguous(), perm=permT.contiguous(),
                lin=(z[:, 0] * N + z[:, 1]).contiguous()))                      # flat index into a dense [N, N] matrix


class _EVChain(torch.autograd.Function):
    """y_e[b, f, i in A] = sum_g sum_k (Phi^(k) ... Phi^(0) x_g)[i] for one edge feature."""

    @staticmethod
    def forward(ctx, w, xA, pe, NA):
        lib = _cabi.load()
        F_, K, G, nnz = w.shape
        B = xA.shape[0]
        w = w.contiguous()
        xA = xA.contiguous()
        st = torch.cuda.current_stream().cuda_stream
        states = torch.empty((K, F_ * G * B, NA), dtype=w.dtype, device=w.device)
        S = torch.empty((F_, G, B, NA), dtype=w.dtype, device=w.device)
        _cabi.check(lib.b200gf_ev_forward(_ENUM[w.dtype], NA, B, G, F_, K, pe["rowptr"].data_ptr(), pe["col"].data_ptr(), nnz,
                                          w.data_ptr(), xA.data_ptr(), states.data_ptr(), S.data_ptr(), st))
        ctx.pe, ctx.NA = pe, NA
        ctx.save_for_backward(w, xA, states)
        return S.sum(dim=1).permute(1, 0, 2).contiguous()                       # [B, F, NA]

    @staticmethod
    def backward(ctx, dy):
        lib = _cabi.load()
        w, xA, states = ctx.saved_tensors
        pe, NA = ctx.pe, ctx.NA
        F_, K, G, nnz = w.shape
        B = xA.shape[0]
        dy = dy.contiguous()
        st = torch.cuda.current_stream().cuda_stream
        lam = torch.empty((2, F_ * G * B, NA), dtype=w.dtype, device=w.device)
        dw = torch.zeros_like(w)
        dxA = torch.empty_like(xA)
        _cabi.check(lib.b200gf_ev_backward(_ENUM[w.dtype], NA, B, G, F_, K, pe["rowidx"].data_ptr(), pe["col"].data_ptr(),
                                           pe["rowptrT"].data_ptr(), pe["colT"].data_ptr(), pe["perm"].data_ptr(), nnz,
                                           w.data_ptr(), xA.data_ptr(), states.data_ptr(), dy.data_ptr(), lam.data_ptr(),
                                           dw.data_ptr(), dxA.data_ptr(), st))
        return dw, dxA, None, None


def _evgf_sparse(Phi, struct, x, b):
    F_, E, K, G, N, _ = Phi.shape
    B = x.shape[0]
    if x.device.type != "cuda":
        raise RuntimeError("b200gf: EVGF needs CUDA tensors (there is no CPU fallback); got x on %s" % x.device)
    if Phi.dtype != x.dtype or x.dtype not in _ENUM:
        raise RuntimeError("b200gf: EVGF expects Phi and x of one dtype (float32 / float64)")
    y = torch.zeros((B, F_, N), dtype=x.dtype, device=x.device)
    if struct.NA > 0:
        xA = x.index_select(2, struct.A)
        yA = None
        for e in range(E):
            pe = struct.per_e[e]
            if pe["nnz"] == 0:
                continue
            w = Phi[:, e].reshape(F_, K, G, N * N).index_select(3, pe["lin"])   # [F, K, G, nnz], differentiable gather
            ye = _EVChain.apply(w, xA, pe, struct.NA)
            yA = ye if yA is None else yA + ye
        if yA is not None:
            y = y.index_add(2, struct.A, yA)
    if b is not None:
        y = y + b
    return y


def EVGF(S, x, b=None):
    """EVGF(filter_matrices, input, bias=None): same contract as alegnn/utils/graphML.py:389-488.
    S [F, E, K, G, N, N] dense filter matrices (zeros off the pattern), x [B, G, N], b [F, 1] / [F, N] / None."""
    F_, E, K, G, N = S.shape[0], S.shape[1], S.shape[2], S.shape[3], S.shape[4]
    assert S.shape[5] == N                       # graphML.py:441
    assert x.shape[1] == G                       # graphML.py:443
    assert x.shape[2] == N                       # graphML.py:444
    pattern = (S.detach() != 0).any(dim=0).any(dim=1).any(dim=1)                # [E, N, N]
    return _evgf_sparse(S, EVStructure(pattern), x, b)


class EdgeVariantGF(nn.Module):
    """EdgeVariantGF(in_features, out_features, shift_taps, selected_nodes, number_nodes, edge_features=1, bias=True)
    — same surface as the reference layer (graphML.py:2511-2712)."""

    def __init__(self, G, F, K, M, N, E=1, bias=True):
        super().__init__()
        self.G, self.F, self.K, self.E, self.M, self.N = G, F, K, E, M, N
        self.S = None
        self.weightEV = nn.parameter.Parameter(torch.Tensor(F, E, K, G, N, N))
        if self.M < self.N:
            self.weightLSI = nn.parameter.Parameter(torch.Tensor(F, E, K, G))
        else:
            self.register_parameter("weightLSI", None)
        if bias:
            self.bias = nn.parameter.Parameter(torch.Tensor(F, 1))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self):
        stdv = 1. / math.sqrt(self.G * self.K * self.N)      # graphML.py:2601
        self.weightEV.data.uniform_(-stdv, stdv)
        if self.weightLSI is not None:
            self.weightLSI.data.uniform_(-stdv, stdv)
        if self.bias is not None:
            self.bias.data.uniform_(-stdv, stdv)

    def addGSO(self, S):
        assert len(S.shape) == 3
        assert S.shape[0] == self.E
        self.N = S.shape[1]
        assert S.shape[2] == self.N
        self.S = S
        N, M, E, K = self.N, self.M, self.E, self.K
        eye = torch.eye(N, device=S.device, dtype=S.dtype).reshape(1, N, N).repeat(E, 1, 1)
        pattern = (torch.abs(S) + eye) > zeroTolerance                          # graphML.py:2620
        idx = torch.arange(N, device=S.device)
        hybrid = (idx[:, None] < M) | (idx[None, :] < M) if M < N else torch.ones(N, N, dtype=torch.bool, device=S.device)
        pattern = pattern & hybrid[None]                                        # graphML.py:2624-2644
        self.sparsityPattern = pattern.to(S.dtype)
        ident = (eye > 0) & hybrid[None]
        if K > 1:                                                               # graphML.py:2653-2663: k = 0 is the identity
            full = torch.cat([ident.reshape(1, E, 1, 1, N, N),
                              pattern.reshape(1, E, 1, 1, N, N).repeat(1, 1, K - 1, 1, 1, 1)], dim=2)
        else:
            full = ident.reshape(1, E, 1, 1, N, N)
        self.sparsityPatternFull = full.to(S.dtype)
        self._struct = EVStructure(pattern | ident)

    def forward(self, x):
        B, Fin, Nin = x.shape
        self.Phi = self.weightEV * self.sparsityPatternFull                     # graphML.py:2676
        if Nin < self.N:
            x = torch.cat((x, torch.zeros(B, Fin, self.N - Nin, dtype=x.dtype, device=x.device)), dim=2)
        uEV = _evgf_sparse(self.Phi, self._struct, x, self.bias)
        if self.M < self.N:
            uLSI = LSIGF(self.weightLSI, self.S, x, self.bias)                  # bias again, as the reference does (:2686)
        else:
            uLSI = torch.tensor(0., dtype=uEV.dtype, device=uEV.device)
        u = uEV + uLSI
        if Nin < self.N:
            u = u[:, :, :Nin]
        return u

    def extra_repr(self):
        s = "in_features=%d, out_features=%d, " % (self.G, self.F) + "shift_taps=%d, " % (self.K) + \
            "selected_nodes=%d, " % (self.M) + "number_nodes=%d, " % (self.N) + "edge_features=%d, " % (self.E) + \
            "bias=%s, " % (self.bias is not None)
        return s + ("GSO stored" if self.S is not None else "no GSO stored")
