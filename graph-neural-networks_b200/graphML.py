"""Drop-in replacements for the LSIGF path of `alegnn.utils.graphML` (reference alegnn/utils/graphML.py).

    LSIGF(h, S, x, b=None)                     <- graphML.py:83-176
    GraphFilter(G, F, K, E=1, bias=True)       <- graphML.py:2036-2155   (same attributes, parameter names and
                                                  shapes, so reference checkpoints `GFL.<i>.weight/bias` round-trip)
    install() / uninstall()                    retarget `alegnn.utils.graphML.LSIGF` / `.GraphFilter` in place
                                                  (SURVEY.md §8b): no reference file is edited.

Host code is PyTorch (tensors, autograd plumbing, streams); the arithmetic runs in libb200gf.so through the
C ABI of include/b200gf.h.  There is no CPU or eager fallback: CPU tensors raise.

Layout.  The reference keeps activations [B, G, N] with the node axis contiguous and returns y as a permuted view
of a [B, N, F] buffer (graphML.py:170-171).  Here y is a permuted view of a node-major [N, B*F(+pad)] buffer (same
logical shape [B, F, N]); when such a view comes back in as the next layer's x — directly, or through an
element-wise op such as nn.ReLU, which preserves strides — it is consumed in place with no transpose.
"""
import math

import torch
import torch.nn as nn

from . import _cabi
from .gso import Plan, SparseGSO, plan_for

_ENUM = {torch.float32: _cabi.F32, torch.float64: _cabi.F64}
_ELEMS_PER_SECTOR = {torch.float32: 8, torch.float64: 4}


def padded_ld(C, dtype):
    q = _ELEMS_PER_SECTOR[dtype]
    return (C + q - 1) // q * q


def node_major_ld(t):
    """If `t` ([B, C, N] logical) is a view of a node-major [N, ld] buffer with columns (b, c), returns ld, else None."""
    if t.dim() != 3:
        return None
    B, C, N = t.shape
    sb, sc, sn = t.stride()
    if N == 0 or B * C == 0:
        return None
    if C > 1 and sc != 1:
        return None
    if B > 1 and sb != C:
        return None
    if N > 1 and sn < B * C:
        return None
    if N == 1:
        sn = B * C
    return sn


def _stream():
    return torch.cuda.current_stream().cuda_stream


def to_node_major(x):
    """[B, C, N] tensor -> (node-major buffer [N, ld], ld).  No copy when x already is such a view."""
    B, C, N = x.shape
    ld = node_major_ld(x)
    if ld is not None:
        return x, ld
    lib = _cabi.load()
    xc = x.contiguous()
    ld = padded_ld(B * C, x.dtype)
    buf = torch.empty((N, ld), dtype=x.dtype, device=x.device)
    _cabi.check(lib.b200gf_to_node_major(_ENUM[x.dtype], xc.data_ptr(), buf.data_ptr(), ld, N, B * C, _stream()))
    return buf, ld


def to_feature_major(y):
    """[B, C, N] result (typically the permuted node-major view LSIGF returns) -> contiguous reference layout, with one
    coalesced tiled transpose (b200gf_to_feature_major) instead of an element-wise strided copy."""
    ld = node_major_ld(y)
    if ld is None or y.device.type != "cuda" or y.dtype not in _ENUM:
        return y.contiguous()
    B, C, N = y.shape
    out = torch.empty((B, C, N), dtype=y.dtype, device=y.device)
    _cabi.check(_cabi.load().b200gf_to_feature_major(_ENUM[y.dtype], y.data_ptr(), ld, out.data_ptr(), N, B * C, _stream()))
    return out


def _as_bcn_view(buf, B, C, N):
    """node-major [N, ld] buffer -> logical [B, C, N] strided view."""
    return buf[:, :B * C].view(N, B, C).permute(1, 2, 0)


class _LSIGFFunction(torch.autograd.Function):
    """y = LSIGF(h, S, x, b) with S fixed inside `plan` (graphML.py:83-176)."""

    @staticmethod
    def forward(ctx, h, x, b, plan, act=0):
        lib = _cabi.load()
        F_, E, K, G = h.shape
        B, _, N = x.shape
        dt = x.dtype
        hc = h.contiguous()
        ctx.x_node_major = node_major_ld(x) is not None
        xn, x_ld = to_node_major(x)
        bias_per_node = 0
        bc = None
        if b is not None:
            bias_per_node = 0 if b.shape[1] == 1 else 1
            bc = b.contiguous()
        ldf = padded_ld(B * F_, dt)
        ybuf = torch.empty((N, ldf), dtype=dt, device=x.device)
        ws_bytes = lib.b200gf_workspace_bytes(plan.handle, B, G, F_, K, _cabi.NODE_MAJOR, 0)
        ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=x.device)
        rc = lib.b200gf_forward_act(plan.handle, xn.data_ptr(), _cabi.NODE_MAJOR, x_ld, hc.data_ptr(),
                                    None if bc is None else bc.data_ptr(), bias_per_node,
                                    ybuf.data_ptr(), _cabi.NODE_MAJOR, ldf, ws.data_ptr(), ws_bytes,
                                    B, G, F_, K, int(act), _stream())
        _cabi.check(rc)
        ctx.act = int(act)
        ctx.ldf = ldf
        ctx.plan = plan
        ctx.x_ld = x_ld
        ctx.bias_per_node = bias_per_node
        ctx.has_bias = b is not None
        ctx.bias_shape = None if b is None else tuple(b.shape)
        ctx.dims = (B, G, F_, K, E, N)
        if act:
            ctx.save_for_backward(hc, xn, ybuf)      # the fused ReLU's backward needs only the layer output (y > 0)
        else:
            ctx.save_for_backward(hc, xn)
        return _as_bcn_view(ybuf, B, F_, N)

    @staticmethod
    def backward(ctx, dy):
        lib = _cabi.load()
        if ctx.act:
            hc, xn, ybuf = ctx.saved_tensors
        else:
            hc, xn = ctx.saved_tensors
        B, G, F_, K, E, N = ctx.dims
        plan = ctx.plan
        dt = hc.dtype
        dyn, dy_ld = to_node_major(dy)
        if ctx.act:                                   # dy_pre = dy * (y > 0), one pass, node-major
            masked = torch.empty((N, ctx.ldf), dtype=dt, device=dy.device)
            _cabi.check(lib.b200gf_relu_backward(_ENUM[dt], ybuf.data_ptr(), ctx.ldf, dyn.data_ptr(), dy_ld,
                                                 masked.data_ptr(), ctx.ldf, N, B * F_, _stream()))
            dyn, dy_ld = masked, ctx.ldf
        need_dh, need_dx, need_db = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.needs_input_grad[2]
        dh = torch.empty_like(hc)
        ldc = padded_ld(B * G, dt)
        dxbuf = torch.empty((N, ldc), dtype=dt, device=dy.device) if need_dx else None
        db = None
        if ctx.has_bias and need_db:
            db = torch.empty(ctx.bias_shape, dtype=dt, device=dy.device)
        ws_bytes = lib.b200gf_workspace_bytes(plan.handle, B, G, F_, K, _cabi.NODE_MAJOR, 1)
        ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dy.device)
        rc = lib.b200gf_backward(plan.handle, dyn.data_ptr(), _cabi.NODE_MAJOR, dy_ld,
                                 xn.data_ptr(), _cabi.NODE_MAJOR, ctx.x_ld, hc.data_ptr(),
                                 None if dxbuf is None else dxbuf.data_ptr(), _cabi.NODE_MAJOR, ldc,
                                 dh.data_ptr(), None if db is None else db.data_ptr(), ctx.bias_per_node,
                                 ws.data_ptr(), ws_bytes, B, G, F_, K, _stream())
        _cabi.check(rc)
        dx = None
        if need_dx:
            if ctx.x_node_major:
                dx = _as_bcn_view(dxbuf, B, G, N)       # the producer of x reads it through the same strides
            else:
                # x was a plain [B, G, N] tensor: hand back a contiguous gradient (one coalesced tiled transpose)
                # instead of a strided view that autograd would re-copy element-wise (measured 1.1 ms vs 0.1 ms)
                dx = torch.empty((B, G, N), dtype=dt, device=dy.device)
                _cabi.check(lib.b200gf_to_feature_major(_ENUM[dt], dxbuf.data_ptr(), ldc, dx.data_ptr(), N, B * G,
                                                        _stream()))
        return (dh if need_dh else None), dx, db, None, None


def LSIGF(h, S, x, b=None, activation=None):
    """LSIGF(filter_taps, GSO, input, bias=None): linear shift-invariant graph filter, then bias.
    `activation="relu"` (extension) fuses the layer's ReLU into the contraction epilogue.

    Same contract as the reference (alegnn/utils/graphML.py:83-176):
        h [F, E, K, G]; S [E, N, N] (dense tensor, or SparseGSO / Plan); x [B, G, N]; b [F, 1] or [F, N] or None
        returns y [B, F, N],  y_f = sum_e sum_k sum_g h[f,e,k,g] (x_g S_e^k) + b_f
    """
    F_ = h.shape[0]
    E = h.shape[1]
    K = h.shape[2]
    G = h.shape[3]
    assert S.shape[0] == E                       # graphML.py:135
    N = S.shape[1]
    assert S.shape[2] == N                       # graphML.py:137
    B = x.shape[0]
    assert x.shape[1] == G                       # graphML.py:139
    assert x.shape[2] == N                       # graphML.py:140
    if b is not None:
        # the reference adds b by broadcasting (graphML.py:174-175): GraphFilter passes [F, 1], and the reference's own
        # GatedGRNN reshapes its biases to (1, F, 1) before calling LSIGF (graphML.py:1394-1404, :1461)
        if b.dim() == 3 and b.shape[0] == 1:
            b = b[0]
        elif b.dim() == 1 and N == 1:
            b = b.reshape(-1, 1)
        if not (b.dim() == 2 and b.shape[0] in (1, F_) and b.shape[1] in (1, N)):
            raise RuntimeError("b200gf: LSIGF bias must broadcast against [B, F, N] as [F, 1], [F, N], [1, F, 1] or "
                               "[1, F, N]; got %s" % (tuple(b.shape),))
        if b.shape[0] == 1 and F_ > 1:
            b = b.expand(F_, b.shape[1])
    if activation is None:
        return _dispatch(h, S, x, b)
    if activation != "relu":
        raise ValueError("b200gf: fused activation must be None or 'relu', got %r" % (activation,))
    return _dispatch(h, S, x, b, 1)


def _dispatch_cuda(h, S, x, b, act=0):
    """Device part of LSIGF: loud checks (there is no CPU path), plan lookup, the autograd function over the C ABI.
    `_dispatch` is the single hook the CPU tests replace with the oracle to exercise the argument handling above."""
    if x.device.type != "cuda":
        raise RuntimeError("b200gf: LSIGF needs CUDA tensors (there is no CPU fallback); got x on %s" % x.device)
    if x.dtype not in _ENUM:
        raise RuntimeError("b200gf: LSIGF supports float32 and float64, got %s" % x.dtype)
    if h.dtype != x.dtype or S.dtype != x.dtype or (b is not None and b.dtype != x.dtype):
        # torch.matmul in the reference raises on mixed dtypes too ("expected scalar type ...")
        raise RuntimeError("b200gf: LSIGF expects h, S, x, b of one dtype, got h=%s S=%s x=%s" % (h.dtype, S.dtype, x.dtype))
    plan = plan_for(S, x.device)
    if plan.device != x.device and not (plan.device.index == (x.device.index or 0)):
        raise RuntimeError("b200gf: GSO plan lives on %s but x is on %s" % (plan.device, x.device))
    return _LSIGFFunction.apply(h, x, b, plan, act)


_dispatch = _dispatch_cuda


class GraphFilter(nn.Module):
    """GraphFilter(in_features, out_features, filter_taps, edge_features=1, bias=True)

    Same surface as the reference layer (alegnn/utils/graphML.py:2036-2155): attributes G, F, K, E, S, N;
    parameters `weight` [F, E, K, G] and `bias` [F, 1] (or None); `addGSO(S)`, `forward(x)`, `extra_repr()`.
    `addGSO` additionally accepts a SparseGSO and builds the device plan once (cached per device).
    """

    def __init__(self, G, F, K, E=1, bias=True):
        super().__init__()
        self.G = G
        self.F = F
        self.K = K
        self.E = E
        self.S = None
        self.fused_activation = None                 # "relu" after fuse_layers(): the next module became nn.Identity
        self.weight = nn.parameter.Parameter(torch.Tensor(F, E, K, G))
        if bias:
            self.bias = nn.parameter.Parameter(torch.Tensor(F, 1))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self):
        stdv = 1. / math.sqrt(self.G * self.K)      # graphML.py:2109-2114
        self.weight.data.uniform_(-stdv, stdv)
        if self.bias is not None:
            self.bias.data.uniform_(-stdv, stdv)

    def addGSO(self, S):
        assert len(S.shape) == 3                    # graphML.py:2118
        assert S.shape[0] == self.E                 # graphML.py:2120
        self.N = S.shape[1]
        assert S.shape[2] == self.N                 # graphML.py:2122
        self.S = S
        # Build (or fetch from the cache) the device CSR plan now.  A dense GSO still on the CPU gets its plan
        # when the architecture's .to(device) re-attaches it (architectures.py:463-479) or on first use.
        if torch.cuda.is_available() and (isinstance(S, (SparseGSO, Plan)) or
                                          (isinstance(S, torch.Tensor) and S.device.type == "cuda")):
            plan_for(S)

    def forward(self, x):
        B = x.shape[0]
        F = x.shape[1]
        Nin = x.shape[2]
        if Nin < self.N:                            # zero-pad the node axis, graphML.py:2131-2135
            x = torch.cat((x, torch.zeros(B, F, self.N - Nin, dtype=x.dtype, device=x.device)), dim=2)
        if self.fused_activation is None:           # the reference's call, argument for argument (graphML.py:2137)
            u = LSIGF(self.weight, self.S, x, self.bias)  # plan lookup is cached per (tensor, version, device)
        else:
            u = LSIGF(self.weight, self.S, x, self.bias, activation=self.fused_activation)
        if Nin < self.N:                            # keep the first Nin nodes, graphML.py:2142-2143
            u = u[:, :, :Nin]
        return u

    def extra_repr(self):
        reprString = "in_features=%d, out_features=%d, " % (self.G, self.F) + "filter_taps=%d, " % (self.K) + \
                     "edge_features=%d, " % (self.E) + "bias=%s, " % (self.bias is not None)
        if self.S is not None:
            reprString += "GSO stored"
        else:
            reprString += "no GSO stored"
        return reprString


def fuse_layers(model):
    """Fuse `GraphFilter -> nn.ReLU -> (NoPool | MaxPoolLocal)` triples inside every nn.Sequential of `model` (the
    reference builds its graph-filtering layers exactly like that, alegnn/modules/architectures.py:274-296):
    the ReLU moves into the filter's contraction epilogue (the nn.ReLU module is replaced by nn.Identity, so the
    Sequential keeps its indices and the checkpoint keys `GFL.<3l>.weight/bias` are unchanged) and a reference
    MaxPoolLocal is swapped for this package's CUDA gather on the node-major output.  NoPool is an identity already.
    Returns the number of fused layers."""
    from . import pooling
    fused = 0
    for seq in [m for m in model.modules() if isinstance(m, nn.Sequential)]:
        mods = list(seq._modules.items())
        for i, (name, m) in enumerate(mods):
            if isinstance(m, GraphFilter) and i + 1 < len(mods) and type(mods[i + 1][1]) is nn.ReLU:
                m.fused_activation = "relu"
                seq._modules[mods[i + 1][0]] = nn.Identity()
                fused += 1
                if i + 2 < len(mods):
                    pname, pm = mods[i + 2]
                    if type(pm).__name__ == "MaxPoolLocal" and not isinstance(pm, pooling.MaxPoolLocal):
                        seq._modules[pname] = pooling.MaxPoolLocal.from_reference(pm)
    return fused


# ---------------------------------------------------------------------------------------------------
# retargeting the reference
# ---------------------------------------------------------------------------------------------------
_SAVED = {}


def install(gml=None):
    """Point `alegnn.utils.graphML.LSIGF`, `.GraphFilter`, `.EVGF`, `.EdgeVariantGF`, the local pooling / activation
    layers and the static-GSO recurrent layers at this package.

    `GraphFilter.forward` in the reference looks `LSIGF` up as a module global at call time (graphML.py:2137), so
    this also accelerates its hybrid EdgeVariantGF (:2686), jARMA (:592) and GatedGRNN (:1403,:1461) call sites.
    Architectures built AFTER install() get this package's layers (plan cached in addGSO).
    """
    from . import activations, delayed, edgevariant, pooling, recurrent
    if gml is None:
        import alegnn.utils.graphML as gml
    if id(gml) not in _SAVED:
        _SAVED[id(gml)] = (gml, {n: getattr(gml, n) for n in ("LSIGF", "GraphFilter", "EVGF", "EdgeVariantGF",
                                                             "MaxPoolLocal", "MaxLocalActivation",
                                                             "MedianLocalActivation", "HiddenState",
                                                             "TimeGatedHiddenState", "NodeGatedHiddenState",
                                                             "LSIGF_DB", "GraphFilter_DB", "GRNN_DB",
                                                             "HiddenState_DB")})
    gml.LSIGF = LSIGF
    gml.GraphFilter = GraphFilter
    gml.EVGF = edgevariant.EVGF
    gml.EdgeVariantGF = edgevariant.EdgeVariantGF
    gml.MaxPoolLocal = pooling.MaxPoolLocal      # same layer, neighbourhoods from the CSR routine (scales past dense N x N)
    gml.MaxLocalActivation = activations.MaxLocalActivation
    gml.MedianLocalActivation = activations.MedianLocalActivation
    # static-GSO recurrent layers: node-major recursion (recurrent.py).  gml.GatedGRNN itself is left alone so that the
    # reference's EdgeGatedHiddenState (dense per-sample gated GSOs, not on this path) keeps working.
    gml.HiddenState = recurrent.HiddenState
    gml.TimeGatedHiddenState = recurrent.TimeGatedHiddenState
    gml.NodeGatedHiddenState = recurrent.NodeGatedHiddenState
    # batch-/time-varying GSOs: one space-time sparse operator per batch (delayed.py)
    gml.LSIGF_DB = delayed.LSIGF_DB
    gml.GraphFilter_DB = delayed.GraphFilter_DB
    # the recursion over the same GSO batch: the delay line advanced by one CSR hop per time step (delayed.GRNN_DB)
    gml.GRNN_DB = delayed.GRNN_DB
    gml.HiddenState_DB = delayed.HiddenState_DB
    return gml


def uninstall(gml=None):
    for key, (mod, saved) in list(_SAVED.items()):
        if gml is None or mod is gml:
            for name, obj in saved.items():
                setattr(mod, name, obj)
            del _SAVED[key]
