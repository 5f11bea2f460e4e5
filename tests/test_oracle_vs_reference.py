"""Live check of the CPU oracle against the UNMODIFIED reference imported from /root/reference (skipped on the GPU box,
where the reference tree does not exist).  Also checks that install() retargets the reference module in place."""
import numpy as np
import pytest
import torch

import lsigf_oracle as orc
import ref_import

pytestmark = pytest.mark.skipif(not ref_import.reference_available(), reason="reference tree not present")


def _rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.mark.parametrize("seed", range(6))
def test_lsigf_random_cases_vs_reference(seed):
    gml = ref_import.import_reference()
    rng = np.random.default_rng(seed)
    N = int(rng.integers(5, 40)); B = int(rng.integers(1, 4)); G = int(rng.integers(1, 6)); F = int(rng.integers(1, 6))
    K = int(rng.integers(1, 6)); E = int(rng.integers(1, 4))
    c = orc.random_case(1000 + seed, N, B, G, F, K, E, avg_deg=4, bias=["F1", "FN", None][seed % 3])
    t = lambda a: None if a is None else torch.tensor(a)  # noqa: E731
    h, x = t(c["h"]).requires_grad_(True), t(c["x"]).requires_grad_(True)
    y = gml.LSIGF(h, t(c["S"]), x, t(c["b"]))
    y.backward(t(c["dy"]))
    assert _rel(orc.lsigf_dense(c["h"], c["S"], c["x"], c["b"]), y.detach().numpy()) < 1e-12
    dh, dx, _ = orc.lsigf_grads_dense(c["h"], c["S"], c["x"], c["dy"])
    assert _rel(dh, h.grad.numpy()) < 1e-12 and _rel(dx, x.grad.numpy()) < 1e-12
    ys = orc.lsigf_sparse(c["h"], list(c["S"]), c["x"], c["b"])
    assert _rel(ys, y.detach().numpy()) < 1e-12


def test_install_retargets_reference_module():
    import gnn_b200
    gml = ref_import.import_reference()
    orig = (gml.LSIGF, gml.GraphFilter, gml.EVGF, gml.EdgeVariantGF, gml.MaxPoolLocal)
    try:
        gnn_b200.install(gml)
        assert gml.LSIGF is gnn_b200.LSIGF and gml.GraphFilter is gnn_b200.GraphFilter
        assert gml.EVGF is gnn_b200.EVGF and gml.EdgeVariantGF is gnn_b200.EdgeVariantGF
        assert gml.MaxPoolLocal is gnn_b200.MaxPoolLocal
        # an architecture built now gets the B200 layer, with the reference's parameter names
        import torch.nn as nn
        import alegnn.modules.architectures as archit
        S = np.eye(8) * 0.5 + np.diag(np.ones(7), 1) * 0.25
        net = archit.SelectionGNN([1, 4], [3], True, nn.ReLU, [8], gml.NoPool, [1], [2], S)
        assert isinstance(net.GFL[0], gnn_b200.GraphFilter)
        assert sorted(net.state_dict().keys()) == ["GFL.0.bias", "GFL.0.weight", "MLP.0.bias", "MLP.0.weight"]
        assert tuple(net.GFL[0].weight.shape) == (4, 1, 3, 1)
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            net(torch.zeros(2, 1, 8, dtype=net.GFL[0].weight.dtype))
    finally:
        gnn_b200.uninstall(gml)
    assert (gml.LSIGF, gml.GraphFilter, gml.EVGF, gml.EdgeVariantGF, gml.MaxPoolLocal) == orig


@pytest.mark.parametrize("kind", ["HiddenState", "TimeGatedHiddenState", "NodeGatedHiddenState"])
@pytest.mark.parametrize("seed", range(3))
def test_recurrent_layers_vs_reference_live(kind, seed, monkeypatch):
    """gnn_b200.recurrent against the reference layers (graphML.py:3540-4031) on random shapes.  The dense CPU oracle
    stands in for the CUDA filter, so this checks the recursion / gating / autograd wiring, not the kernels.  Seeded
    construction + addGSO consume the RNG in the reference's order, so both layers hold identical parameters."""
    import gnn_b200
    from gnn_b200 import recurrent as rec
    gml = ref_import.import_reference()
    lsigf = lambda h, S, x, b=None: orc.lsigf_dense_torch(h, S, x, b)  # noqa: E731
    monkeypatch.setattr(rec, "_lsigf", lsigf)
    monkeypatch.setattr(gnn_b200.graphML, "LSIGF", lsigf)
    rng = np.random.default_rng(7000 + seed)
    N, B, T = int(rng.integers(4, 20)), int(rng.integers(1, 4)), int(rng.integers(1, 6))
    F, H, K = int(rng.integers(1, 4)), int(rng.integers(1, 5)), int(rng.integers(1, 4))
    E = int(rng.integers(1, 3)) if kind == "HiddenState" else 1     # the gate GRNNs are built with E = 1 (:3757)
    bias = bool(seed % 2 == 0)
    S = torch.tensor(orc.random_sparse_gso(rng, N, 4, E))
    x, z0 = rng.standard_normal((B, T, F, N)), rng.standard_normal((B, H, N))
    dz = rng.standard_normal((B, T, H, N))
    outs = []
    for mod in (gml, rec):
        torch.manual_seed(seed)
        torch.set_default_dtype(torch.float64)
        try:
            layer = getattr(mod, kind)(F, H, K, E=E, bias=bias)
            layer.addGSO(S)
        finally:
            torch.set_default_dtype(torch.float32)
        xt, zt = torch.tensor(x, requires_grad=True), torch.tensor(z0, requires_grad=True)
        z, zT = layer(xt, zt)
        z.backward(torch.tensor(dz))
        outs.append((z.detach().numpy(), zT.detach().numpy(), xt.grad.numpy(), zt.grad.numpy(),
                     {n: (p.detach().numpy(), p.grad.numpy()) for n, p in layer.named_parameters()}))
    (z_r, zT_r, dx_r, dz0_r, p_r), (z_m, zT_m, dx_m, dz0_m, p_m) = outs
    assert sorted(p_r) == sorted(p_m)
    for n in p_r:
        assert np.array_equal(p_r[n][0], p_m[n][0]), n             # same init
        assert _rel(p_m[n][1], p_r[n][1]) < 1e-11, n
    assert z_m.shape == z_r.shape and zT_m.shape == zT_r.shape
    assert _rel(z_m, z_r) < 1e-12 and _rel(zT_m, zT_r) < 1e-12
    assert _rel(dx_m, dx_r) < 1e-11 and _rel(dz0_m, dz0_r) < 1e-11


def test_install_retargets_recurrent_layers(monkeypatch):
    import gnn_b200
    from gnn_b200 import recurrent as rec
    gml = ref_import.import_reference()
    import alegnn.modules.architectures as archit
    orig = (gml.HiddenState, gml.TimeGatedHiddenState, gml.NodeGatedHiddenState, gml.GatedGRNN, gml.EdgeGatedHiddenState)
    S = np.eye(6) * 0.5 + np.diag(np.ones(5), 1) * 0.25
    x = torch.tensor(np.random.default_rng(0).standard_normal((3, 4, 2, 6)), dtype=torch.float32)

    def build_and_run():
        torch.manual_seed(11)
        net = archit.GraphRecurrentNN(2, 3, 3, [2, 2], True, torch.tanh, torch.relu, torch.nn.ReLU, [2], S.astype(np.float32))
        torch.manual_seed(12)                       # z0 is drawn inside splitForward (architectures.py:4547)
        return net, net(x)

    _, y_ref = build_and_run()
    try:
        gnn_b200.install(gml)
        # the whole architecture through the retargeted layers, with the CPU oracle standing in for the CUDA filter
        lsigf = lambda h, S_, x_, b=None: orc.lsigf_dense_torch(h, S_, x_, b)  # noqa: E731
        monkeypatch.setattr(rec, "_lsigf", lsigf)
        monkeypatch.setattr(gnn_b200.graphML, "LSIGF", lsigf)
        net, y = build_and_run()
        assert isinstance(net.hiddenState, gnn_b200.HiddenState) and isinstance(net.outputState, gnn_b200.GraphFilter)
        assert {"hiddenState.aWeights", "hiddenState.bWeights", "hiddenState.xBias", "hiddenState.zBias",
                "outputState.weight", "outputState.bias"} <= set(net.state_dict())
        assert gml.GatedGRNN is orig[3] and gml.EdgeGatedHiddenState is orig[4]   # edge gating stays with the reference
        assert y.shape == y_ref.shape and torch.allclose(y, y_ref, rtol=1e-5, atol=1e-6)
    finally:
        gnn_b200.uninstall(gml)
    assert (gml.HiddenState, gml.TimeGatedHiddenState, gml.NodeGatedHiddenState) == orig[:3]


def test_install_retargets_batch_delay_filter(monkeypatch):
    """LocalGNN_DB (architecturesTime.py:33-205: two GraphFilter_DB layers + tanh + per-node readout) built from the
    retargeted module equals the unmodified architecture; the dense CPU oracle applied to the space-time CSR stands in
    for the CUDA filter."""
    import gnn_b200
    from gnn_b200 import delayed
    gml = ref_import.import_reference()
    import alegnn.modules.architecturesTime as architTime
    orig = (gml.LSIGF_DB, gml.GraphFilter_DB)
    rng = np.random.default_rng(3)
    B, T, N, E = 3, 5, 7, 2
    S = torch.tensor(np.stack([np.stack([orc.random_sparse_gso(rng, N, 3, E) for _ in range(T)]) for _ in range(B)]),
                     dtype=torch.float32)
    x = torch.tensor(rng.standard_normal((B, T, 2, N)), dtype=torch.float32)

    def build_and_run():
        torch.manual_seed(21)
        net = architTime.LocalGNN_DB([2, 4, 3], [3, 2], True, torch.nn.Tanh, [5, 2], E)
        y = net(x, S)
        y.sum().backward()
        return net, y, [p.grad.clone() for p in net.parameters()]

    _, y_ref, g_ref = build_and_run()

    def apply(h, S_, x_big, b_big):
        csr, M = delayed.block_delay_csr(S_)
        dense = torch.zeros(len(csr), M, M, dtype=S_.dtype)
        for e, (rowptr, col, val) in enumerate(csr):
            dense[e, torch.repeat_interleave(torch.arange(M), rowptr[1:] - rowptr[:-1]), col.long()] = val
        return orc.lsigf_dense_torch(h, dense, x_big, b_big)

    try:
        gnn_b200.install(gml)
        monkeypatch.setattr(delayed, "_apply", apply)
        net, y, g = build_and_run()
        assert isinstance(net.GFL[0], gnn_b200.GraphFilter_DB) and isinstance(net.GFL[2], gnn_b200.GraphFilter_DB)
        assert y.shape == y_ref.shape and torch.allclose(y, y_ref, rtol=1e-5, atol=1e-6)
        for a, b in zip(g, g_ref):
            assert torch.allclose(a, b, rtol=1e-4, atol=1e-5)
    finally:
        gnn_b200.uninstall(gml)
    assert (gml.LSIGF_DB, gml.GraphFilter_DB) == orig


def test_install_retargets_batch_delay_recurrence(monkeypatch):
    """GraphRecurrentNN_DB (architecturesTime.py:273-470: HiddenState_DB -> GraphFilter_DB -> tanh -> per-node readout) built
    from the retargeted module equals the unmodified architecture, output and every parameter gradient; torch.sparse on the
    per-time-step CSR operators stands in for the hop kernel, the dense CPU oracle for the filters."""
    import gnn_b200
    from gnn_b200 import delayed
    gml = ref_import.import_reference()
    import alegnn.modules.architecturesTime as architTime
    orig = (gml.GRNN_DB, gml.HiddenState_DB)
    rng = np.random.default_rng(4)
    B, T, N, E = 2, 6, 7, 2
    S = torch.tensor(np.stack([np.stack([orc.random_sparse_gso(rng, N, 3, E) for _ in range(T)]) for _ in range(B)]),
                     dtype=torch.float32)
    x = torch.tensor(rng.standard_normal((B, T, 2, N)), dtype=torch.float32)

    def build_and_run():
        torch.manual_seed(22)                       # parameters, then the random initial hidden state (:447)
        net = architTime.GraphRecurrentNN_DB(2, 3, 4, [3, 2], True, torch.tanh, torch.tanh, torch.nn.Tanh, [5, 2], E)
        y = net(x, S)
        y.sum().backward()
        return net, y, [p.grad.clone() for p in net.parameters()]

    _, y_ref, g_ref = build_and_run()

    def apply(h, S_, x_big, b_big):
        csr, M = delayed.block_delay_csr(S_)
        dense = torch.zeros(len(csr), M, M, dtype=S_.dtype)
        for e, (rowptr, col, val) in enumerate(csr):
            dense[e, torch.repeat_interleave(torch.arange(M), rowptr[1:] - rowptr[:-1]), col.long()] = val
        return orc.lsigf_dense_torch(h, dense, x_big, b_big)

    class SparseSlabOps:
        def __init__(self, S_):
            fwd, _, R = delayed.slab_csr(S_)
            self.A = [torch.sparse_coo_tensor(torch.stack((torch.repeat_interleave(torch.arange(R), rp[1:] - rp[:-1]), col.long())),
                                              val, (R, R)) for (rp, col, val) in fwd]

        def hop(self, o, src):
            return torch.sparse.mm(self.A[o], src)

    try:
        gnn_b200.install(gml)
        monkeypatch.setattr(delayed, "_apply", apply)
        monkeypatch.setattr(delayed, "_slab_ops", SparseSlabOps)
        net, y, g = build_and_run()
        assert isinstance(net.hiddenState, gnn_b200.HiddenState_DB) and isinstance(net.outputState, gnn_b200.GraphFilter_DB)
        assert {"hiddenState.aWeights", "hiddenState.bWeights", "hiddenState.xBias", "hiddenState.zBias",
                "outputState.weight", "outputState.bias"} <= set(net.state_dict())
        assert y.shape == y_ref.shape and torch.allclose(y, y_ref, rtol=1e-5, atol=1e-6)
        for a, b in zip(g, g_ref):
            assert torch.allclose(a, b, rtol=1e-4, atol=1e-5)
    finally:
        gnn_b200.uninstall(gml)
    assert (gml.GRNN_DB, gml.HiddenState_DB) == orig


@pytest.mark.parametrize("dataType", [np.float64, torch.float64])
def test_sparse_source_localization_matches_reference_dataset(dataType):
    """gnn_b200.datasets_sparse.SourceLocalization (sparse mat-vec diffusion) == the reference data class
    (dataTools.py:472-592, dense matrix powers) for the same numpy seed: signals, labels, splits and the helper methods."""
    import scipy.sparse as sp
    from gnn_b200 import datasets_sparse
    ref_import.import_reference()
    import alegnn.utils.graphTools as graphTools
    import alegnn.utils.dataTools as dataTools
    np.random.seed(3)
    G = graphTools.Graph("SBM", 30, {"nCommunities": 3, "probIntra": 0.7, "probInter": 0.2})
    sources = [2, 11, 23]

    class SparseG:                                   # what a large-graph caller would hold: no dense W anywhere
        N, W = G.N, sp.csr_matrix(G.W)

    sets = []
    for cls, graph in ((dataTools.SourceLocalization, G), (datasets_sparse.SourceLocalization, SparseG)):
        np.random.seed(5)
        d = cls(graph, 20, 6, 7, sources, tMax=9, dataType=dataType)
        d.expandDims()
        np.random.seed(6)
        sets.append((d, d.getSamples("train", 5), d.getSamples("test", [1, 3]), d.getSamples("valid", 2)))
    (ref, *ref_draws), (mine, *my_draws) = sets
    num = lambda a: a.numpy() if isinstance(a, torch.Tensor) else np.asarray(a)  # noqa: E731
    for part in ("train", "valid", "test"):
        xs, ys = mine.samples[part]["signals"], mine.samples[part]["targets"]
        xr, yr = ref.samples[part]["signals"], ref.samples[part]["targets"]
        assert type(xs) is type(xr) and xs.dtype == xr.dtype and ys.dtype == yr.dtype and tuple(xs.shape) == tuple(xr.shape)
        assert np.abs(num(xs) - num(xr)).max() < 1e-13 and np.array_equal(num(ys), num(yr))
    for (xa, ya), (xb, yb) in zip(my_draws, ref_draws):
        assert tuple(xa.shape) == tuple(xb.shape) and np.abs(num(xa) - num(xb)).max() < 1e-13 and np.array_equal(num(ya), num(yb))
    scores = np.random.default_rng(0).standard_normal((7, 3))
    yhat = torch.tensor(scores) if dataType is torch.float64 else scores
    assert float(mine.evaluate(yhat, mine.samples["test"]["targets"])) == float(ref.evaluate(yhat, ref.samples["test"]["targets"]))


def test_reference_gatedgrnn_with_biases_after_install(monkeypatch):
    """the reference's own GatedGRNN (graphML.py:1292-1527) reshapes its biases to (1, H, 1) before calling
    the module-global LSIGF (:1394-1404, :1403, :1461).  After install() that call lands in gnn_b200.LSIGF, whose argument
    handling must accept what the reference's broadcast add accepts.  The dense CPU oracle stands in for the CUDA
    dispatch (`graphML._dispatch`), so the shape handling of the product function itself is what runs here."""
    import gnn_b200
    gml = ref_import.import_reference()
    seen = []

    def dispatch(h, S, x, b):
        seen.append(None if b is None else tuple(b.shape))
        return orc.lsigf_dense_torch(h, S, x, b)

    rng = np.random.default_rng(5)
    B, T, F, H, N, K, E = 2, 3, 2, 4, 7, 3, 1
    a = torch.tensor(rng.standard_normal((H, E, K, F)))
    bt = torch.tensor(rng.standard_normal((H, E, K, H)))
    S = torch.tensor(orc.random_sparse_gso(rng, N, 3, E))
    x = torch.tensor(rng.standard_normal((B, T, F, N)))
    z0 = torch.tensor(rng.standard_normal((B, H, N)))
    xb, zb = torch.tensor(rng.standard_normal((H, 1))), torch.tensor(rng.standard_normal((H, 1)))
    want = gml.GatedGRNN(a, bt, S, x, z0, torch.tanh, xBias=xb, zBias=zb)
    try:
        gnn_b200.install(gml)
        monkeypatch.setattr(gnn_b200.graphML, "_dispatch", dispatch)
        got = gml.GatedGRNN(a, bt, S, x, z0, torch.tanh, xBias=xb, zBias=zb)
    finally:
        gnn_b200.uninstall(gml)
    assert seen and all(s == (H, 1) for s in seen)          # (1, H, 1) was normalised to the [F, 1] the C ABI takes
    for w, g in zip(want, got):
        assert _rel(g.detach().numpy(), w.detach().numpy()) < 1e-12
    # shapes the reference's broadcast would reject are still rejected loudly
    with pytest.raises(RuntimeError, match="bias must broadcast"):
        gnn_b200.LSIGF(a, S, x[:, 0], torch.zeros(H + 1, 1, dtype=torch.float64))


def test_fuse_layers_on_a_reference_architecture(monkeypatch):
    """SURVEY.md §8 f-1 at the architecture level: an unmodified reference `SelectionGNN` with two graph-convolutional
    layers and `MaxPoolLocal` (architectures.py:166-296), built once with the reference's own layers and once after
    install() + fuse_layers() with this package's — same seeds, same parameters.  Output, input gradient and every
    parameter gradient must agree; the state_dict keys must not change.  CPU: the oracle / a torch gather stand in for the
    two CUDA dispatch hooks, everything else (fused-activation plumbing, Identity rewiring, neighbourhoods) is product code."""
    import torch.nn as nn
    import gnn_b200
    from gnn_b200 import graphML, pooling
    gml = ref_import.import_reference()
    import alegnn.modules.architectures as archit

    def dispatch(h, S, x, b, act=0):
        y = orc.lsigf_dense_torch(h, S, x, b)
        return torch.relu(y) if act else y

    def gather_max(x, nb32, n_out, max_nb):
        B, F, _ = x.shape
        return x.index_select(2, nb32.reshape(-1).long()).reshape(B, F, n_out, max_nb).max(dim=3)[0]

    rng = np.random.default_rng(21)
    N = 24
    A = (rng.random((N, N)) < 0.2).astype(np.float64)
    A = np.maximum(A, A.T)
    np.fill_diagonal(A, 0)
    S = A / max(np.abs(np.linalg.eigvalsh(A)).max(), 1e-9)
    x = torch.tensor(rng.standard_normal((3, 2, N)))

    def build():
        torch.manual_seed(5)
        torch.set_default_dtype(torch.float64)
        try:
            return archit.SelectionGNN([2, 4, 3], [3, 2], True, nn.ReLU, [10, 6], gml.MaxPoolLocal, [1, 2], [5], S)
        finally:
            torch.set_default_dtype(torch.float32)

    ref_net = build()
    xr = x.clone().requires_grad_(True)
    y_ref = ref_net(xr)
    y_ref.sum().backward()
    try:
        gnn_b200.install(gml)
        monkeypatch.setattr(graphML, "_dispatch", dispatch)
        monkeypatch.setattr(pooling, "_gather_max", gather_max)
        net = build()
        assert isinstance(net.GFL[0], gnn_b200.GraphFilter) and isinstance(net.GFL[2], gnn_b200.MaxPoolLocal)
        assert sorted(net.state_dict().keys()) == sorted(ref_net.state_dict().keys())
        net.load_state_dict(ref_net.state_dict())
        assert gnn_b200.fuse_layers(net) == 2
        assert isinstance(net.GFL[1], nn.Identity) and isinstance(net.GFL[4], nn.Identity)
        assert sorted(net.state_dict().keys()) == sorted(ref_net.state_dict().keys())
        xm = x.clone().requires_grad_(True)
        y = net(xm)
        y.sum().backward()
    finally:
        gnn_b200.uninstall(gml)
    assert _rel(y.detach().numpy(), y_ref.detach().numpy()) < 1e-12
    assert _rel(xm.grad.numpy(), xr.grad.numpy()) < 1e-11
    for (n1, p1), (n2, p2) in zip(sorted(ref_net.named_parameters()), sorted(net.named_parameters())):
        assert n1 == n2 and _rel(p2.grad.numpy(), p1.grad.numpy()) < 1e-11, n1
