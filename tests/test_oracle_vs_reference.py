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
