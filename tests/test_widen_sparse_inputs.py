"""torch sparse tensors as the GSO argument of LSIGF / GraphFilter.addGSO (SURVEY.md §8b "Extension over reference"):
converted once to host CSR (never densified), cached per tensor, same results as the dense route."""
import numpy as np
import pytest
import torch

import lsigf_oracle as orc


@pytest.mark.gpu
@pytest.mark.filterwarnings("ignore:Sparse")
@pytest.mark.parametrize("layout", ["coo", "csr"])
@pytest.mark.parametrize("where", ["cuda", "cpu"])
def test_torch_sparse_gso_matches_dense_route(layout, where):
    import gnn_b200
    c = orc.random_case(31, N=80, B=3, G=4, F=5, K=4, E=2, avg_deg=5, bias="F1")
    S = c["S"].copy()
    S[1] = np.where(S[0] != 0, S[1] + 0.1 * (S[1] == 0), 0.0)             # same pattern for both e (batched CSR needs it)
    t = lambda a: torch.tensor(a, dtype=torch.float64, device="cuda")  # noqa: E731
    Sd = torch.tensor(S, dtype=torch.float64, device=where)
    Ss = Sd.to_sparse() if layout == "coo" else Sd.to_sparse_csr()
    h, x, b = t(c["h"]).requires_grad_(True), t(c["x"]).requires_grad_(True), t(c["b"])
    y = gnn_b200.LSIGF(h, Ss, x, b)
    y.backward(t(c["dy"]))
    rel = lambda a, r: float(np.abs(a.detach().cpu().numpy() - r).max() / np.abs(r).max())  # noqa: E731
    assert rel(y, orc.lsigf_dense(c["h"], S, c["x"], c["b"])) < 1e-12
    dh, dx, _ = orc.lsigf_grads_dense(c["h"], S, c["x"], c["dy"])
    assert rel(h.grad, dh) < 1e-12 and rel(x.grad, dx) < 1e-12
    # cached: the second call reuses the plan built for this tensor object
    from gnn_b200.gso import plan_for
    assert plan_for(Ss, x.device) is plan_for(Ss, x.device)
    layer = gnn_b200.GraphFilter(4, 5, 4, E=2).cuda().double()
    layer.addGSO(Ss)
    with torch.no_grad():
        layer.weight.copy_(h)
        layer.bias.copy_(b)
    assert rel(layer(x.detach()), orc.lsigf_dense(c["h"], S, c["x"], c["b"])) < 1e-12
