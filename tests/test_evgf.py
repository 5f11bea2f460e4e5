"""Edge-variant filter: EVGF / EdgeVariantGF.
CPU: the numpy oracle vs fixtures produced by the unmodified reference (oracle/make_golden.py::gen_evgf).
GPU: the sparse CUDA execution (csrc/ev.cu through gnn_b200.EVGF / EdgeVariantGF) vs the same fixtures."""
import os

import numpy as np
import pytest
import torch

import lsigf_oracle as orc


def _rel(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.fixture(scope="module")
def z(golden_dir):
    return np.load(os.path.join(golden_dir, "evgf_cases.npz"))


def test_oracle_evgf_functional(z):
    y = orc.evgf_dense(z["f_Phi"], z["f_x"], z["f_b"])
    assert _rel(y, z["f_y"]) < 1e-12


@pytest.mark.parametrize("tag", ["full", "hyb"])
def test_oracle_edge_variant_layer(z, tag):
    N, M, E, K, G, F, B, Nin = [int(v) for v in z[tag + "_meta"]]
    x = z[tag + "_x"]
    if Nin < N:
        x = np.concatenate([x, np.zeros((B, G, N - Nin))], axis=2)
    wl = z[tag + "_p_weightLSI"] if (tag + "_p_weightLSI") in z.files else None
    y = orc.edge_variant_gf_forward(z[tag + "_p_weightEV"], wl, z[tag + "_p_bias"], z[tag + "_S"], M, x)[:, :, :Nin]
    assert _rel(y, z[tag + "_y"]) < 1e-12


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-11), (torch.float32, 1e-4)])
def test_evgf_functional_gpu(z, dtype, tol):
    import gnn_b200
    Phi = torch.tensor(z["f_Phi"], dtype=dtype, device="cuda", requires_grad=True)
    x = torch.tensor(z["f_x"], dtype=dtype, device="cuda", requires_grad=True)
    b = torch.tensor(z["f_b"], dtype=dtype, device="cuda", requires_grad=True)
    y = gnn_b200.EVGF(Phi, x, b)
    y.backward(torch.tensor(z["f_dy"], dtype=dtype, device="cuda"))
    assert _rel(y.detach().cpu().numpy(), z["f_y"]) < tol
    assert _rel(x.grad.cpu().numpy(), z["f_dx"]) < tol
    assert _rel(b.grad.cpu().numpy(), z["f_db"]) < tol
    # gradient on the sparsity pattern (entries of Phi that are exactly zero are not parameters of the sparse filter)
    assert _rel(Phi.grad.cpu().numpy(), z["f_dPhi"]) < tol


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["full", "hyb"])
@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-11), (torch.float32, 1e-4)])
def test_edge_variant_layer_gpu(z, tag, dtype, tol):
    """Reference parameter names/shapes load as a state_dict; output, dx and every parameter gradient match the
    reference layer (hybrid: LSI part through the B200 LSIGF, bias counted twice as in graphML.py:2682,2686)."""
    import gnn_b200
    N, M, E, K, G, F, B, Nin = [int(v) for v in z[tag + "_meta"]]
    layer = gnn_b200.EdgeVariantGF(G, F, K, M, N, E, True).to("cuda", dtype)
    sd = {k[len(tag) + 3:]: torch.tensor(z[k], dtype=dtype, device="cuda") for k in z.files if k.startswith(tag + "_p_")}
    layer.load_state_dict(sd)
    layer.addGSO(torch.tensor(z[tag + "_S"], dtype=dtype, device="cuda"))
    x = torch.tensor(z[tag + "_x"], dtype=dtype, device="cuda", requires_grad=True)
    y = layer(x)
    assert tuple(y.shape) == z[tag + "_y"].shape
    y.backward(torch.tensor(z[tag + "_dy"], dtype=dtype, device="cuda"))
    assert _rel(y.detach().cpu().numpy(), z[tag + "_y"]) < tol
    assert _rel(x.grad.cpu().numpy(), z[tag + "_dx"]) < tol
    for name, p in layer.named_parameters():
        assert _rel(p.grad.cpu().numpy(), z[tag + "_g_" + name]) < tol, name


@pytest.mark.gpu
def test_hybrid_edge_variant_at_scale():
    """cfg4-shaped hybrid at a size the dense reference cannot hold: the EV part lives on the compact node set; checked
    against the fp64 oracle chain evaluated sparsely on that set."""
    import scipy.sparse as sp
    import gnn_b200
    from gnn_b200 import edgevariant as evm
    rng = np.random.default_rng(5)
    N, M, E, K, G, F, B = 3000, 24, 2, 3, 4, 4, 3
    pats = []
    for e in range(E):
        m = sp.random(N, N, density=8.0 / N, format="csr", random_state=np.random.RandomState(e)) + sp.eye(N)
        pats.append(torch.tensor((m != 0).toarray()))
    pattern = torch.stack(pats).cuda()
    idx = torch.arange(N, device="cuda")
    pattern = pattern & ((idx[:, None] < M) | (idx[None, :] < M))[None]
    st = evm.EVStructure(pattern)
    assert st.NA < N // 2                                   # the compact set is what makes the hybrid layer tractable
    xA = torch.randn(B, G, st.NA, dtype=torch.float64, device="cuda")
    y_ref = torch.zeros(B, F, st.NA, dtype=torch.float64, device="cuda")
    y = torch.zeros_like(y_ref)
    for e in range(E):
        pe = st.per_e[e]
        w = torch.randn(F, K, G, pe["nnz"], dtype=torch.float64, device="cuda") * 0.3
        y = y + evm._EVChain.apply(w, xA, pe, st.NA)
        rows, cols = pe["rowidx"].long(), pe["col"].long()
        for f in range(F):
            for g in range(G):
                u = xA[:, g, :]
                for k in range(K):
                    Phi = torch.zeros(st.NA, st.NA, dtype=torch.float64, device="cuda")
                    Phi[rows, cols] = w[f, k, g]
                    u = u @ Phi.t()
                    y_ref[:, f, :] += u
    assert _rel(y.cpu().numpy(), y_ref.cpu().numpy()) < 1e-11
