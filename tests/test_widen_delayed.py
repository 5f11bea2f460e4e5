"""Batch- and time-varying graph filter (gnn_b200.delayed: LSIGF_DB, GraphFilter_DB) against fixtures produced by the
unmodified reference (tests/golden/lsigf_db_cases.npz <- oracle/make_golden.py gen_lsigf_db; graphML.py:977-1094,
:3278-3393).

The B200 path folds the B*T per-sample graphs into one space-time sparse operator and runs the ordinary LSIGF kernels
on it.  CPU tests check that construction (CSR layout, delay / zero-history semantics, bias tiling, autograd wiring)
with the dense CPU oracle applied to the same CSR; GPU tests run it through the CUDA filter."""
import os

import numpy as np
import pytest
import torch

import lsigf_oracle as orc

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "lsigf_db_cases.npz"))
TAGS = ["fa", "fb", "fc", "fd", "fe"]


def _rel(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def _dense_from_csr(csr, M, dtype):
    S = torch.zeros(len(csr), M, M, dtype=dtype)
    for e, (rowptr, col, val) in enumerate(csr):
        rows = torch.repeat_interleave(torch.arange(M), rowptr[1:] - rowptr[:-1])
        S[e, rows, col.long()] = val
    return S


@pytest.fixture
def oracle_filter(monkeypatch):
    from gnn_b200 import delayed
    seen = {}

    def apply(h, S, x_big, b_big):
        import gnn_b200
        csr, M = delayed.block_delay_csr(S)
        seen["ld"] = gnn_b200.node_major_ld(x_big)
        return orc.lsigf_dense_torch(h, _dense_from_csr(csr, M, S.dtype), x_big, b_big)

    monkeypatch.setattr(delayed, "_apply", apply)
    return seen


def _case(tag, dtype, device, grad=True):
    t = lambda a: torch.tensor(GOLD[tag + "_" + a], dtype=dtype, device=device)  # noqa: E731
    h, x = t("h").requires_grad_(grad), t("x").requires_grad_(grad)
    b = t("b").requires_grad_(grad) if (tag + "_b") in GOLD.files else None
    return h, t("S"), x, b, t("dy")


def _check_case(tag, dtype, device, tol):
    from gnn_b200 import delayed
    h, S, x, b, dy = _case(tag, dtype, device)
    y = delayed.LSIGF_DB(h, S, x, b)
    assert tuple(y.shape) == GOLD[tag + "_y"].shape
    y.backward(dy)
    assert _rel(y.detach().cpu().numpy(), GOLD[tag + "_y"]) < tol
    assert _rel(h.grad.cpu().numpy(), GOLD[tag + "_dh"]) < tol
    assert _rel(x.grad.cpu().numpy(), GOLD[tag + "_dx"]) < tol
    if b is not None:
        assert _rel(b.grad.cpu().numpy(), GOLD[tag + "_db"]) < tol


def _check_layer(dtype, device, tol):
    from gnn_b200 import delayed
    B, T, N, G, F, K, E = (int(v) for v in GOLD["layer_meta"])
    layer = delayed.GraphFilter_DB(G, F, K, E, True)
    assert sorted(layer.state_dict()) == ["bias", "weight"]
    layer.load_state_dict({"weight": torch.tensor(GOLD["layer_weight"]), "bias": torch.tensor(GOLD["layer_bias"])})
    layer = layer.to(device=device, dtype=dtype)
    layer.addGSO(torch.tensor(GOLD["layer_S"], dtype=dtype, device=device))
    x = torch.tensor(GOLD["layer_x"], dtype=dtype, device=device, requires_grad=True)
    y = layer(x)
    y.backward(torch.tensor(GOLD["layer_dy"], dtype=dtype, device=device))
    assert _rel(y.detach().cpu().numpy(), GOLD["layer_y"]) < tol
    assert _rel(x.grad.cpu().numpy(), GOLD["layer_dx"]) < tol
    assert _rel(layer.weight.grad.cpu().numpy(), GOLD["layer_dweight"]) < tol
    assert _rel(layer.bias.grad.cpu().numpy(), GOLD["layer_dbias"]) < tol
    return layer, x


# ------------------------------------------------------------------------------------------------ CPU host logic
def test_space_time_csr_layout():
    from gnn_b200 import delayed
    rng = np.random.default_rng(1)
    B, T, E, N = 2, 4, 2, 5
    S = torch.tensor(rng.standard_normal((B, T, E, N, N)) * (rng.random((B, T, E, N, N)) < 0.4))
    csr, M = delayed.block_delay_csr(S)
    assert M == B * T * N and len(csr) == E
    big = _dense_from_csr(csr, M, S.dtype)
    for e in range(E):
        rowptr, col, val = csr[e]
        assert rowptr.dtype == torch.int64 and col.dtype == torch.int32 and rowptr[0] == 0 and rowptr[-1] == col.numel()
        for r in range(M):                                    # columns ascending inside every row
            c = col[rowptr[r]:rowptr[r + 1]]
            assert torch.all(c[1:] > c[:-1])
        ref = torch.zeros(M, M, dtype=S.dtype)
        for b in range(B):
            for t in range(1, T):                              # S(b, t) links copy (b, t-1) -> (b, t); S(b, 0) is unused
                ref[(b * T + t - 1) * N:(b * T + t) * N, (b * T + t) * N:(b * T + t + 1) * N] = S[b, t, e]
        assert torch.equal(big[e], ref)
    one, M1 = delayed.block_delay_csr(S[:, :1])               # a single time step: no edges at all
    assert M1 == B * N and all(c.numel() == 0 and r[-1] == 0 for (r, c, v) in one)


@pytest.mark.parametrize("tag", TAGS)
def test_host_logic_matches_reference_fixtures(tag, oracle_filter):
    _check_case(tag, torch.float64, "cpu", 1e-12)
    assert oracle_filter["ld"] == int(GOLD[tag + "_meta"][3])   # the filter sees a node-major [M, G] view: no transpose


def test_layer_host_logic_and_output_view(oracle_filter):
    layer, x = _check_layer(torch.float64, "cpu", 1e-12)
    with pytest.raises(AssertionError):
        layer.addGSO(torch.zeros(2, 3, 3))                     # the reference's 5-D check (graphML.py:3362)
    with pytest.raises(AssertionError):
        layer(x[:, :2])                                        # T must match the stored GSO (graphML.py:3376)


def test_product_path_is_loud_on_cpu():
    from gnn_b200 import delayed
    h, S, x, b, _ = _case("fa", torch.float64, "cpu", grad=False)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        delayed.LSIGF_DB(h, S, x, b)


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("tag", TAGS)
@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-11), (torch.float32, 1e-4)])
def test_gpu_lsigf_db_matches_reference_fixtures(tag, dtype, tol):
    _check_case(tag, dtype, "cuda", tol)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-11), (torch.float32, 1e-4)])
def test_gpu_graphfilter_db_layer(dtype, tol):
    _check_layer(dtype, "cuda", tol)


@pytest.mark.gpu
def test_gpu_flocking_sized_batch_against_batched_matmul():
    """B = 20 trajectories x T = 60 steps x N = 50 agents (the reference's flocking setup, 60 000 space-time nodes):
    compared with the reference's own formulation — per-(b, t) dense products with a unit delay — done with
    torch.matmul in fp64 on the GPU."""
    from gnn_b200 import delayed
    torch.manual_seed(0)
    B, T, N, G, F, K, E = 20, 60, 50, 6, 32, 3, 1
    S = (torch.rand(B, T, E, N, N, device="cuda") < 0.15).double()
    S = S / S.sum(-1, keepdim=True).clamp(min=1.0)
    x = torch.randn(B, T, G, N, device="cuda", dtype=torch.float64)
    h = torch.randn(F, E, K, G, device="cuda", dtype=torch.float64) / (G * K) ** 0.5
    b = torch.randn(F, 1, device="cuda", dtype=torch.float64)
    y = delayed.LSIGF_DB(h, S, x, b)
    z, ref = x.unsqueeze(2).expand(B, T, E, G, N), 0.0
    for k in range(K):
        if k > 0:                                              # z_k(t) = z_{k-1}(t-1) S(t), zero history
            z = torch.matmul(torch.cat((torch.zeros_like(z[:, :1]), z[:, :-1]), 1), S)
        ref = ref + torch.einsum("feg,btegn->btfn", h[:, :, k], z)
    ref = ref + b
    assert float((y - ref).abs().max() / ref.abs().max()) < 1e-12
    y32 = delayed.LSIGF_DB(h.float(), S.float(), x.float(), b.float())
    assert float((y32.double() - ref).abs().max() / ref.abs().max()) < 1e-4


@pytest.mark.gpu
def test_gpu_plan_build_stays_on_the_device():
    """the space-time plan used to travel through the host once per GSO batch (D2H of the CSR, host
    transpose, H2D).  Now both operators are built with torch kernels on the device and adopted device to device
    (b200gf_plan_create_device); same results as the host builder, timing of both printed."""
    import time
    from gnn_b200 import delayed
    from gnn_b200.gso import Plan
    import gnn_b200
    torch.manual_seed(1)
    B, T, N, E = 20, 60, 50, 1
    S = (torch.rand(B, T, E, N, N, device="cuda") < 0.15).float()
    S = S / S.sum(-1, keepdim=True).clamp(min=1.0)
    csr, M = delayed.block_delay_csr(S)

    def clock(fn, reps=3):
        fn(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps * 1e3

    dev_ms = clock(lambda: Plan.from_device_ops([delayed.transpose_csr_device(c, M) for c in csr], csr, M, S.dtype, S.device))
    host_ms = clock(lambda: Plan.from_host_csr(csr, M, S.dtype, S.device))
    p_dev = Plan.from_device_ops([delayed.transpose_csr_device(c, M) for c in csr], csr, M, S.dtype, S.device)
    p_host = Plan.from_host_csr(csr, M, S.dtype, S.device)
    x = torch.randn(1, 4, M, device="cuda")
    h = torch.randn(3, E, 3, 4, device="cuda")
    assert torch.equal(gnn_b200.LSIGF(h, p_dev, x, None), gnn_b200.LSIGF(h, p_host, x, None))
    print("LSIGF_DB plan for B=%d T=%d N=%d (%d space-time nodes, %d nnz): device build %.2f ms, host build %.2f ms" %
          (B, T, N, M, p_dev.nnz, dev_ms, host_ms))


def test_transpose_csr_device_matches_scipy():
    import scipy.sparse as sp
    from gnn_b200 import delayed
    for M, seed in ((1, 0), (7, 1), (60, 2)):
        m = sp.random(M, M, density=0.2, format="csr", random_state=seed)
        m.sort_indices()
        tr, tc, tv = delayed.transpose_csr_device((torch.tensor(m.indptr, dtype=torch.int64),
                                                   torch.tensor(m.indices, dtype=torch.int32), torch.tensor(m.data)), M)
        mt = m.T.tocsr()
        mt.sort_indices()
        assert np.array_equal(tr.numpy(), mt.indptr) and np.array_equal(tc.numpy(), mt.indices)
        assert np.array_equal(tv.numpy(), mt.data)


# ------------------------------------------------------------------------------------------------ CPU, random shapes
from hypothesis import given, settings, strategies as st  # noqa: E402


@settings(max_examples=25, deadline=None, derandomize=True)
@given(B=st.integers(1, 3), T=st.integers(1, 5), N=st.integers(1, 6), G=st.integers(1, 3), F=st.integers(1, 3),
       K=st.integers(1, 5), E=st.integers(1, 2), bias=st.sampled_from([None, "F1", "FN"]), seed=st.integers(0, 10 ** 6))
def test_space_time_filter_equals_delayed_products(B, T, N, G, F, K, E, bias, seed):
    """LSIGF_DB through the space-time operator == the definition (graphML.py:990-997): per-(b, t) dense products with
    one unit delay per tap and zero history, for random shapes incl. T = 1, N = 1, K > T."""
    from gnn_b200 import delayed
    rng = np.random.default_rng(seed)
    S = torch.tensor(rng.standard_normal((B, T, E, N, N)) * (rng.random((B, T, E, N, N)) < 0.6))
    x = torch.tensor(rng.standard_normal((B, T, G, N)))
    h = torch.tensor(rng.standard_normal((F, E, K, G)))
    b = None if bias is None else torch.tensor(rng.standard_normal((F, 1 if bias == "F1" else N)))

    def apply(h_, S_, x_big, b_big):
        csr, M = delayed.block_delay_csr(S_)
        return orc.lsigf_dense_torch(h_, _dense_from_csr(csr, M, S_.dtype), x_big, b_big)

    old = delayed._apply
    delayed._apply = apply
    try:
        y = delayed.LSIGF_DB(h, S, x, b)
    finally:
        delayed._apply = old
    z, ref = x.unsqueeze(2).expand(B, T, E, G, N), 0.0
    for k in range(K):
        if k > 0:
            z = torch.matmul(torch.cat((torch.zeros_like(z[:, :1]), z[:, :-1]), 1), S)
        ref = ref + torch.einsum("feg,btegn->btfn", h[:, :, k], z)
    if b is not None:
        ref = ref + b
    assert y.shape == (B, T, F, N) and torch.allclose(y, ref, rtol=1e-12, atol=1e-12)
