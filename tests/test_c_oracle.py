"""The plain-C oracle (oracle/lsigf_oracle.c) against the reference fixtures and against the numpy oracle. CPU only."""
import os

import numpy as np
import pytest

import c_oracle
import lsigf_oracle as orc


def _rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.fixture(scope="module")
def z(golden_dir):
    return np.load(os.path.join(golden_dir, "lsigf_cases.npz"))


def test_c_oracle_matches_reference_fixtures(z):
    keys = sorted({k.split("_")[0] for k in z.files if k.startswith("c")})
    assert len(keys) >= 8
    for c in keys:
        b = z[c + "_b"] if (c + "_b") in z.files else None
        y = c_oracle.lsigf_forward(z[c + "_h"], list(z[c + "_S"]), z[c + "_x"], b)
        assert _rel(y, z[c + "_y"]) < 1e-12, c
        dh, dx, db = c_oracle.lsigf_backward(z[c + "_h"], list(z[c + "_S"]), z[c + "_x"], z[c + "_dy"],
                                             None if b is None else b.shape)
        assert _rel(dh, z[c + "_dh"]) < 1e-12 and _rel(dx, z[c + "_dx"]) < 1e-12, c
        if b is not None:
            assert _rel(db, z[c + "_db"]) < 1e-12, c


def test_c_oracle_matches_numpy_oracle_on_a_sparse_graph():
    import scipy.sparse as sp
    rng = np.random.default_rng(4)
    N, B, G, F, K, E = 3000, 2, 5, 4, 4, 2
    mats = [sp.random(N, N, density=6.0 / N, format="csr", random_state=np.random.RandomState(e)) * 0.3 for e in range(E)]
    h = rng.standard_normal((F, E, K, G)) * 0.3
    x = rng.standard_normal((B, G, N))
    b = rng.standard_normal((F, N))
    dy = rng.standard_normal((B, F, N))
    assert _rel(c_oracle.lsigf_forward(h, mats, x, b), orc.lsigf_sparse(h, mats, x, b)) < 1e-12
    dh, dx, db = c_oracle.lsigf_backward(h, mats, x, dy, b.shape)
    dh2, dx2, db2 = orc.lsigf_grads_sparse(h, mats, x, dy, b.shape)
    assert _rel(dh, dh2) < 1e-12 and _rel(dx, dx2) < 1e-12 and _rel(db, db2) < 1e-12
