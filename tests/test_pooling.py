"""gnn_b200.MaxPoolLocal vs the reference layer (alegnn/utils/graphML.py:1850-2028), live on CPU."""
import numpy as np
import pytest
import torch

import lsigf_oracle as orc
import ref_import

pytestmark = pytest.mark.skipif(not ref_import.reference_available(), reason="reference tree not present")


@pytest.mark.parametrize("E,K,Nin,Nout", [(1, 1, 20, 20), (1, 2, 20, 9), (2, 3, 17, 5), (1, 0, 12, 7)])
def test_max_pool_local_matches_reference(E, K, Nin, Nout, monkeypatch):
    gml = ref_import.import_reference()
    import gnn_b200
    from gnn_b200 import pooling
    from gnn_b200.pooling import MaxPoolLocal

    def gather_max(x, nb32, n_out, max_nb):          # torch stand-in for the CUDA gather (csrc/layer.cu); CPU leg only
        B, F, _ = x.shape
        return x.index_select(2, nb32.reshape(-1).long()).reshape(B, F, n_out, max_nb).max(dim=3)[0]

    monkeypatch.setattr(pooling, "_gather_max", gather_max)
    rng = np.random.default_rng(E * 100 + K)
    N = Nin
    S = np.abs(orc.random_sparse_gso(rng, N, 3, E, symmetric=True))     # the reference keeps entries > 1e-9 only
    x = torch.tensor(rng.standard_normal((3, 4, Nin)), requires_grad=True)
    ref = gml.MaxPoolLocal(Nin, Nout, K)
    ref.addGSO(torch.tensor(S))
    mine = MaxPoolLocal(Nin, Nout, K)
    mine.addGSO(torch.tensor(S))
    assert mine.maxNeighborhoodSize == ref.maxNeighborhoodSize
    assert torch.equal(mine.neighborhood.sort(dim=1)[0], ref.neighborhood.to(mine.neighborhood.dtype).sort(dim=1)[0])
    y_ref = ref(x)
    y = mine(x)
    assert torch.equal(y, y_ref)
    g = torch.tensor(rng.standard_normal(tuple(y.shape)))
    (gx_ref,) = torch.autograd.grad(y_ref, x, g, retain_graph=True)
    (gx,) = torch.autograd.grad(y, x, g)
    assert torch.allclose(gx, gx_ref)
    # the sparse description gives the same layer, and it works on a node-major strided view (what LSIGF returns)
    sparse = MaxPoolLocal(Nin, Nout, K)
    sparse.addGSO(gnn_b200.SparseGSO.from_dense(torch.tensor(S)))
    buf = x.detach().permute(2, 0, 1).contiguous()                       # [N, B, F] node-major
    assert torch.equal(sparse(buf.permute(1, 2, 0)), y_ref.detach())
    assert "neighborhood stored" in mine.extra_repr()
