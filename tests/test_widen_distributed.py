"""Multi-GPU backward of the partitioned LSIGF over NCCL (2 GPUs): forward + dh / dx / db against the sparse oracle,
for the node sharding, the feature sharding with the NCCL all-to-all, and the feature sharding with the fused
hop + NVLink scatter forward.  The same choreography runs on gloo (CPU) in tests/test_distributed.py."""
import pytest
import torch

from test_distributed import _run


@pytest.mark.gpu
@pytest.mark.parametrize("mode,G,F", [("nodes", 6, 8), ("nodes", 48, 40), ("features", 6, 8), ("features", 16, 8)])
@pytest.mark.parametrize("dtype_name,tol", [("float32", 1e-4), ("float64", 1e-11)])
def test_partitioned_backward_nccl_world2(mode, G, F, dtype_name, tol):
    """nodes / G = 48, F = 40: forward AND backward shift chains run the fused hop + all-gather kernel."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    err = _run("nccl", mode, dtype_name, G=G, backward=True, F=F)
    assert err < tol, err
