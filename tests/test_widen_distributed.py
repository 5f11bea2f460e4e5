"""Multi-GPU backward of the partitioned LSIGF over NCCL (2 GPUs): forward + dh / dx / db against the sparse oracle,
for the node sharding, the feature sharding with the NCCL all-to-all, and the feature sharding with the fused
hop + NVLink scatter forward.  The same choreography runs on gloo (CPU) in tests/test_distributed.py."""
import pytest
import torch

from test_distributed import _run


@pytest.mark.gpu
@pytest.mark.parametrize("mode,G", [("nodes", 6), ("features", 6), ("features", 16)])
@pytest.mark.parametrize("dtype_name,tol", [("float32", 1e-4), ("float64", 1e-11)])
def test_partitioned_backward_nccl_world2(mode, G, dtype_name, tol):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    err = _run("nccl", mode, dtype_name, G=G, backward=True)
    assert err < tol, err
