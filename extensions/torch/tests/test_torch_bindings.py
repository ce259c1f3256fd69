"""Smoke test with the reference's shape (extensions/torch/tests/test_torch_bindings.py:11-31);
the value checks live in tests/test_gpu_extensions.py at the repository root."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def test_flash_solve_shape_dtype_device():
    if not torch.cuda.is_available():
        import pytest
        pytest.skip("CUDA not available")
    import torch_ext
    torch.manual_seed(42)
    N, d_model, num_heads = 256, 32, 4
    Q, K, V = (torch.randn(N, d_model, device="cuda") for _ in range(3))
    out = torch_ext.flash_solve(Q, K, V, d_model, num_heads)
    assert out.shape == (N, d_model) and out.dtype == torch.float32 and out.is_cuda
