"""The compiled Python entry points that keep the reference's names: extensions/torch (module
torch_ext, extensions/torch/tests/test_torch_bindings.py:11-31 in the reference) and
extensions/jax (module jax_ext, pointer ABI of jax_ext.cpp:12-36)."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def torch():
    import torch as t
    if not t.cuda.is_available():
        pytest.skip("no CUDA device")
    return t


def _import(ext, name):
    d = os.path.join(ROOT, "extensions", ext)
    if d not in sys.path:
        sys.path.insert(0, d)
    try:
        return __import__(name)
    except ImportError as e:  # not built: __graft_entry__.build() builds both
        pytest.fail(f"{name} is not built ({e}); run __graft_entry__.build()")


def test_torch_ext_flash_solve_like_the_reference_test(torch, oracle):
    torch_ext = _import("torch", "torch_ext")
    torch.manual_seed(42)
    N, d_model, num_heads = 256, 32, 4          # the reference test's shape
    Q, K, V = (torch.randn(N, d_model, device="cuda", dtype=torch.float32) for _ in range(3))
    out = torch_ext.flash_solve(Q, K, V, d_model, num_heads, kernel="fa_tc_int8_b")
    assert out.shape == (N, d_model) and out.dtype == torch.float32 and out.is_cuda
    ref = oracle.mha(Q.cpu().numpy(), K.cpu().numpy(), V.cpu().numpy(), num_heads, "f64")
    torch.cuda.synchronize()
    assert np.abs(out.cpu().numpy() - ref).max() <= 5e-2
    out16 = torch_ext.flash_solve(Q, K, V, d_model, num_heads, kernel="fa_tc_v2a")
    assert np.abs(out16.cpu().numpy() - ref).max() <= 2e-3
    with pytest.raises(RuntimeError):
        torch_ext.flash_solve(Q.double(), K, V, d_model, num_heads)
    # batched input is an extension over the reference
    Qb = torch.rand(2, 300, 128, device="cuda")
    ob = torch_ext.flash_solve(Qb, Qb, Qb, 128, 2)
    refb = oracle.mha(Qb.cpu().numpy(), Qb.cpu().numpy(), Qb.cpu().numpy(), 2, "f64")
    assert np.abs(ob.cpu().numpy() - refb).max() <= 2e-2


def test_jax_ext_pointer_abi(torch, oracle):
    jax_ext = _import("jax", "jax_ext")
    q, k, v = oracle.profile_inputs(512, 256)
    tq, tk, tv = (torch.from_numpy(a).cuda() for a in (q, k, v))
    out = torch.empty_like(tq)
    torch.cuda.synchronize()
    jax_ext.flash_solve(tq.data_ptr(), tk.data_ptr(), tv.data_ptr(), out.data_ptr(), 512, 256, 4, "fa_tc_int8_b")
    ref = oracle.mha(q, k, v, 4, "f64")
    assert np.abs(out.cpu().numpy() - ref).max() <= 2e-2
    jax_ext.flash_solve(tq.data_ptr(), tk.data_ptr(), tv.data_ptr(), out.data_ptr(), 512, 256, 4, "fa_tc_v2a")
    assert np.abs(out.cpu().numpy() - ref).max() <= 2e-3
    with pytest.raises(ValueError):
        jax_ext.flash_solve(0, 0, 0, 0, 8, 32, 4, "not_a_kernel")
    jax_ext.flash_solve(tq.data_ptr(), tk.data_ptr(), tv.data_ptr(), out.data_ptr(), 512, 256, 4)
