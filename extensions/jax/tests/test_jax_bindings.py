"""Smoke test mirroring the reference's extensions/jax/tests/test_jax_bindings.py:13-34: skips
without jax / a GPU; the pointer ABI is value-checked in tests/test_gpu_extensions.py."""
import os
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def test_flash_solve_jax_shape_dtype():
    jax = pytest.importorskip("jax")
    pytest.importorskip("cupy")
    if not any(d.platform == "gpu" for d in jax.devices()):
        pytest.skip("no GPU")
    import jax.numpy as jnp
    from jax_binding import flash_solve_jax
    key = jax.random.PRNGKey(0)
    q = jax.random.normal(key, (256, 32), dtype=jnp.float32)
    out = flash_solve_jax(q, q, q, 32, 4)
    assert out.shape == (256, 32) and out.dtype == jnp.float32
