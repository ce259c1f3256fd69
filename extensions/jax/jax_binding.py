"""flash_solve_jax(q, k, v, d_model, num_heads, kernel) — JAX arrays in/out, like the reference's
extensions/jax/jax_binding.py:25-77 (JAX -> DLPack -> CuPy -> raw pointers -> jax_ext).  Needs
jax and cupy at run time (neither ships in this image; the pointer ABI itself is tested with
torch pointers in tests/test_gpu_parity.py)."""


def flash_solve_jax(q, k, v, d_model: int, num_heads: int, kernel: str = "fa_tc_int8_b"):
    import cupy as cp
    import jax.dlpack as jdl

    import jax_ext

    def to_cp(x):
        return cp.ascontiguousarray(cp.from_dlpack(jdl.to_dlpack(x)).astype(cp.float32))

    cq, ck, cv = to_cp(q), to_cp(k), to_cp(v)
    out = cp.empty_like(cq)
    n = cq.size // d_model
    jax_ext.flash_solve(int(cq.data.ptr), int(ck.data.ptr), int(cv.data.ptr), int(out.data.ptr),
                        int(n), int(d_model), int(num_heads), kernel)
    return jdl.from_dlpack(out.toDlpack())
