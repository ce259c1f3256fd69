"""quantizedmha_b200 — B200-native (sm_100a) quantised multi-head-attention forward.

Host-side mirror of the reference's Python surface (extensions/torch/torch_ext.cpp:11-57 and
extensions/jax/jax_ext.cpp:12-36) on top of the C-ABI shared library declared in
include/qmha.h.  There is no CPU or PyTorch fallback: if libqmha.so is missing, or no B200 is
present, every compute call raises.
"""
from .binding import (  # noqa: F401
    QmhaArgs,
    QmhaError,
    GRAN_BLOCK,
    GRAN_HEAD,
    GRAN_TENSOR,
    DTYPE_BF16,
    DTYPE_F16,
    DTYPE_F32,
    KERNEL_BF16,
    KERNEL_F16,
    KERNEL_INT8,
    KERNEL_INT8_PV8,
    attention_prepared,
    convert_qkv_f16,
    flash_solve,
    flash_solve_ptr,
    forward,
    forward_host,
    get_rope,
    kernel_id,
    launch_count,
    lib,
    lib_path,
    quantize_blocks,
    quantize_qkv,
    quantize_static,
    set_rope,
    solve,
    workspace_dims,
)

__all__ = [n for n in dir() if not n.startswith("_")]
