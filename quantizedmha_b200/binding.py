"""ctypes binding of quantizedmha_b200/lib/libqmha.so (include/qmha.h).

PyTorch appears here only for device memory, streams and pointers — the reference's own torch
extension does the same (extensions/torch/torch_ext.cpp:36-40 passes raw data_ptr()s to solve).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Tuple

KERNEL_INT8, KERNEL_F16, KERNEL_BF16, KERNEL_INT8_PV8 = 0, 1, 2, 3
DTYPE_F32, DTYPE_F16, DTYPE_BF16 = 0, 1, 2
GRAN_TENSOR, GRAN_HEAD, GRAN_BLOCK = 0, 1, 2

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB: Optional[C.CDLL] = None


class QmhaError(RuntimeError):
    pass


def lib_path() -> str:
    return os.environ.get("QMHA_LIB", os.path.join(_HERE, "lib", "libqmha.so"))


class QmhaArgs(C.Structure):
    """include/qmha.h: qmha_args"""
    _fields_ = [("struct_size", C.c_size_t), ("Q", C.c_void_p), ("K", C.c_void_p), ("V", C.c_void_p), ("O", C.c_void_p),
                ("B", C.c_int), ("N", C.c_int), ("d_model", C.c_int), ("h", C.c_int), ("kernel", C.c_int),
                ("gran", C.c_int), ("in_dtype", C.c_int), ("out_dtype", C.c_int), ("rope", C.c_int),
                ("rope_base", C.c_float), ("variant", C.c_int), ("stream", C.c_void_p),
                ("o_row_stride", C.c_int64), ("o_batch_stride", C.c_int64), ("n_peers", C.c_int),
                ("peer_O", C.c_void_p * 7), ("in_row_stride", C.c_int64), ("in_batch_stride", C.c_int64),
                ("device", C.c_int)]


MAX_PEERS = 7


def declare(L: C.CDLL) -> C.CDLL:
    """Sets the prototypes of include/qmha.h on a loaded library (also used by the A/B tools that
    load several builds of the library side by side)."""
    vp, i, f = C.c_void_p, C.c_int, C.c_float
    L.solve.argtypes = [vp, vp, vp, vp, i, i, i]
    L.solve.restype = None
    L.qmha_forward.argtypes = [vp, vp, vp, vp, i, i, i, i, i, i, vp]
    L.qmha_forward_host.argtypes = [vp, vp, vp, vp, i, i, i, i, i, i]
    L.qmha_forward_host_ex.argtypes = [vp, vp, vp, vp, i, i, i, i, i, i, i, i]
    L.qmha_workspace_dims.argtypes = [i, i, i, C.POINTER(i), C.POINTER(i)]
    L.qmha_quantize_qkv.argtypes = [vp, vp, vp, i, i, i, i, i, vp, vp, vp, vp, vp]
    L.qmha_convert_qkv_f16.argtypes = [vp, vp, vp, i, i, i, i, vp, vp, vp, vp]
    L.qmha_quantize_blocks.argtypes = [vp, i, i, i, i, i, vp, vp, vp]
    L.qmha_quantize_static.argtypes = [vp, C.c_int64, f, f, vp, vp]
    L.qmha_attention_prepared.argtypes = [vp, vp, vp, vp, vp, i, i, i, i, i, i, vp]
    L.qmha_quantize_qkv_ex.argtypes = [vp, vp, vp, i, i, i, i, i, i, i, f, vp, vp, vp, vp, vp]
    L.qmha_quantize_qkv_k.argtypes = [vp, vp, vp, i, i, i, i, i, i, i, i, f, vp, vp, vp, vp, vp]
    L.qmha_convert_qkv_16.argtypes = [vp, vp, vp, i, i, i, i, i, i, i, f, vp, vp, vp, vp]
    L.qmha_attention_prepared_ex.argtypes = [vp, vp, vp, vp, vp, i, i, i, i, i, i, i, vp]
    L.qmha_forward_ex.argtypes = [C.POINTER(QmhaArgs)]
    L.qmha_args_init.argtypes = [C.POINTER(QmhaArgs)]
    L.qmha_args_init.restype = None
    L.qmha_granularity_for.argtypes = [i, i, i]
    L.qmha_debug_inject_stall.argtypes = [i]
    L.qmha_synchronize.argtypes = [vp]
    L.qmha_check_async_error.argtypes = []
    L.qmha_last_error.restype = C.c_char_p
    L.qmha_set_kernel.argtypes = [C.c_char_p]
    L.qmha_get_kernel.restype = C.c_char_p
    L.qmha_kernel_from_name.argtypes = [C.c_char_p]
    L.qmha_default_granularity.argtypes = [i, i]
    L.qmha_set_rope.argtypes = [i, f]
    L.qmha_get_rope.argtypes = []
    if hasattr(L, "qmha_ipc_export"):     # (older builds loaded side by side by the A/B tools lack these)
        L.qmha_ipc_export.argtypes = [vp, C.c_char_p, C.POINTER(C.c_int64)]
        L.qmha_ipc_open.argtypes = [C.c_char_p, C.c_int64, C.POINTER(vp)]
        L.qmha_ipc_close_all.argtypes = []
        if hasattr(L, "qmha_ipc_close"):
            L.qmha_ipc_close.argtypes = [vp]
        L.qmha_enable_peer_access.argtypes = [i, i]
    L.qmha_launch_count.restype = C.c_int64
    L.qmha_version.restype = C.c_char_p
    L.qmha_shutdown.restype = None
    if hasattr(L, "qmha_debug_cycles"):
        L.qmha_debug_cycles.argtypes = [C.POINTER(C.c_ulonglong), i]
    return L


def lib() -> C.CDLL:
    """Loads the CUDA library; raises (never falls back) when it has not been built."""
    global _LIB
    if _LIB is None:
        p = lib_path()
        if not os.path.exists(p):
            raise QmhaError(f"{p} not found: build it with `make lib` or __graft_entry__.build(); "
                            "there is no fallback path")
        _LIB = declare(C.CDLL(p))
    return _LIB


def _check(rc: int) -> None:
    if rc != 0:
        raise QmhaError(lib().qmha_last_error().decode() or "qmha call failed")


def kernel_id(kernel) -> int:
    """Accepts KERNEL_* ints, 'int8'/'f16', or any of the reference's kernel names."""
    if isinstance(kernel, int):
        return kernel
    k = lib().qmha_kernel_from_name(str(kernel).encode())
    if k < 0:
        raise QmhaError(f"unknown kernel {kernel!r}")
    return k


def set_rope(enable: bool, base: float = 10000.0) -> None:
    """Fused RoPE on Q and K inside the quantise / convert pass (utils/verify.cu:9-23 semantics;
    the reference's CPU check rotates, its GPU kernels never did).  Process-wide."""
    _check(lib().qmha_set_rope(int(bool(enable)), float(base)))


def get_rope() -> bool:
    return bool(lib().qmha_get_rope())


def launch_count() -> int:
    return int(lib().qmha_launch_count())


def workspace_dims(N: int, d_model: int, h: int) -> Tuple[int, int]:
    n_pad, d_pad = C.c_int(), C.c_int()
    _check(lib().qmha_workspace_dims(N, d_model, h, C.byref(n_pad), C.byref(d_pad)))
    return n_pad.value, d_pad.value


def _torch():
    import torch
    return torch


def _stream_ptr(stream=None, device=None) -> int:
    torch = _torch()
    s = stream if stream is not None else torch.cuda.current_stream(device)
    return int(s.cuda_stream)


def _shape3(t) -> Tuple[int, int, int]:
    if t.dim() == 2:
        return 1, t.shape[0], t.shape[1]
    if t.dim() == 3:
        return t.shape[0], t.shape[1], t.shape[2]
    raise QmhaError("expected [N, d_model] or [B, N, d_model]")


def _dtype_id(dt) -> int:
    torch = _torch()
    try:
        return {torch.float32: DTYPE_F32, torch.float16: DTYPE_F16, torch.bfloat16: DTYPE_BF16}[dt]
    except KeyError:
        raise QmhaError(f"unsupported dtype {dt}: float32, float16 or bfloat16") from None


def _check_inputs(Q, K, V, allow16: bool = False):
    torch = _torch()
    for name, t in (("Q", Q), ("K", K), ("V", V)):
        if not t.is_cuda:
            raise QmhaError("Inputs must be CUDA tensors")  # torch_ext.cpp:14
        if t.dtype != torch.float32 and not (allow16 and t.dtype in (torch.float16, torch.bfloat16)):
            raise QmhaError(f"{name} must be float32")  # torch_ext.cpp:15-17
    if Q.shape != K.shape or Q.shape != V.shape:
        raise QmhaError("Q, K, V must have the same shape")
    if Q.dtype != K.dtype or Q.dtype != V.dtype:
        raise QmhaError("Q, K, V must have the same dtype")


def solve(Q, K, V, N: int, d_model: int, h: int, out=None):
    """The reference's C entry point on torch device tensors (synchronous, default variant)."""
    torch = _torch()
    out = torch.empty_like(Q) if out is None else out
    torch.cuda.current_stream().synchronize()
    lib().solve(Q.data_ptr(), K.data_ptr(), V.data_ptr(), out.data_ptr(), N, d_model, h)
    err = lib().qmha_last_error().decode()
    if err:
        raise QmhaError(err)
    return out


def _out_strides(out, B: int, N: int, d_model: int) -> Tuple[int, int]:
    """(row stride, batch stride) in elements of an output tensor that may be a (batch, head-range) slab view of
    a larger tensor: the last dimension must be contiguous."""
    if tuple(out.shape) not in ((N, d_model), (B, N, d_model)) or (out.dim() == 2 and B != 1):
        raise QmhaError(f"out has shape {tuple(out.shape)}, expected {(B, N, d_model)}")
    if d_model > 1 and out.stride(-1) != 1:
        raise QmhaError("the last dimension of out must be contiguous")
    ld = out.stride(-2) if N > 1 else d_model
    bs = out.stride(0) if (out.dim() == 3 and B > 1) else N * ld
    return int(ld), int(bs)


def forward(Q, K, V, num_heads: int, kernel="int8", gran: int = GRAN_HEAD, out=None, stream=None,
            rope=None, rope_base: float = 10000.0, out_dtype=None, peer_outs=None):
    """Stream-ordered forward on [N, d_model] or [B, N, d_model] CUDA tensors (qmha_forward_ex).
    Q, K, V: float32 like the reference, or float16 / bfloat16 (the quantise pass then reads 2 bytes per
    element); out_dtype: dtype of the result (default: that of Q).  rope: True / False per call, None = the
    process default (set_rope / QMHA_ROPE).  gran: GRAN_* or -1 = default for the shape.
    out may be a slab VIEW of a larger tensor (rows / batch entries strided, last dimension contiguous): the
    kernel's epilogue writes it in place.  peer_outs: up to 7 further destinations that receive the same bytes
    with the same strides — raw device addresses (e.g. other ranks' tensors mapped with ipc_open) or tensors
    shaped and strided like out (replicas on this or on peer-accessible devices)."""
    torch = _torch()
    _check_inputs(Q, K, V, allow16=True)
    B, N, d_model = _shape3(Q)
    # Q, K, V may be slab views (rows / batch entries strided, last dimension contiguous) with one common pitch and
    # 16-byte aligned origins: the quantise pass reads them in place; anything else is made contiguous first
    in_ld = in_bs = 0
    esz = Q.element_size()
    try:
        pitches = {_out_strides(t, B, N, d_model) for t in (Q, K, V)}
    except QmhaError:
        pitches = set()
    if (len(pitches) == 1 and all(t.data_ptr() % 16 == 0 for t in (Q, K, V))
            and all((p * esz) % 16 == 0 for p in next(iter(pitches)))):
        in_ld, in_bs = next(iter(pitches))
        if (in_ld, in_bs) == (d_model, N * d_model):
            in_ld = in_bs = 0
    else:
        Q, K, V = Q.contiguous(), K.contiguous(), V.contiguous()
    if out is None:
        out = torch.empty(Q.shape, dtype=out_dtype or Q.dtype, device=Q.device)
    ld, bs = _out_strides(out, B, N, d_model)
    a = QmhaArgs()
    lib().qmha_args_init(C.byref(a))
    a.Q, a.K, a.V, a.O = Q.data_ptr(), K.data_ptr(), V.data_ptr(), out.data_ptr()
    a.B, a.N, a.d_model, a.h = B, N, d_model, num_heads
    a.kernel, a.gran = kernel_id(kernel), gran
    a.in_dtype, a.out_dtype = _dtype_id(Q.dtype), _dtype_id(out.dtype)
    a.rope = -1 if rope is None else int(bool(rope))
    a.rope_base = float(rope_base)
    a.stream = _stream_ptr(stream, Q.device)           # the tensors' device, which need not be the current one
    a.device = Q.device.index if Q.device.index is not None else -1
    if (ld, bs) != (d_model, N * d_model):
        a.o_row_stride, a.o_batch_stride = ld, bs
    peers = list(peer_outs or [])
    if len(peers) > MAX_PEERS:
        raise QmhaError(f"at most {MAX_PEERS} peer outputs")
    for j, pz in enumerate(peers):
        if isinstance(pz, int):
            a.peer_O[j] = pz
        else:
            if pz.dtype != out.dtype or _out_strides(pz, B, N, d_model) != (ld, bs):
                raise QmhaError("peer outputs must have the dtype, shape and strides of out")
            a.peer_O[j] = pz.data_ptr()
    a.n_peers = len(peers)
    a.in_row_stride, a.in_batch_stride = in_ld, in_bs
    _check(lib().qmha_forward_ex(C.byref(a)))
    return out


def ipc_export(t) -> Tuple[bytes, int]:
    """(64-byte CUDA IPC handle of the allocation that holds tensor t, byte offset of t inside it) — qmha_ipc_export."""
    h = C.create_string_buffer(64)
    off = C.c_int64()
    _check(lib().qmha_ipc_export(t.data_ptr(), h, C.byref(off)))
    return h.raw, int(off.value)


def ipc_open(handle: bytes, offset: int) -> int:
    """Device address, valid in THIS process on the current device, of another rank's exported tensor."""
    if len(handle) != 64:
        raise QmhaError("a CUDA IPC handle is 64 bytes")
    p = C.c_void_p()
    _check(lib().qmha_ipc_open(C.create_string_buffer(handle, 64), offset, C.byref(p)))
    return int(p.value)


def ipc_close(addr: int) -> None:
    """Drops one reference to the mapping behind an address returned by ipc_open."""
    _check(lib().qmha_ipc_close(addr))


def ipc_close_all() -> None:
    _check(lib().qmha_ipc_close_all())


def enable_peer_access(dev: int, peer: int) -> None:
    """Several devices in one process: lets kernels on `dev` write tensors that live on `peer`."""
    _check(lib().qmha_enable_peer_access(dev, peer))


def flash_solve(Q, K, V, d_model: int, num_heads: int, kernel: str = "fa_tc_int8_b"):
    """Mirror of torch_ext.flash_solve (extensions/torch/torch_ext.cpp:11-43): fp32 CUDA tensors
    whose numel is a multiple of d_model; returns a tensor like Q.  Unlike the reference, `kernel`
    actually selects the variant and the work is enqueued on torch's current stream."""
    torch = _torch()
    _check_inputs(Q, K, V)
    Qc, Kc, Vc = Q.contiguous(), K.contiguous(), V.contiguous()
    if Qc.numel() % d_model != 0:
        raise QmhaError("Q.numel() must be divisible by d_model")  # torch_ext.cpp:24
    if Qc.dim() == 3:
        B, N = Qc.shape[0], Qc.shape[1]
    else:
        B, N = 1, Qc.numel() // d_model
    out = torch.empty_like(Qc)
    # gran -1: the same choice as the C solve() (block scales when they fit, per-head otherwise)
    _check(lib().qmha_forward(Qc.data_ptr(), Kc.data_ptr(), Vc.data_ptr(), out.data_ptr(), B, N, d_model,
                              num_heads, kernel_id(kernel), -1, _stream_ptr()))
    return out


def flash_solve_ptr(q_ptr: int, k_ptr: int, v_ptr: int, out_ptr: int, N: int, d_model: int,
                    num_heads: int, kernel: str = "fa_tc_int8_b") -> None:
    """Mirror of jax_ext.flash_solve (extensions/jax/jax_ext.cpp:12-28): raw device addresses."""
    L = lib()
    _check(L.qmha_forward(q_ptr, k_ptr, v_ptr, out_ptr, 1, N, d_model, num_heads, kernel_id(kernel), -1, None))
    _torch().cuda.synchronize()
    _check(L.qmha_check_async_error())


def forward_host(Q, K, V, num_heads: int, kernel="int8", gran: int = GRAN_HEAD, out=None):
    """Host (CPU, ideally pinned) fp32 / fp16 / bf16 tensors in, host tensor out (dtype of `out`, default that of
    Q); copies are pipelined over (batch entry, head group) chunks inside."""
    torch = _torch()
    B, N, d_model = _shape3(Q)
    out = torch.empty_like(Q) if out is None else out
    _check(lib().qmha_forward_host_ex(Q.data_ptr(), K.data_ptr(), V.data_ptr(), out.data_ptr(), B, N, d_model,
                                      num_heads, kernel_id(kernel), gran, _dtype_id(Q.dtype), _dtype_id(out.dtype)))
    return out


def quantize_qkv(Q, K, V, num_heads: int, gran: int = GRAN_HEAD, stream=None, rope=None, rope_base: float = 10000.0,
                 kernel="int8"):
    """Kernel (a).  Returns (Qp int8 [B*h,n_pad,d_pad], Kp, Vt fp16 [B*h,d_pad,n_pad], scales [3,B*h]
    or, for GRAN_BLOCK, [3,B*h,n_pad/32]).  Inputs: float32, float16 or bfloat16.  kernel="int8_pv8": Vt holds
    the int8 codes themselves (int8 tensor) for the INT8 P.V mode."""
    torch = _torch()
    _check_inputs(Q, K, V, allow16=True)
    Q, K, V = Q.contiguous(), K.contiguous(), V.contiguous()
    B, N, d_model = _shape3(Q)
    n_pad, d_pad = workspace_dims(N, d_model, num_heads)
    u = B * num_heads
    Qp = torch.empty((u, n_pad, d_pad), dtype=torch.int8, device=Q.device)
    Kp = torch.empty_like(Qp)
    kid = kernel_id(kernel)
    Vt = torch.empty((u, d_pad, n_pad), dtype=torch.int8 if kid == KERNEL_INT8_PV8 else torch.float16, device=Q.device)
    if gran == GRAN_BLOCK:
        scales = torch.empty((3, u, n_pad // 32), dtype=torch.float32, device=Q.device)
    else:
        scales = torch.empty((3, u), dtype=torch.float32, device=Q.device)
    _check(lib().qmha_quantize_qkv_k(Q.data_ptr(), K.data_ptr(), V.data_ptr(), _dtype_id(Q.dtype), B, N, d_model,
                                     num_heads, kid, gran, -1 if rope is None else int(bool(rope)), float(rope_base),
                                     Qp.data_ptr(), Kp.data_ptr(), Vt.data_ptr(), scales.data_ptr(),
                                     _stream_ptr(stream)))
    return Qp, Kp, Vt, scales


def convert_qkv_f16(Q, K, V, num_heads: int, stream=None, kernel="f16"):
    """Operands of the 16-bit kernels: fp16 (kernel="f16") or bf16 (kernel="bf16") tensors in the prepared layout."""
    torch = _torch()
    _check_inputs(Q, K, V, allow16=True)
    Q, K, V = Q.contiguous(), K.contiguous(), V.contiguous()
    B, N, d_model = _shape3(Q)
    n_pad, d_pad = workspace_dims(N, d_model, num_heads)
    u = B * num_heads
    kid = kernel_id(kernel)
    dt = torch.bfloat16 if kid == KERNEL_BF16 else torch.float16
    Qp = torch.empty((u, n_pad, d_pad), dtype=dt, device=Q.device)
    Kp = torch.empty_like(Qp)
    Vt = torch.empty((u, d_pad, n_pad), dtype=dt, device=Q.device)
    _check(lib().qmha_convert_qkv_16(Q.data_ptr(), K.data_ptr(), V.data_ptr(), _dtype_id(Q.dtype), B, N, d_model,
                                     num_heads, kid, -1, 0.0, Qp.data_ptr(), Kp.data_ptr(), Vt.data_ptr(),
                                     _stream_ptr(stream)))
    return Qp, Kp, Vt


def attention_prepared(Qp, Kp, Vt, scales, B: int, N: int, d_model: int, num_heads: int, kernel="int8",
                       out=None, stream=None, gran: int = GRAN_HEAD):
    """Kernel (b)/(c) on prepared operands; returns O [B, N, d_model] fp32.  `gran` must be the
    granularity the scales were produced with (GRAN_BLOCK: scales [3, B*h, n_pad/32])."""
    torch = _torch()
    out = torch.empty((B, N, d_model), dtype=torch.float32, device=Qp.device) if out is None else out
    _check(lib().qmha_attention_prepared(Qp.data_ptr(), Kp.data_ptr(), Vt.data_ptr(),
                                         scales.data_ptr() if scales is not None else None, out.data_ptr(),
                                         B, N, d_model, num_heads, kernel_id(kernel), gran, _stream_ptr(stream)))
    return out


def check_async_error() -> None:
    _check(lib().qmha_check_async_error())


def quantize_blocks(X, num_heads: int, block_rows: int = 32, stream=None):
    """Reference-granularity quantiser (one scale per 32-row block per head), input layout."""
    torch = _torch()
    X = X.contiguous()
    B, N, d_model = _shape3(X)
    nblk = -(-N // block_rows)
    q = torch.empty(X.shape, dtype=torch.int8, device=X.device)
    s = torch.empty((B * num_heads * nblk,), dtype=torch.float32, device=X.device)
    _check(lib().qmha_quantize_blocks(X.data_ptr(), B, N, d_model, num_heads, block_rows, q.data_ptr(),
                                      s.data_ptr(), _stream_ptr(stream)))
    return q, s


def quantize_static(X, scale: float, zero_point: float = 0.0, stream=None):
    """Golden-spec quantiser (generate_golden.cpp:94-101)."""
    torch = _torch()
    X = X.contiguous()
    q = torch.empty(X.shape, dtype=torch.int8, device=X.device)
    _check(lib().qmha_quantize_static(X.data_ptr(), X.numel(), scale, zero_point, q.data_ptr(),
                                      _stream_ptr(stream)))
    return q
