"""ctypes bindings for oracle/libqmha_oracle.so (our CPU restatement) and, when present,
oracle/_ref/libqmha_ref.so (the reference's own host sources compiled where they lie).

TEST INFRASTRUCTURE ONLY — see the header of qmha_oracle.cpp.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_i8p = np.ctypeslib.ndpointer(dtype=np.int8, flags="C_CONTIGUOUS")


def build_oracle(with_ref: bool = True) -> None:
    """Compile the oracle (and oracle/_ref when /root/reference exists).  Building the checker
    is not using it."""
    subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    if with_ref and os.path.isdir("/root/reference"):
        subprocess.run(["make", "-C", _HERE, "-s", "ref"], check=True)


class Oracle:
    """Restatement of the reference CPU path (file:line citations in qmha_oracle.cpp)."""

    def __init__(self, path: Optional[str] = None):
        path = path or os.path.join(_HERE, "libqmha_oracle.so")
        if not os.path.exists(path):
            build_oracle(with_ref=False)
        self.lib = L = C.CDLL(path)
        L.oracle_init_profile_inputs.argtypes = [_f32p, _f32p, _f32p, C.c_int64, C.c_int]
        L.oracle_init_golden_inputs.argtypes = [_f32p, _f32p, _f32p, C.c_int, C.c_int, C.c_int, C.c_int]
        L.oracle_apply_rope.argtypes = [_f32p, C.c_int, C.c_int, C.c_int]
        L.oracle_mha_forward.argtypes = [_f32p, _f32p, _f32p, _f32p] + [C.c_int] * 6
        L.oracle_mha_head_rows.argtypes = [_f32p, _f32p, _f32p, _f32p] + [C.c_int] * 5
        L.oracle_cpu_reference_rope.argtypes = [_f32p, _f32p, _f32p, _f32p] + [C.c_int] * 4
        L.oracle_verify_results.argtypes = [_f32p, _f32p, C.c_int64, C.c_float, C.c_float, C.POINTER(C.c_int64)]
        L.oracle_verify_results.restype = C.c_int
        L.oracle_quantize_dynamic.argtypes = [_f32p] + [C.c_int] * 6 + [_i8p, _f32p]
        L.oracle_quantize_dynamic.restype = C.c_int64
        L.oracle_quantize_static.argtypes = [_f32p, C.c_int64, C.c_float, C.c_float, _i8p]
        L.oracle_mha_int8_emulated.argtypes = [_i8p, _i8p, _i8p, _f32p, _f32p, _f32p, _f32p] + [C.c_int] * 6
        L.oracle_mha_int8_emulated_block.argtypes = [_i8p, _i8p, _i8p, _f32p, _f32p, _f32p, _f32p] + [C.c_int] * 7
        L.oracle_save_reference.argtypes = [C.c_char_p, _f32p, C.c_int, C.c_int]
        L.oracle_save_reference.restype = C.c_int
        L.oracle_load_reference.argtypes = [C.c_char_p, _f32p, C.c_int, C.c_int]
        L.oracle_load_reference.restype = C.c_int
        L.oracle_num_threads.restype = C.c_int

    # -- inputs ----------------------------------------------------------------------------
    def profile_inputs(self, rows: int, d_model: int, use_random: bool = True):
        """inputs/data.cu:9-30.  rows = B*N (the stream continues across batches)."""
        q, k, v = (np.empty((rows, d_model), np.float32) for _ in range(3))
        self.lib.oracle_init_profile_inputs(q, k, v, rows * d_model, int(use_random))
        return q, k, v

    def golden_inputs(self, N: int, d_model: int, h: int, rope: bool = True):
        """tests/generate_golden.cpp:123-138."""
        q, k, v = (np.empty((N, d_model), np.float32) for _ in range(3))
        self.lib.oracle_init_golden_inputs(q, k, v, N, d_model, h, int(rope))
        return q, k, v

    def apply_rope(self, x: np.ndarray, h: int) -> np.ndarray:
        x = np.ascontiguousarray(x, np.float32).copy()
        N, d_model = x.shape
        self.lib.oracle_apply_rope(x, N, d_model, h)
        return x

    # -- attention -------------------------------------------------------------------------
    def mha(self, q, k, v, h: int, precision: str = "f32", threads: int = 0) -> np.ndarray:
        """softmax(QK^T/sqrt(d))V per head; q,k,v [N,d_model] or [B,N,d_model]."""
        q, k, v = (np.ascontiguousarray(a, np.float32) for a in (q, k, v))
        shp = q.shape
        B = 1 if q.ndim == 2 else shp[0]
        N, d_model = shp[-2], shp[-1]
        o = np.empty_like(q)
        self.lib.oracle_mha_forward(q, k, v, o, B, N, d_model, h, 0 if precision == "f32" else 1, threads)
        return o

    def mha_head_rows(self, q_rows, k, v, precision: str = "f64", threads: int = 0) -> np.ndarray:
        """A sample of query rows [nr, d] of one head against its full K / V [N, d]."""
        q_rows, k, v = (np.ascontiguousarray(a, np.float32) for a in (q_rows, k, v))
        nr, d = q_rows.shape
        assert k.shape == v.shape and k.shape[1] == d
        o = np.empty_like(q_rows)
        self.lib.oracle_mha_head_rows(q_rows, k, v, o, nr, k.shape[0], d, 0 if precision == "f32" else 1, threads)
        return o

    def cpu_reference_rope(self, q, k, v, h: int, threads: int = 0) -> np.ndarray:
        """utils/verify.cu:25-104 (attention with RoPE on q and k)."""
        q, k, v = (np.ascontiguousarray(a, np.float32) for a in (q, k, v))
        N, d_model = q.shape
        o = np.empty_like(q)
        self.lib.oracle_cpu_reference_rope(q, k, v, o, N, d_model, h, threads)
        return o

    def verify_results(self, out, ref, eps: float = 1e-3, rel: float = 1e-3) -> bool:
        out = np.ascontiguousarray(out, np.float32).ravel()
        ref = np.ascontiguousarray(ref, np.float32).ravel()
        if out.size != ref.size:
            return False
        return bool(self.lib.oracle_verify_results(out, ref, out.size, eps, rel, None))

    # -- quantisation ----------------------------------------------------------------------
    GRAN = {"tensor": 0, "head": 1, "block": 2}

    def quantize(self, x, h: int, gran: str = "head", block_rows: int = 32):
        """Kernel spec (fa_tc_int8_b.cu:104-106,138-140).  Returns (int8 like x, scales)."""
        x = np.ascontiguousarray(x, np.float32)
        B = 1 if x.ndim == 2 else x.shape[0]
        N, d_model = x.shape[-2], x.shape[-1]
        g = self.GRAN[gran]
        ns = 1 if g == 0 else B * h * (1 if g == 1 else -(-N // block_rows))
        q = np.empty(x.shape, np.int8)
        s = np.empty(ns, np.float32)
        got = self.lib.oracle_quantize_dynamic(x, B, N, d_model, h, g, block_rows, q, s)
        assert got == ns
        return q, s

    def quantize_static(self, x, scale: float, zero_point: float = 0.0) -> np.ndarray:
        """Golden spec (generate_golden.cpp:94-101)."""
        x = np.ascontiguousarray(x, np.float32)
        q = np.empty(x.shape, np.int8)
        self.lib.oracle_quantize_static(x.ravel(), x.size, scale, zero_point, q.ravel())
        return q

    def mha_int8_emulated(self, qq, kq, vq, sq, sk, sv, h: int, p_format: str = "f16", threads: int = 0):
        """Attention on given int8 tensors with per-(batch,head) scales; float64 softmax."""
        qq, kq, vq = (np.ascontiguousarray(a, np.int8) for a in (qq, kq, vq))
        sq, sk, sv = (np.ascontiguousarray(a, np.float32).ravel() for a in (sq, sk, sv))
        B = 1 if qq.ndim == 2 else qq.shape[0]
        N, d_model = qq.shape[-2], qq.shape[-1]
        assert sq.size == B * h and sk.size == B * h and sv.size == B * h
        o = np.empty(qq.shape, np.float32)
        self.lib.oracle_mha_int8_emulated(qq, kq, vq, sq, sk, sv, o, B, N, d_model, h,
                                          1 if p_format == "f16" else 0, threads)
        return o

    def mha_int8_emulated_block(self, qq, kq, vq, sq, sk, sv, h: int, block_rows: int = 32,
                                p_format: str = "f16", threads: int = 0):
        """Same with the reference's per-(head, 32-row block) scales; scales are [B*h*nblk]."""
        qq, kq, vq = (np.ascontiguousarray(a, np.int8) for a in (qq, kq, vq))
        sq, sk, sv = (np.ascontiguousarray(a, np.float32).ravel() for a in (sq, sk, sv))
        B = 1 if qq.ndim == 2 else qq.shape[0]
        N, d_model = qq.shape[-2], qq.shape[-1]
        nblk = -(-N // block_rows)
        assert sq.size == B * h * nblk and sk.size == sq.size and sv.size == sq.size
        o = np.empty(qq.shape, np.float32)
        self.lib.oracle_mha_int8_emulated_block(qq, kq, vq, sq, sk, sv, o, B, N, d_model, h, block_rows,
                                                1 if p_format == "f16" else 0, threads)
        return o

    def num_threads(self) -> int:
        return int(self.lib.oracle_num_threads())


class RefLib:
    """The reference's own host code (tests/generate_golden.cpp, utils/verify.cu,
    inputs/data.cu) compiled from /root/reference into oracle/_ref/libqmha_ref.so."""

    def __init__(self, path: Optional[str] = None):
        path = path or os.path.join(_HERE, "_ref", "libqmha_ref.so")
        self.lib = L = C.CDLL(path)
        L.ref_cpu_mha.argtypes = [_f32p, _f32p, _f32p, _f32p, C.c_int, C.c_int, C.c_int]
        L.ref_quantize_int8.argtypes = [_f32p, C.c_longlong, C.c_float, C.c_float, _i8p]
        L.ref_cpu_reference.argtypes = [_f32p, _f32p, _f32p, _f32p, C.c_int, C.c_int, C.c_int]
        L.ref_verify_results.argtypes = [_f32p, _f32p, C.c_longlong, C.c_float, C.c_float]
        L.ref_verify_results.restype = C.c_int
        L.ref_initialize_host_data.argtypes = [_f32p, _f32p, _f32p, C.c_int, C.c_int, C.c_int]
        L.ref_apply_rope_row.argtypes = [_f32p, C.c_int, C.c_int]
        L.ref_generate_golden_main.restype = C.c_int
        L.ref_save_reference.argtypes = [_f32p, C.c_char_p, C.c_int, C.c_int]
        L.ref_save_reference.restype = C.c_int
        L.ref_load_reference.argtypes = [_f32p, C.c_char_p, C.c_int, C.c_int]
        L.ref_load_reference.restype = C.c_int
        if hasattr(L, "ref_load_inputs"):
            L.ref_save_inputs.argtypes = [_f32p, _f32p, _f32p, C.c_char_p, C.c_int, C.c_int]
            L.ref_save_inputs.restype = C.c_int
            L.ref_load_inputs.argtypes = [_f32p, _f32p, _f32p, C.c_char_p, C.c_int, C.c_int]
            L.ref_load_inputs.restype = C.c_int

    def cpu_mha(self, q, k, v, h):
        q, k, v = (np.ascontiguousarray(a, np.float32) for a in (q, k, v))
        o = np.empty_like(q)
        self.lib.ref_cpu_mha(q, k, v, o, q.shape[0], q.shape[1], h)
        return o

    def cpu_reference(self, q, k, v, h):
        q, k, v = (np.ascontiguousarray(a, np.float32) for a in (q, k, v))
        o = np.empty_like(q)
        self.lib.ref_cpu_reference(q, k, v, o, q.shape[0], q.shape[1], h)
        return o

    def quantize_int8(self, x, scale, zp=0.0):
        x = np.ascontiguousarray(x, np.float32)
        q = np.empty(x.shape, np.int8)
        self.lib.ref_quantize_int8(x.ravel(), x.size, scale, zp, q.ravel())
        return q

    def initialize_host_data(self, N, d_model, use_random=True):
        q, k, v = (np.empty((N, d_model), np.float32) for _ in range(3))
        self.lib.ref_initialize_host_data(q, k, v, N, d_model, int(use_random))
        return q, k, v

    def verify_results(self, out, ref, eps=1e-3, rel=1e-3):
        out = np.ascontiguousarray(out, np.float32).ravel()
        ref = np.ascontiguousarray(ref, np.float32).ravel()
        return bool(self.lib.ref_verify_results(out, ref, out.size, eps, rel))

    def load_inputs(self, path, N, d_model):
        """inputs/data.cu:84-109 load_inputs; None when the file is missing or its header does not match."""
        q, k, v = (np.empty((N, d_model), np.float32) for _ in range(3))
        ok = self.lib.ref_load_inputs(q, k, v, str(path).encode(), N, d_model)
        return (q, k, v) if ok else None

    def save_inputs(self, q, k, v, path):
        q, k, v = (np.ascontiguousarray(a, np.float32) for a in (q, k, v))
        return bool(self.lib.ref_save_inputs(q, k, v, str(path).encode(), q.shape[0], q.shape[1]))

    def load_reference(self, path, N, d_model):
        """utils/verify.cu:128-151 load_reference."""
        o = np.empty((N, d_model), np.float32)
        return o if self.lib.ref_load_reference(o, str(path).encode(), N, d_model) else None

    def save_reference(self, data, path):
        data = np.ascontiguousarray(data, np.float32)
        return bool(self.lib.ref_save_reference(data, str(path).encode(), data.shape[0], data.shape[1]))

    def apply_rope_row(self, row, pos):
        row = np.ascontiguousarray(row, np.float32).copy()
        self.lib.ref_apply_rope_row(row, pos, row.size)
        return row


_ORACLE: Optional[Oracle] = None


def load_oracle() -> Oracle:
    global _ORACLE
    if _ORACLE is None:
        _ORACLE = Oracle()
    return _ORACLE


def load_ref() -> Optional[RefLib]:
    """None when oracle/_ref has not been built (no /root/reference and no prebuilt .so)."""
    p = os.path.join(_HERE, "_ref", "libqmha_ref.so")
    return RefLib(p) if os.path.exists(p) else None
