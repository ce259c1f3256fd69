"""INT8 P.V mode (QMHA_KERNEL_INT8_PV8, "int8_pv8"): the reference's P.V semantics — P quantised to 8-bit codes and
multiplied with the int8 V codes on the INT8 tensor pipe, int32 accumulation (fa_tc_int8_b.cu:359-371) — as an opt-in
variant of the INT8 kernel (run with -m gpu on a B200).

Tolerances: the INT8 contract (max-abs <= 2e-2, rel-L2 <= 1e-2 vs the float64 oracle on inputs/data.cu inputs).  The
8-bit P adds its own rounding on top of the Q/K/V quantisation: against the emulated-INT8 model with exact P on the SAME
codes the mode is gated at rel-L2 <= 6e-3 (fp16 P: 2e-3)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

INT8_MAX_ABS, INT8_REL_L2 = 2e-2, 1e-2
PV8_VS_EMU_REL_L2 = 6e-3


@pytest.fixture(scope="module")
def torch():
    import torch as t
    if not t.cuda.is_available():
        pytest.skip("no CUDA device")
    return t


@pytest.fixture(scope="module")
def qm(torch):
    import quantizedmha_b200 as q
    return q


def _err(got, ref):
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    assert np.isfinite(got).all()
    return float(np.abs(got - ref).max()), float(np.linalg.norm(got - ref) / max(np.linalg.norm(ref), 1e-30))


def _dev(torch, *arrs):
    return [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in arrs]


def _unpack_rows(Qp, B, N, h, d):
    a = Qp.cpu().numpy().reshape(B, h, Qp.shape[1], Qp.shape[2])[:, :, :N, :d]
    return np.ascontiguousarray(a.transpose(0, 2, 1, 3)).reshape(B, N, h * d)


def _unpack_vt(Vt, B, N, h, d):
    a = Vt.cpu().numpy().reshape(B, h, Vt.shape[1], Vt.shape[2])[:, :, :d, :N]
    return np.ascontiguousarray(a.transpose(0, 3, 1, 2)).reshape(B, N, h * d)


@pytest.mark.parametrize("shape", [(1, 8, 32, 4), (2, 300, 256, 2), (1, 1024, 512, 4), (3, 129, 96, 3), (1, 50, 40, 2)])
@pytest.mark.parametrize("gran_name", ["block", "head", "tensor"])
def test_int8_v_operand_is_bit_exact(qm, torch, oracle, shape, gran_name):
    """Vt of the INT8 P.V mode holds the int8 codes themselves, transposed (keys contiguous), zero padded; Q / K operands
    and scales are those of the default INT8 mode."""
    B, N, dm, h = shape
    d = dm // h
    gran = {"block": qm.GRAN_BLOCK, "head": qm.GRAN_HEAD, "tensor": qm.GRAN_TENSOR}[gran_name]
    q, k, v = (np.stack(x) for x in zip(*[oracle.golden_inputs(N, dm, h) for _ in range(B)]))
    v[1:] *= 0.3
    tq, tk, tv = _dev(torch, q, k, v)
    Qp, Kp, Vt, sc = qm.quantize_qkv(tq, tk, tv, h, gran, kernel="int8_pv8")
    Q0, K0, V0, s0 = qm.quantize_qkv(tq, tk, tv, h, gran, kernel="int8")
    torch.cuda.synchronize()
    assert Vt.dtype == torch.int8 and torch.equal(Qp, Q0) and torch.equal(Kp, K0) and torch.equal(sc, s0)
    codes, _ = oracle.quantize(v, h, gran_name, 32)
    assert np.array_equal(_unpack_vt(Vt, B, N, h, d), codes)
    assert not bool(Vt[:, :, N:].any()) and not bool(Vt[:, d:, :].any())
    assert torch.equal(Vt.float(), V0.float())      # same codes as the fp16-stored ones


@pytest.mark.parametrize("gran_name", ["GRAN_BLOCK", "GRAN_HEAD"])
@pytest.mark.parametrize("shape", [(1, 2048, 512, 4), (2, 640, 256, 2), (1, 4096, 512, 8), (1, 1000, 128, 4), (1, 2304, 128, 1), (1, 300, 96, 3)])
def test_profile_inputs_tolerance_and_two_level(qm, torch, oracle, shape, gran_name):
    B, N, dm, h = shape
    gran = getattr(qm, gran_name)
    name = "block" if gran == qm.GRAN_BLOCK else "head"
    q, k, v = (a.reshape(B, N, dm) for a in oracle.profile_inputs(B * N, dm))
    tq, tk, tv = _dev(torch, q, k, v)
    out = qm.forward(tq, tk, tv, h, kernel="int8_pv8", gran=gran)
    torch.cuda.synchronize()
    qm.binding.check_async_error()
    o = out.cpu().numpy()
    mx, rel = _err(o, oracle.mha(q, k, v, h, "f64"))
    assert mx <= INT8_MAX_ABS and rel <= INT8_REL_L2, (mx, rel)
    (qq, sq), (kq, sk), (vq, sv) = (oracle.quantize(x, h, name, 32) for x in (q, k, v))
    if gran == qm.GRAN_BLOCK:
        emu = oracle.mha_int8_emulated_block(qq, kq, vq, sq, sk, sv, h, 32, "exact")
    else:
        emu = oracle.mha_int8_emulated(qq, kq, vq, sq, sk, sv, h, "exact")
    _, rel_k = _err(o, emu)
    assert rel_k <= PV8_VS_EMU_REL_L2, rel_k
    # and it agrees with the default (fp16 P) mode to the same level
    ref16 = qm.forward(tq, tk, tv, h, kernel="int8", gran=gran)
    assert _err(o, ref16.cpu().numpy())[1] <= PV8_VS_EMU_REL_L2


@pytest.mark.parametrize("case", [(512, 256, 2), (1024, 128, 2), (640, 64, 2)])
def test_golden_inputs_and_rescale_path(qm, torch, oracle, case):
    """Signed inputs: the row max keeps growing, so the int32 O accumulators are rescaled (float round trip)."""
    N, dm, h = case
    q, k, v = oracle.golden_inputs(N, dm, h)
    for gran in (qm.GRAN_BLOCK, qm.GRAN_HEAD):
        out = qm.forward(*_dev(torch, q[None], k[None], v[None]), h, kernel="int8_pv8", gran=gran)
        torch.cuda.synchronize()
        qm.binding.check_async_error()
        mx, rel = _err(out[0].cpu().numpy(), oracle.mha(q, k, v, h, "f64"))
        assert mx <= INT8_MAX_ABS, (gran, mx, rel)


@pytest.mark.parametrize("N", [1, 2, 63, 64, 65, 127, 129, 191, 193, 255, 257, 449])
def test_every_pipeline_depth_and_ragged_tail(qm, torch, oracle, N):
    dm, h = 256, 2
    q, k, v = (a[None] for a in oracle.golden_inputs(N, dm, h))
    ref = oracle.mha(q, k, v, h, "f64")
    tq, tk, tv = _dev(torch, q, k, v)
    for gran in (qm.GRAN_BLOCK, qm.GRAN_HEAD):
        out = qm.forward(tq, tk, tv, h, kernel="int8_pv8", gran=gran)
        torch.cuda.synchronize()
        qm.binding.check_async_error()
        assert _err(out.cpu().numpy(), ref)[0] <= INT8_MAX_ABS, gran


def test_repeatable_and_headline_geometry(qm, torch, oracle):
    """C4 geometry (N=8192, d=128), 2 heads: sampled rows vs the float64 oracle; repeated launches bit-identical."""
    N, H, d = 8192, 2, 128
    q, k, v = oracle.profile_inputs(N, H * d)
    tq, tk, tv = _dev(torch, q, k, v)
    first = qm.forward(tq, tk, tv, H, kernel="int8_pv8", gran=qm.GRAN_BLOCK).clone()
    for _ in range(4):
        again = qm.forward(tq, tk, tv, H, kernel="int8_pv8", gran=qm.GRAN_BLOCK)
        torch.cuda.synchronize()
        assert torch.equal(first, again)
    qm.binding.check_async_error()
    rows = np.arange(0, N, 509)
    got = first.cpu().numpy()
    for hh in range(H):
        sl = slice(hh * d, (hh + 1) * d)
        ref = oracle.mha_head_rows(q[rows][:, sl], k[:, sl], v[:, sl], "f64")
        mx, rel = _err(got[rows][:, sl], ref)
        assert mx <= INT8_MAX_ABS and rel <= INT8_REL_L2, (mx, rel)
