"""Extended C-ABI (include/qmha.h: qmha_forward_ex and friends) and the CLI half of the boundary, on a B200
(run with -m gpu).  BF16 kernel, 16-bit inputs / outputs, per-call options, failure reporting, the
host-buffer path for the reference's B = 1 call shape, long sequences, two devices in one process and
bin/profile_* with its .cache files read back through the reference's own loaders (oracle/_ref).

Tolerances: INT8 max-abs <= 2e-2 / rel-L2 <= 1e-2 and FP16 max-abs <= 2e-3 as in BASELINE.json; the BF16
anchor is gated at max-abs <= 1e-2 against the float64 oracle (bf16 carries 8 significant bits: Q, K, P, V are
rounded to 2^-9 relative) and at 2e-3 against the oracle evaluated on the bf16-rounded inputs.
"""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INT8_MAX_ABS, INT8_REL_L2 = 2e-2, 1e-2
F16_MAX_ABS = 2e-3
BF16_MAX_ABS, BF16_VS_ROUNDED = 1e-2, 2e-3


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


def _sync(qm, torch):
    torch.cuda.synchronize()
    qm.binding.check_async_error()


def _unpack_rows(Qp, B, N, h, d):
    a = Qp.cpu().numpy().reshape(B, h, Qp.shape[1], Qp.shape[2])[:, :, :N, :d]
    return np.ascontiguousarray(a.transpose(0, 2, 1, 3)).reshape(B, N, h * d)


# ---------------------------------------------------------------------------------- BF16 anchor
@pytest.mark.parametrize("shape", [(1, 2048, 512, 4), (2, 640, 256, 2), (1, 1000, 128, 4), (1, 300, 96, 3), (1, 4096, 64, 2)])
def test_bf16_kernel_against_the_oracle(qm, torch, oracle, shape):
    B, N, dm, h = shape
    q, k, v = (a.reshape(B, N, dm) for a in oracle.profile_inputs(B * N, dm))
    tq, tk, tv = _dev(torch, q, k, v)
    out = qm.forward(tq, tk, tv, h, kernel="bf16")
    _sync(qm, torch)
    o = out.cpu().numpy()
    assert _err(o, oracle.mha(q, k, v, h, "f64"))[0] <= BF16_MAX_ABS
    rb = lambda a: torch.from_numpy(a).bfloat16().float().numpy()
    assert _err(o, oracle.mha(rb(q), rb(k), rb(v), h, "f64"))[0] <= BF16_VS_ROUNDED
    # golden (signed) inputs as well
    q, k, v = oracle.golden_inputs(N, dm, h)
    out = qm.forward(*_dev(torch, q, k, v), h, kernel="bf16")
    _sync(qm, torch)
    assert _err(out.cpu().numpy(), oracle.mha(q, k, v, h, "f64"))[0] <= BF16_MAX_ABS
    assert qm.kernel_id("fa_b200_bf16") == qm.KERNEL_BF16


def test_bf16_operands_are_round_to_nearest_bf16(qm, torch, oracle):
    q, k, v = oracle.golden_inputs(130, 192, 3)
    tq, tk, tv = _dev(torch, q[None], k[None], v[None])
    Qp, Kp, Vt = qm.convert_qkv_f16(tq, tk, tv, 3, kernel="bf16")
    assert Qp.dtype == torch.bfloat16
    assert np.array_equal(_unpack_rows(Qp.float(), 1, 130, 3, 64)[0], tq[0].bfloat16().float().cpu().numpy())
    vt = Vt.float().cpu().numpy().reshape(1, 3, Vt.shape[1], Vt.shape[2])[:, :, :64, :130]
    assert np.array_equal(np.ascontiguousarray(vt.transpose(0, 3, 1, 2)).reshape(130, 192), tv[0].bfloat16().float().cpu().numpy())


# ---------------------------------------------------------------------------------- 16-bit inputs / outputs
@pytest.mark.parametrize("shape", [(2, 300, 128, 2), (1, 130, 96, 4), (1, 50, 40, 2), (1, 1024, 512, 4), (3, 257, 128, 1)])
@pytest.mark.parametrize("dt", ["float16", "bfloat16"])
def test_16_bit_inputs_and_outputs(qm, torch, oracle, shape, dt):
    """fp16 / bf16 callers skip the fp32 round trip.  Exactness, not tolerance: (1) codes and scales from 16-bit
    inputs are bit-identical to the CPU restatement on the same (exactly representable) values; (2) the fp32
    result from 16-bit inputs equals the result from the widened fp32 inputs bit for bit; (3) a 16-bit output is
    the fp32 output rounded to nearest (TMA-store path for d % 32 == 0, plain stores otherwise); (4) nothing is
    written outside the output tensor."""
    B, N, dm, h = shape
    d = dm // h
    tdt = getattr(torch, dt)
    q, k, v = (a.reshape(B, N, dm) for a in oracle.profile_inputs(B * N, dm))
    t16 = [t.to(tdt) for t in _dev(torch, q - 0.5, k - 0.5, v)]
    t32 = [t.float() for t in t16]
    w = [t.cpu().numpy() for t in t32]
    for gran, name in ((qm.GRAN_BLOCK, "block"), (qm.GRAN_HEAD, "head"), (qm.GRAN_TENSOR, "tensor")):
        Qp, Kp, Vt, sc = qm.quantize_qkv(*t16, h, gran)
        codes, s = oracle.quantize(w[0], h, name, 32)
        assert np.array_equal(_unpack_rows(Qp, B, N, h, d), codes), (name, "Q codes")
        nb = -(-N // 32)
        got_s = sc[0].cpu().numpy()
        if gran == qm.GRAN_BLOCK:
            assert np.array_equal(got_s[:, :nb].reshape(-1), s)
        else:
            assert np.array_equal(got_s, s if gran == qm.GRAN_HEAD else np.full(B * h, s[0], np.float32))
    modes = [("int8", qm.GRAN_BLOCK), ("int8", qm.GRAN_HEAD), ("f16", qm.GRAN_HEAD), ("bf16", qm.GRAN_HEAD)]
    for kern, gran in modes:
        ref32 = qm.forward(*t32, h, kernel=kern, gran=gran)
        o32 = qm.forward(*t16, h, kernel=kern, gran=gran, out_dtype=torch.float32)
        _sync(qm, torch)
        assert torch.equal(o32, ref32), (kern, gran, "fp32 result differs between 16-bit and widened inputs")
        n = B * N * dm
        pad = 4096
        buf = torch.full((n + 2 * pad,), -7.0, dtype=tdt, device="cuda")
        o16 = buf[pad:pad + n].view(B, N, dm)
        qm.forward(*t16, h, kernel=kern, gran=gran, out=o16)
        _sync(qm, torch)
        assert torch.equal(o16, ref32.to(tdt)), (kern, gran, "16-bit output is not the rounded fp32 output")
        assert bool((buf[:pad] == -7.0).all()) and bool((buf[pad + n:] == -7.0).all()), (kern, gran)
    ref = oracle.mha(*w, h, "f64")
    assert _err(ref32.cpu().numpy(), ref)[0] <= BF16_MAX_ABS


def test_16_bit_inputs_with_odd_head_dimension_fail_loudly(qm, torch):
    t = torch.rand((1, 64, 30), device="cuda").half()   # d = 15
    with pytest.raises(qm.QmhaError):
        qm.forward(t, t, t, 2, kernel="f16")


# ---------------------------------------------------------------------------------- per-call options
def test_rope_and_kernel_are_per_call_arguments(qm, torch, oracle):
    """qmha_args carries kernel, granularity and RoPE per call; nothing process-wide is touched."""
    N, dm, h = 300, 256, 2
    q, k, v = oracle.golden_inputs(N, dm, h, rope=False)
    tq, tk, tv = _dev(torch, q[None], k[None], v[None])
    assert not qm.get_rope()
    per_call = qm.forward(tq, tk, tv, h, kernel="int8", gran=qm.GRAN_BLOCK, rope=True)
    assert not qm.get_rope()
    qm.set_rope(True)
    try:
        glob = qm.forward(tq, tk, tv, h, kernel="int8", gran=qm.GRAN_BLOCK)
        off = qm.forward(tq, tk, tv, h, kernel="int8", gran=qm.GRAN_BLOCK, rope=False)
    finally:
        qm.set_rope(False)
    plain = qm.forward(tq, tk, tv, h, kernel="int8", gran=qm.GRAN_BLOCK)
    _sync(qm, torch)
    assert torch.equal(per_call, glob) and torch.equal(off, plain) and not torch.equal(per_call, plain)
    ref = oracle.cpu_reference_rope(q, k, v, h)
    mx, rel = _err(per_call[0].cpu().numpy(), ref)
    assert mx <= INT8_MAX_ABS and rel <= INT8_REL_L2
    other_base = qm.forward(tq, tk, tv, h, kernel="f16", rope=True, rope_base=500.0)
    _sync(qm, torch)
    assert not torch.equal(other_base, qm.forward(tq, tk, tv, h, kernel="f16", rope=True))
    # gran = -1: the granularity solve() would use
    assert torch.equal(qm.forward(tq, tk, tv, h, kernel="int8", gran=-1), plain)


@pytest.mark.parametrize("shape", [(1, 300, 256, 2), (2, 128, 128, 4), (1, 1000, 192, 3)])
def test_fused_rope_with_per_tensor_scales(qm, torch, oracle, shape):
    """The two-pass (global absmax) quantiser rotates before it takes the maxima too: codes and the single
    scale per tensor bit-exact vs "RoPE on the host, then quantise" with the oracle."""
    B, N, dm, h = shape
    d = dm // h
    q, k, v = (np.stack([a] * B) for a in oracle.golden_inputs(N, dm, h, rope=False))
    if B > 1:
        k[1] *= 2.5
    qr = np.stack([oracle.apply_rope(x.copy(), h) for x in q])
    kr = np.stack([oracle.apply_rope(x.copy(), h) for x in k])
    tq, tk, tv = _dev(torch, q, k, v)
    Qp, Kp, Vt, sc = qm.quantize_qkv(tq, tk, tv, h, qm.GRAN_TENSOR, rope=True)
    for i, (x, packed) in enumerate(((qr, Qp), (kr, Kp))):
        codes, s = oracle.quantize(x, h, "tensor")
        assert np.array_equal(_unpack_rows(packed, B, N, h, d), codes), "QK"[i]
        assert np.array_equal(sc[i].cpu().numpy(), np.full(B * h, s[0], np.float32)), "QK"[i]
    out = qm.forward(tq, tk, tv, h, kernel="int8", gran=qm.GRAN_TENSOR, rope=True)
    _sync(qm, torch)
    ref = np.stack([oracle.cpu_reference_rope(q[b], k[b], v[b], h) for b in range(B)])
    assert _err(out.cpu().numpy(), ref)[0] <= INT8_MAX_ABS


# ---------------------------------------------------------------------------------- failure reporting
def test_a_stalled_launch_is_reported_by_the_next_call_and_poisons_nothing(qm, torch, oracle):
    """A bounded mbarrier wait that gives up records (launch id, wait site).  The record must surface even when the
    caller never polls qmha_check_async_error() — the NEXT entry into the library fails with it — and later
    launches must run normally (an earlier version made every later launch return at kernel entry)."""
    q, k, v = oracle.profile_inputs(512, 256)
    tq, tk, tv = _dev(torch, q, k, v)
    good = qm.forward(tq, tk, tv, 2, kernel="int8", gran=qm.GRAN_BLOCK).clone()
    _sync(qm, torch)
    assert qm.lib().qmha_debug_inject_stall(204) == 0
    with pytest.raises(qm.QmhaError, match="wait site 204"):
        qm.forward(tq, tk, tv, 2, kernel="int8", gran=qm.GRAN_BLOCK)
    again = qm.forward(tq, tk, tv, 2, kernel="int8", gran=qm.GRAN_BLOCK)   # reported once, then business as usual
    _sync(qm, torch)
    assert torch.equal(again, good)
    # the polling interface sees a record too, and clears both copies of it
    assert qm.lib().qmha_debug_inject_stall(301) == 0
    torch.cuda.synchronize()
    with pytest.raises(qm.QmhaError, match="wait site 301"):
        qm.binding.check_async_error()
    qm.binding.check_async_error()
    assert torch.equal(qm.forward(tq, tk, tv, 2, kernel="int8", gran=qm.GRAN_BLOCK), good)
    # ... and so do the synchronous entries
    assert qm.lib().qmha_debug_inject_stall(102) == 0
    with pytest.raises(qm.QmhaError, match="wait site 102"):
        qm.solve(tq, tk, tv, 512, 256, 2)
    assert torch.equal(qm.solve(tq, tk, tv, 512, 256, 2), good)


# ---------------------------------------------------------------------------------- host buffers, B = 1
@pytest.mark.parametrize("kern,gran_name", [("int8", "GRAN_BLOCK"), ("int8", "GRAN_HEAD"), ("f16", "GRAN_HEAD"), ("bf16", "GRAN_HEAD")])
def test_host_buffer_entry_pipelines_over_head_groups(qm, torch, oracle, kern, gran_name, monkeypatch):
    """The reference's call shape is B = 1 (include/launchers.h:41-62): the host-buffer path cuts the heads into
    groups (strided 2-D copies) so that copies and kernels still overlap.  Whatever the grouping, the result is
    bit-identical to the device path (scales are per head or finer)."""
    gran = getattr(qm, gran_name)
    N, dm, h = 1000, 512, 8
    q, k, v = oracle.profile_inputs(N, dm)
    hq, hk, hv = (torch.from_numpy(a).pin_memory() for a in (q, k, v))
    dev_out = qm.forward(*_dev(torch, q, k, v), h, kernel=kern, gran=gran)
    _sync(qm, torch)
    for group in ("", "1", "2", "4", "8"):
        if group:
            monkeypatch.setenv("QMHA_HOST_HEAD_GROUP", group)
        else:
            monkeypatch.delenv("QMHA_HOST_HEAD_GROUP", raising=False)
        ho = qm.forward_host(hq, hk, hv, h, kernel=kern, gran=gran)
        assert np.array_equal(ho.numpy(), dev_out.cpu().numpy()), (kern, gran_name, group)
    monkeypatch.delenv("QMHA_HOST_HEAD_GROUP", raising=False)
    # 16-bit host buffers in and out (half the PCIe bytes): identical to the device path on the same 16-bit tensors
    for tdt in (torch.float16, torch.bfloat16):
        h16 = [t.to(tdt).pin_memory() for t in (hq, hk, hv)]
        want16 = qm.forward(*(t.cuda() for t in h16), h, kernel=kern, gran=gran)
        _sync(qm, torch)
        got16 = qm.forward_host(*h16, h, kernel=kern, gran=gran)
        assert got16.dtype == tdt and torch.equal(got16, want16.cpu()), (kern, gran_name, tdt)
        got32 = qm.forward_host(*h16, h, kernel=kern, gran=gran, out=torch.empty((N, dm), dtype=torch.float32).pin_memory())
        want32 = qm.forward(*(t.cuda() for t in h16), h, kernel=kern, gran=gran, out_dtype=torch.float32)
        _sync(qm, torch)
        assert torch.equal(got32, want32.cpu()), (kern, gran_name, tdt, "fp32 out")
    # batched, pageable memory, gran -1
    q3, k3, v3 = (np.stack([a, a * 0.5, a * 2.0]) for a in (q[:300], k[:300], v[:300]))
    ho = qm.forward_host(*(torch.from_numpy(a) for a in (q3, k3, v3)), h, kernel=kern, gran=-1)
    want = qm.forward(*_dev(torch, q3, k3, v3), h, kernel=kern, gran=-1)
    _sync(qm, torch)
    assert np.array_equal(ho.numpy(), want.cpu().numpy())


# ---------------------------------------------------------------------------------- long sequences
def test_solve_falls_back_to_per_head_scales_when_the_block_table_no_longer_fits(qm, torch, oracle):
    L = qm.lib()
    assert L.qmha_granularity_for(8192, 4096, 32) == qm.GRAN_BLOCK
    assert L.qmha_granularity_for(49152, 128, 1) == qm.GRAN_BLOCK
    assert L.qmha_granularity_for(70000, 128, 1) == qm.GRAN_HEAD
    assert L.qmha_granularity_for(300000, 64, 2) == qm.GRAN_BLOCK     # d = 32: small tiles leave room for a longer table
    assert L.qmha_granularity_for(1000000, 64, 2) == qm.GRAN_HEAD
    N, d = 70000, 128
    gen = torch.Generator(device="cuda").manual_seed(3)
    tq, tk, tv = (torch.rand((N, d), device="cuda", generator=gen) for _ in range(3))
    with pytest.raises(qm.QmhaError, match="too long"):
        qm.forward(tq, tk, tv, 1, kernel="int8", gran=qm.GRAN_BLOCK)
    out = qm.solve(tq, tk, tv, N, d, 1)          # the reference's entry point degrades instead of failing
    rows = np.array([0, 1, 31999, 69998, 69999])
    ref = oracle.mha_head_rows(tq.cpu().numpy()[rows], tk.cpu().numpy(), tv.cpu().numpy(), "f64")
    mx, rel = _err(out.cpu().numpy()[rows], ref)
    assert mx <= INT8_MAX_ABS and rel <= INT8_REL_L2, (mx, rel)


# ---------------------------------------------------------------------------------- two devices, one process
def test_two_devices_in_one_process(qm, torch, oracle):
    """One workspace per device (api.cu: std::map<int, Workspace>): interleaved asynchronous calls on two GPUs
    from one process, then a host-buffer call on each."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run with gpurun --gpus 2)")
    q, k, v = (a.reshape(2, 640, 256) for a in oracle.profile_inputs(2 * 640, 256))
    ref = oracle.mha(q, k, v, 2, "f64")
    outs = {}
    for rep in range(3):
        for dev in (0, 1):
            with torch.cuda.device(dev):
                t = [torch.from_numpy(a).to(f"cuda:{dev}") * (1.0 + dev) for a in (q, k, v)]
                outs[dev] = qm.forward(t[0] / (1.0 + dev), t[1] / (1.0 + dev), t[2] / (1.0 + dev), 2, kernel="int8", gran=qm.GRAN_BLOCK)
    for dev in (0, 1):
        with torch.cuda.device(dev):
            torch.cuda.synchronize()
            qm.binding.check_async_error()
            mx, rel = _err(outs[dev].cpu().numpy(), ref)
            assert mx <= INT8_MAX_ABS and rel <= INT8_REL_L2, (dev, mx, rel)
            ho = qm.forward_host(*(torch.from_numpy(a) for a in (q, k, v)), 2, kernel="f16")
            assert _err(ho.numpy(), ref)[0] <= F16_MAX_ABS
    assert torch.equal(outs[0].cpu(), outs[1].cpu())


def test_plain_c_consumer_of_the_abi(torch, tmp_path):
    """tests/c/abi_consumer.c: solve() and qmha_forward_ex() (FP16 kernel, strided output slab) called from C with
    memory from the CUDA runtime's C API — no Python, no torch on the path under test."""
    import quantizedmha_b200 as qm
    from test_host_cpu import _build_c_consumer
    exe = _build_c_consumer(qm, tmp_path)
    r = subprocess.run([exe, "run"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "abi_consumer ok" in r.stdout, r.stdout + r.stderr


# ---------------------------------------------------------------------------------- the CLI half of the boundary
def _profile_binary(kernel):
    path = os.path.join(ROOT, "bin", f"profile_{kernel}")
    if not os.path.exists(path):
        subprocess.run(["make", "-C", ROOT, "-s", "driver", f"KERNEL={kernel}"], check=True)
    return path


@pytest.mark.parametrize("kernel", ["fa_tc_int8_b", "fa_tc_v2a"])
def test_profile_binary_flags_exit_codes_and_cache_files(torch, oracle, reflib, tmp_path, kernel):
    """bin/profile_<KERNEL> (drivers/main.cu of the reference: flags :45-58, exit 1 on a failed check :97-99) and the
    .cache files it writes, read back through the REFERENCE's own loaders (inputs/data.cu:84-109 load_inputs,
    utils/verify.cu:128-151 load_reference, compiled where they lie as oracle/_ref)."""
    exe = _profile_binary(kernel)
    N, dm, h = 512, 256, 2
    args = [exe, f"--N={N}", f"--d_model={dm}", f"--h={h}", "--warmup=1", "--runs=2", "--json"]
    r = subprocess.run(args, cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "Correctness check PASSED." in r.stdout and "Profiling complete." in r.stdout
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1]
    rec = json.loads(line)
    assert rec["kernel"] == kernel and rec["N"] == N and rec["d_model"] == dm and rec["h"] == h and rec["ms_median"] > 0
    inp = tmp_path / ".cache" / f"input_random_N{N}_d{dm}.bin"
    refc = tmp_path / ".cache" / f"ref_N{N}_d{dm}.bin"
    assert inp.exists() and refc.exists()
    assert inp.stat().st_size == 8 + 3 * N * dm * 4 and refc.stat().st_size == 8 + N * dm * 4
    q, k, v = oracle.profile_inputs(N, dm)
    if reflib is not None:
        got = reflib.load_inputs(inp, N, dm)
        assert got is not None, "the reference's load_inputs rejected the driver's input cache"
        assert all(np.array_equal(a, b) for a, b in zip(got, (q, k, v)))
        assert reflib.load_inputs(inp, N + 1, dm) is None          # header is checked like the reference does
        ref = reflib.load_reference(refc, N, dm)
        assert ref is not None and np.array_equal(ref, np.ones((N, dm), np.float32))
        # and the other direction: files written by the reference's writers are accepted by the driver
        (tmp_path / "w").mkdir()
        (tmp_path / "w" / ".cache").mkdir()
        assert reflib.save_inputs(q, k, v, tmp_path / "w" / ".cache" / f"input_random_N{N}_d{dm}.bin")
        assert reflib.save_reference(np.ones((N, dm), np.float32), tmp_path / "w" / ".cache" / f"ref_N{N}_d{dm}.bin")
        r2 = subprocess.run(args, cwd=tmp_path / "w", capture_output=True, text=True, timeout=300)
        assert r2.returncode == 0 and "Loaded input matrices from" in r2.stdout and "Loaded CPU reference from" in r2.stdout
    # second run: the caches are found; --no-check skips the check like the reference
    r = subprocess.run(args + ["--no-check"], cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "Loaded input matrices from" in r.stdout and "Skipping correctness check" in r.stdout
    # a failed check exits with 1 (drivers/main.cu:97-99)
    r = subprocess.run(args, cwd=tmp_path, capture_output=True, text=True, timeout=300, env=dict(os.environ, QMHA_DRIVER_CORRUPT="1"))
    assert r.returncode == 1 and "Correctness check FAILED" in r.stderr
    # unknown kernel name: usage error, not a crash
    r = subprocess.run([exe, "--kernel=nope", "--runs=0"], cwd=tmp_path, capture_output=True, text=True, timeout=60)
    assert r.returncode == 2
    # -k <name> form, batch flag, RoPE flag
    r = subprocess.run([exe, "-k", kernel, f"--N={N}", f"--d_model={dm}", f"--h={h}", "--B=2", "--runs=1", "--rope", "--json"],
                       cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr


# ---------------------------------------------------------------------------------- per-head quantiser variants
@pytest.mark.parametrize("shape", [(1, 50, 64, 8), (2, 300, 256, 2), (1, 1024, 512, 4), (3, 129, 96, 3), (2, 4096, 256, 2)])
def test_persistent_per_head_quantiser_is_bit_identical(qm, torch, oracle, shape, monkeypatch):
    """QMHA_STREAM_QUANT=1 selects the persistent-grid per-head quantiser (queue of absmax / quantise items, no
    clusters): codes, scales and padding identical to the cluster kernel and to the CPU restatement, with and
    without fused RoPE, for fp32 and 16-bit inputs."""
    B, N, dm, h = shape
    d = dm // h
    q, k, v = (np.stack(x) for x in zip(*[oracle.golden_inputs(N, dm, h, rope=False) for _ in range(B)]))
    q[1:] *= 1.7
    tq, tk, tv = _dev(torch, q, k, v)
    for rope in ((False, True) if d % 8 == 0 else (False,)):
        monkeypatch.delenv("QMHA_STREAM_QUANT", raising=False)
        ref = qm.quantize_qkv(tq, tk, tv, h, qm.GRAN_HEAD, rope=rope)
        monkeypatch.setenv("QMHA_STREAM_QUANT", "1")
        got = qm.quantize_qkv(tq, tk, tv, h, qm.GRAN_HEAD, rope=rope)
        torch.cuda.synchronize()
        assert all(torch.equal(a, b) for a, b in zip(ref, got)), rope
        got16 = qm.quantize_qkv(tq.half(), tk.half(), tv.half(), h, qm.GRAN_HEAD, rope=rope)
        monkeypatch.delenv("QMHA_STREAM_QUANT", raising=False)
        ref16 = qm.quantize_qkv(tq.half(), tk.half(), tv.half(), h, qm.GRAN_HEAD, rope=rope)
        torch.cuda.synchronize()
        assert all(torch.equal(a, b) for a, b in zip(ref16, got16)), rope
    codes, s = oracle.quantize(k, h, "head")
    monkeypatch.setenv("QMHA_STREAM_QUANT", "1")
    Qp, Kp, Vt, sc = qm.quantize_qkv(tq, tk, tv, h, qm.GRAN_HEAD, rope=False)
    monkeypatch.delenv("QMHA_STREAM_QUANT", raising=False)
    assert np.array_equal(_unpack_rows(Kp, B, N, h, d), codes) and np.array_equal(sc[1].cpu().numpy(), s)
    assert (Qp[:, N:, :] == 0).all() and (Vt[:, :, N:] == 0).all()


@pytest.mark.gpu
def test_development_cycle_counters_account_for_every_cta(tmp_path):
    """qmha_debug_cycles / qmha_debug_sm_spans (QMHA_CYCLES=1, set before the first call, hence a fresh process; counted for
    qmha_attention_prepared launches): one launch
    of 2 x 4 units x 2 query blocks reports 16 CTAs, a positive residency sum, and per-SM spans that cover it."""
    script = tmp_path / "cycles.py"
    script.write_text(
        "import ctypes as C, sys\n"
        f"sys.path.insert(0, {ROOT!r})\n"
        "import torch, quantizedmha_b200 as qm\n"
        "L = qm.lib()\n"
        "L.qmha_debug_sm_spans.argtypes = [C.POINTER(C.c_ulonglong), C.c_int, C.c_int]\n"
        "q, k, v = (torch.rand((2, 512, 512), device='cuda') for _ in range(3))\n"
        "Qp, Kp, Vt, sc = qm.quantize_qkv(q, k, v, 4, qm.GRAN_BLOCK); out = torch.empty_like(q)\n"
        "P = lambda t: C.c_void_p(t.data_ptr())\n"
        "run = lambda: L.qmha_attention_prepared(P(Qp), P(Kp), P(Vt), P(sc), P(out), 2, 512, 512, 4, qm.binding.KERNEL_INT8, qm.GRAN_BLOCK, None)\n"
        "assert run() == 0; torch.cuda.synchronize()\n"
        "c = (C.c_ulonglong * 2)(); assert L.qmha_debug_cycles(c, 1) == 0\n"
        "assert run() == 0; torch.cuda.synchronize()\n"
        "assert L.qmha_debug_cycles(c, 0) == 0\n"
        "sp = (C.c_ulonglong * 384)(); assert L.qmha_debug_sm_spans(sp, 192, 1) == 0\n"
        "spans = [sp[2 * i] for i in range(192) if sp[2 * i]]\n"
        "print(c[1], c[0], len(spans), sum(spans))\n"
        "assert c[1] == 16 and c[0] > 0 and 1 <= len(spans) <= 16 and sum(spans) >= c[0] * 0.99\n"
        "assert L.qmha_debug_sm_spans(sp, 193, 0) != 0\n")
    r = subprocess.run([sys.executable, str(script)], env=dict(os.environ, QMHA_CYCLES="1"), capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
