"""Output placement of the attention epilogue (qmha_args.o_row_stride / o_batch_stride / peer_O; SURVEY §8f row 4,
the reference's concat_mat step, include/launchers.h:59-61 + utils/utils.cu:15-22, folded into the kernel): the
result may be written straight into a (batch, head-range) slab of a larger tensor and replicated into further
destinations.  One GPU is enough here — the replicas are extra buffers on the same device; the NVLink / CUDA-IPC
leg runs under torchrun on a multi-GPU box (tools/nccl_gather_check.py) and, with two visible GPUs, in the last test.

Bar: every placement is BIT-IDENTICAL to the dense single-destination call on the same inputs, and nothing
outside the slab is touched."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch():
    import torch as t
    if not t.cuda.is_available():
        pytest.skip("no CUDA device")
    return t


@pytest.fixture(scope="module")
def qm(torch):
    import quantizedmha_b200 as q
    assert os.path.exists(q.lib_path()), "libqmha.so missing: the GPU tests never fall back"
    return q


def _inputs(torch, B, N, dm, seed, dtype=None):
    g = torch.Generator(device="cpu").manual_seed(seed)
    t = [(torch.rand((B, N, dm), generator=g) - 0.3).cuda() for _ in range(3)]
    return [x.to(dtype) if dtype is not None else x for x in t]


def _sync(torch, qm):
    torch.cuda.synchronize()
    qm.binding.check_async_error()


# (kernel, d, N): every epilogue of the attention kernel — TMA tensor stores (d % 32 == 0 with a 64 KB K ring),
# rows staged through the dead Q tile (d = 120 padded to 128), plain per-row stores (small head dimensions)
EPILOGUES = [("int8", 128, 600), ("f16", 128, 300), ("bf16", 64, 257), ("int8", 120, 300), ("int8", 64, 333),
             ("int8", 32, 130), ("f16", 32, 200), ("int8_pv8", 128, 520)]


@pytest.mark.parametrize("kernel,d,N", EPILOGUES)
@pytest.mark.parametrize("out_dtype", ["float32", "float16"])
def test_slab_of_a_larger_tensor_and_replicas_are_bit_identical(torch, qm, kernel, d, N, out_dtype):
    H_all, h0, h1, B_all, b = 5, 1, 4, 3, 1
    heads = h1 - h0
    odt = getattr(torch, out_dtype)
    q, k, v = _inputs(torch, 1, N, heads * d, seed=d + N)
    gran = qm.GRAN_BLOCK if (kernel.startswith("int8") and d % 4 == 0) else qm.GRAN_HEAD
    dense = qm.forward(q, k, v, heads, kernel=kernel, gran=gran, out_dtype=odt)
    _sync(torch, qm)
    big = [torch.full((B_all, N, H_all * d), -7.0, dtype=odt, device="cuda") for _ in range(3)]
    views = [t[b, :, h0 * d:h1 * d] for t in big]
    qm.forward(q[0], k[0], v[0], heads, kernel=kernel, gran=gran, out=views[0], peer_outs=views[1:])
    _sync(torch, qm)
    for t in big:
        assert torch.equal(t[b, :, h0 * d:h1 * d], dense[0])
        untouched = t.clone()
        untouched[b, :, h0 * d:h1 * d] = -7.0
        assert bool((untouched == -7.0).all()), "the epilogue wrote outside its slab"


@pytest.mark.parametrize("kernel,d", [("int8", 128), ("int8", 64), ("f16", 128)])
def test_batched_slab_with_batch_stride_and_raw_peer_addresses(torch, qm, kernel, d):
    B, N, heads, H_all = 3, 384, 2, 4
    q, k, v = _inputs(torch, B, N, heads * d, seed=7)
    dense = qm.forward(q, k, v, heads, kernel=kernel, gran=-1)
    _sync(torch, qm)
    big = torch.zeros((B + 2, N + 5, H_all * d), device="cuda")
    peers = [torch.zeros_like(big) for _ in range(qm.binding.MAX_PEERS)]
    sl = (slice(1, 1 + B), slice(0, N), slice(2 * d, 4 * d))
    off = big[sl].data_ptr() - big.data_ptr()
    qm.forward(q, k, v, heads, kernel=kernel, gran=-1, out=big[sl], peer_outs=[p.data_ptr() + off for p in peers])
    _sync(torch, qm)
    for t in [big] + peers:
        assert torch.equal(t[sl], dense)
        z = t.clone()
        z[sl] = 0
        assert not bool(z.any())


def test_placement_errors_are_loud(torch, qm):
    N, d, heads = 128, 32, 2
    q, k, v = _inputs(torch, 1, N, heads * d, seed=3)
    big = torch.zeros((N, heads * d + 2), device="cuda")           # row pitch 66 floats = 264 B: not a 16-byte multiple
    with pytest.raises(qm.QmhaError, match="16 bytes"):
        qm.forward(q[0], k[0], v[0], heads, out=big[:, :heads * d])
    with pytest.raises(qm.QmhaError, match="at most"):
        qm.forward(q[0], k[0], v[0], heads, peer_outs=[torch.zeros((N, heads * d), device="cuda")] * 8)
    with pytest.raises(qm.QmhaError, match="strides of out"):
        qm.forward(q[0], k[0], v[0], heads, peer_outs=[torch.zeros((N, 2 * heads * d), device="cuda")[:, :heads * d]])
    with pytest.raises(qm.QmhaError, match="contiguous"):
        qm.forward(q[0], k[0], v[0], heads, out=torch.zeros((heads * d, N), device="cuda").t())
    # and the library still works afterwards
    out = qm.forward(q[0], k[0], v[0], heads)
    _sync(torch, qm)
    assert bool(torch.isfinite(out).all())


def test_fused_gather_plan_on_one_device_matches_the_full_forward(torch, qm):
    """sharding.forward_fused_gather with W simulated ranks on ONE device: each "rank" runs its launch plan with the
    other ranks' replicas as peer destinations; afterwards every replica equals the unsharded forward bit for bit."""
    from quantizedmha_b200 import sharding as sh
    B, N, H, d, W = 2, 300, 6, 64, 4
    Q, K, V = _inputs(torch, B, N, H * d, seed=11)
    full = qm.forward(Q, K, V, H, kernel="int8", gran=qm.GRAN_BLOCK)
    _sync(torch, qm)
    reps = [torch.full_like(full, float("nan")) for _ in range(W)]

    class Rep:   # what ReplicatedOutput provides, without a process group
        def __init__(self, r):
            self.world, self.rank, self.local = W, r, reps[r]
            self.peer_base = {j: reps[j].data_ptr() for j in range(W) if j != r}

    for r in range(W):
        sh.forward_fused_gather(Q, K, V, H, Rep(r), kernel="int8", gran=qm.GRAN_BLOCK, fence=False)
    _sync(torch, qm)
    for t in reps:
        assert torch.equal(t, full)


def test_ipc_export_of_a_torch_tensor(torch, qm):
    """A handle names the allocation that holds the tensor; the offset finds the tensor inside it.  (CUDA refuses to
    open a handle in the process that exported it, so the round trip itself is covered by the 2-GPU run.)"""
    a = torch.zeros(1000, device="cuda")
    t = torch.zeros((64, 256), device="cuda")
    h, off = qm.binding.ipc_export(t)
    assert len(h) == 64 and off >= 0 and off % 16 == 0
    h2, off2 = qm.binding.ipc_export(t[8:])
    assert off2 == off + 8 * 256 * 4
    try:                                   # refused in the exporting process; whatever happens, the library stays usable
        qm.binding.ipc_open(h, off)
    except qm.QmhaError:
        pass
    qm.binding.ipc_close_all()
    out = qm.forward(t, t, t, 2)
    torch.cuda.synchronize()
    qm.binding.check_async_error()
    assert bool(torch.isfinite(out).all())
    del a


def test_device_ordinal_in_the_argument_block(torch, qm):
    """qmha_args.device: the call runs on the device that owns the tensors whatever the caller's current device is,
    and leaves the current device alone; a wrong ordinal fails loudly."""
    import ctypes as C
    q, k, v = _inputs(torch, 1, 300, 128, seed=41)
    ref = qm.forward(q, k, v, 2)
    _sync(torch, qm)
    a = qm.QmhaArgs()
    qm.lib().qmha_args_init(C.byref(a))
    assert a.device == -1
    o = torch.empty_like(q)
    a.Q, a.K, a.V, a.O = q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr()
    a.B, a.N, a.d_model, a.h, a.gran = 1, 300, 128, 2, qm.GRAN_HEAD
    a.device = torch.cuda.device_count()
    assert qm.lib().qmha_forward_ex(C.byref(a)) != 0 and b"no such device" in qm.lib().qmha_last_error()
    a.device = 0
    assert qm.lib().qmha_forward_ex(C.byref(a)) == 0
    _sync(torch, qm)
    assert torch.equal(o, ref)
    if torch.cuda.device_count() >= 2:                    # tensors on GPU 1 while GPU 0 is current
        q1, k1, v1 = (t.to("cuda:1") for t in (q, k, v))
        assert torch.cuda.current_device() == 0
        o1 = qm.forward(q1, k1, v1, 2)
        torch.cuda.synchronize(1)
        assert torch.cuda.current_device() == 0 and torch.equal(o1.cpu(), ref.cpu())


def test_replica_on_a_second_gpu_in_one_process(torch, qm):
    """The NVLink leg without process boundaries: the kernel runs on GPU 0 and its epilogue also stores into a replica
    that lives on GPU 1 (peer access enabled through the library)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run with gpurun --gpus 2)")
    qm.binding.enable_peer_access(0, 1)
    B, N, heads, d = 2, 1000, 3, 128
    q, k, v = _inputs(torch, B, N, heads * d, seed=5)
    for kernel, odt in (("int8", torch.float32), ("bf16", torch.bfloat16), ("int8", torch.float16)):
        remote = torch.full((B, N, 2 * heads * d), -1.0, dtype=odt, device="cuda:1")
        local = torch.empty((B, N, heads * d), dtype=odt, device="cuda:0")
        torch.cuda.synchronize(1)
        # (local is dense, the remote slab is strided: same shape, so give the kernel one pitch for both)
        local_wide = torch.empty((B, N, 2 * heads * d), dtype=odt, device="cuda:0")
        qm.forward(q, k, v, heads, kernel=kernel, gran=-1, out=local_wide[:, :, :heads * d],
                   peer_outs=[remote[:, :, :heads * d]])
        _sync(torch, qm)
        dense = qm.forward(q, k, v, heads, kernel=kernel, gran=-1, out=local)
        _sync(torch, qm)
        assert torch.equal(local_wide[:, :, :heads * d], dense)
        assert torch.equal(remote[:, :, :heads * d].cpu(), dense.cpu())
        assert bool((remote[:, :, heads * d:] == -1.0).all())


@pytest.mark.parametrize("gran", ["block", "head"])
def test_persistent_kernel_is_bit_identical(torch, qm, gran, monkeypatch):
    """QMHA_PERSIST=1 (opt-in): one CTA per SM walks the (unit, query block) items with the barriers, TMEM and
    constant tiles kept alive, the next item's Q / K tiles prefetched under the previous item's output stores.
    256 items on 148 SMs, so CTAs process one or two items; also with a strided slab + replica as output."""
    B, N, H, d = 2, 4096, 8, 128
    q, k, v = _inputs(torch, B, N, H * d, seed=21)
    g = qm.GRAN_BLOCK if gran == "block" else qm.GRAN_HEAD
    monkeypatch.setenv("QMHA_PERSIST", "0")
    ref = qm.forward(q, k, v, H, kernel="int8", gran=g)
    _sync(torch, qm)
    monkeypatch.setenv("QMHA_PERSIST", "1")
    launches = qm.launch_count()
    out = qm.forward(q, k, v, H, kernel="int8", gran=g)
    _sync(torch, qm)
    assert qm.launch_count() > launches
    assert torch.equal(out, ref)
    big = torch.zeros((B, N + 3, 2 * H * d), device="cuda")
    peer = torch.zeros_like(big)
    qm.forward(q, k, v, H, kernel="int8", gran=g, out=big[:, :N, H * d:], peer_outs=[peer[:, :N, H * d:]])
    _sync(torch, qm)
    assert torch.equal(big[:, :N, H * d:], ref) and torch.equal(peer, big)
    for grid in ("1", "37"):                      # many items per CTA
        monkeypatch.setenv("QMHA_PERSIST_GRID", grid)
        small = qm.forward(q[:1, :1024], k[:1, :1024], v[:1, :1024], H, kernel="int8", gran=g)
        _sync(torch, qm)
        monkeypatch.setenv("QMHA_PERSIST", "0")
        small_ref = qm.forward(q[:1, :1024], k[:1, :1024], v[:1, :1024], H, kernel="int8", gran=g)
        _sync(torch, qm)
        monkeypatch.setenv("QMHA_PERSIST", "1")
        assert torch.equal(small, small_ref)


@pytest.mark.parametrize("kernel,gran,dtype", [("int8", "block", "float32"), ("int8", "head", "float32"),
                                                ("int8", "tensor", "float32"), ("f16", "head", "float32"),
                                                ("bf16", "head", "bfloat16"), ("int8", "block", "float16"),
                                                ("int8_pv8", "block", "float32")])
def test_strided_input_slabs_are_read_in_place(torch, qm, kernel, gran, dtype):
    """Q, K, V as (batch, head-range) views of larger tensors (qmha_args.in_row_stride / in_batch_stride): every
    quantise / convert kernel reads them in place; results equal the call on contiguous copies bit for bit."""
    B_all, N, H_all, d, h0, h1 = 3, 333, 5, 64, 1, 4
    heads = h1 - h0
    dt = getattr(torch, dtype)
    g = {"block": qm.GRAN_BLOCK, "head": qm.GRAN_HEAD, "tensor": qm.GRAN_TENSOR}[gran]
    full = _inputs(torch, B_all, N + 2, H_all * d, seed=31, dtype=dt)
    views = [t[1:3, :N, h0 * d:h1 * d] for t in full]
    assert not views[0].is_contiguous()
    ref = qm.forward(*(v.contiguous() for v in views), heads, kernel=kernel, gran=g)
    _sync(torch, qm)
    out = qm.forward(*views, heads, kernel=kernel, gran=g)
    _sync(torch, qm)
    assert torch.equal(out, ref)
    # one batch entry, rope on (the rotation uses the row index inside the slab), 2-D views
    ref2 = qm.forward(*(v[0].contiguous() for v in views), heads, kernel=kernel, gran=g, rope=True)
    out2 = qm.forward(*(v[0] for v in views), heads, kernel=kernel, gran=g, rope=True)
    _sync(torch, qm)
    assert torch.equal(out2, ref2)


def test_strided_inputs_with_the_persistent_per_head_quantiser_and_bad_pitches(torch, qm, monkeypatch):
    import ctypes as C
    B, N, heads, d = 2, 512, 2, 128
    full = _inputs(torch, B, N, 3 * heads * d, seed=33)
    views = [t[:, :, heads * d:2 * heads * d] for t in full]
    ref = qm.forward(*(v.contiguous() for v in views), heads, kernel="int8", gran=qm.GRAN_HEAD)
    monkeypatch.setenv("QMHA_STREAM_QUANT", "1")
    out = qm.forward(*views, heads, kernel="int8", gran=qm.GRAN_HEAD)
    _sync(torch, qm)
    assert torch.equal(out, ref)
    monkeypatch.delenv("QMHA_STREAM_QUANT")
    a = qm.QmhaArgs()
    qm.lib().qmha_args_init(C.byref(a))
    o = torch.empty((B, N, heads * d), device="cuda")
    a.Q, a.K, a.V, a.O = (views[0].data_ptr(), views[1].data_ptr(), views[2].data_ptr(), o.data_ptr())
    a.B, a.N, a.d_model, a.h = B, N, heads * d, heads
    a.in_row_stride = heads * d - 4                      # smaller than a row
    assert qm.lib().qmha_forward_ex(C.byref(a)) != 0 and b"input strides" in qm.lib().qmha_last_error()
    a.in_row_stride = 3 * heads * d + 1                  # not a multiple of 16 bytes
    assert qm.lib().qmha_forward_ex(C.byref(a)) != 0 and b"16 bytes" in qm.lib().qmha_last_error()
    a.in_row_stride, a.in_batch_stride = 3 * heads * d, N * 3 * heads * d
    assert qm.lib().qmha_forward_ex(C.byref(a)) == 0
    _sync(torch, qm)
    assert torch.equal(o, qm.forward(*(v.contiguous() for v in views), heads, kernel="int8", gran=-1))
