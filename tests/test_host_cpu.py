"""CPU-only checks of the host side: the C-ABI library loads and exports every symbol declared
in include/qmha.h, argument validation / error reporting, the loud no-GPU failure, and the
(batch × head) sharding used for multi-GPU runs (world_size 2 over gloo)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def qm():
    import quantizedmha_b200 as q
    if not os.path.exists(q.lib_path()):
        import __graft_entry__ as g
        g.build()
    return q


def test_library_exports_every_declared_symbol(qm):
    hdr = open(os.path.join(ROOT, "include", "qmha.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(solve|qmha_[a-z0-9_]+)\s*\(", hdr))
    assert {"solve", "qmha_forward", "qmha_forward_host", "qmha_quantize_qkv", "qmha_attention_prepared",
            "qmha_quantize_blocks", "qmha_quantize_static", "qmha_convert_qkv_f16", "qmha_last_error"} <= names
    L = ctypes.CDLL(qm.lib_path())
    for n in sorted(names):
        assert hasattr(L, n), f"{n} declared in include/qmha.h but not exported"


def _build_c_consumer(qm, tmp):
    """gcc (C99, warnings are errors) on tests/c/abi_consumer.c against include/qmha.h and libqmha.so."""
    exe = os.path.join(str(tmp), "abi_consumer")
    libdir = os.path.dirname(qm.lib_path())
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(cuda, "include"),
                    os.path.join(ROOT, "tests", "c", "abi_consumer.c"), "-o", exe, "-L", libdir, "-lqmha",
                    "-L", os.path.join(cuda, "lib64"), "-lcudart", "-lm", f"-Wl,-rpath,{libdir}",
                    f"-Wl,-rpath,{os.path.join(cuda, 'lib64')}"], check=True)
    return exe


def test_header_is_valid_c_and_the_argument_block_matches_the_library(qm, tmp_path):
    """The boundary is a C ABI: a C99 translation unit includes qmha.h, links every entry it uses and sees the same
    qmha_args size as the library and as the ctypes mirror."""
    exe = _build_c_consumer(qm, tmp_path)
    r = subprocess.run([exe, "sizes"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stdout + r.stderr
    c_size, stamped, kid = (int(x) for x in r.stdout.split())
    assert c_size == stamped == ctypes.sizeof(qm.QmhaArgs) and kid == qm.KERNEL_INT8


def test_launchers_header_is_the_reference_signature():
    txt = open(os.path.join(ROOT, "include", "qmha.h")).read()
    flat = " ".join(txt.split())
    assert "void solve(const float* Q, const float* K, const float* V, float* output, int N, int d_model, int h);" in flat


def test_kernel_name_aliases(qm):
    for n in ("fa_tc_int8_b", "fa_tc_int8_a", "int8", "fa_b200_int8"):
        assert qm.kernel_id(n) == qm.KERNEL_INT8
    for n in ("fa_tc_v2a", "fa_tc_v1a", "fa", "unfused", "f16", "fa_b200_f16"):
        assert qm.kernel_id(n) == qm.KERNEL_F16
    with pytest.raises(qm.QmhaError):
        qm.kernel_id("no_such_kernel")


def test_extended_entry_host_logic(qm):
    """qmha_forward_ex argument block (struct_size guard, defaults), the BF16 kernel names and the choice of the
    scale granularity for a shape — host logic only, no device needed."""
    import ctypes as C
    L = qm.lib()
    assert qm.kernel_id("bf16") == qm.KERNEL_BF16 and qm.kernel_id("fa_b200_bf16") == qm.KERNEL_BF16
    a = qm.QmhaArgs()
    L.qmha_args_init(C.byref(a))
    assert a.struct_size == C.sizeof(qm.QmhaArgs), "ctypes mirror of qmha_args is out of step with include/qmha.h"
    assert (a.kernel, a.gran, a.rope, a.variant, a.in_dtype, a.out_dtype) == (-1, -1, -1, -1, qm.DTYPE_F32, qm.DTYPE_F32)
    a.struct_size = 8
    assert L.qmha_forward_ex(C.byref(a)) != 0 and b"struct_size" in L.qmha_last_error()
    assert L.qmha_forward_ex(None) != 0
    # block scales while the per-block table of one unit fits in shared memory behind the tiles, per-head beyond
    assert L.qmha_granularity_for(8192, 4096, 32) == qm.GRAN_BLOCK
    assert L.qmha_granularity_for(16384, 4096, 32) == qm.GRAN_BLOCK
    assert L.qmha_granularity_for(56320, 128, 1) == qm.GRAN_BLOCK and L.qmha_granularity_for(56321, 128, 1) == qm.GRAN_HEAD
    assert L.qmha_granularity_for(512, 30, 2) == qm.GRAN_HEAD          # d = 15: scalar two-pass path
    assert L.qmha_default_granularity(4096, 32) == qm.GRAN_BLOCK


def test_workspace_dims_and_shape_validation(qm):
    assert qm.workspace_dims(8192, 4096, 32) == (8192, 128)
    assert qm.workspace_dims(50, 64, 8) == (256, 32)
    assert qm.workspace_dims(4096, 512, 8) == (4096, 64)
    with pytest.raises(qm.QmhaError, match="divisible"):
        qm.workspace_dims(128, 100, 3)
    with pytest.raises(qm.QmhaError, match="<= 128"):
        qm.workspace_dims(128, 512, 2)


def test_no_gpu_means_loud_failure_not_fallback(qm):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    rc = qm.lib().qmha_forward(None, None, None, None, 1, 8, 32, 4, 0, 1, None)
    assert rc != 0 and b"no CPU fallback" in qm.lib().qmha_last_error()
    rc = qm.lib().qmha_forward_host(None, None, None, None, 1, 8, 32, 4, 0, 1)
    assert rc != 0


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "quantizedmha_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in src.lower().replace("# oracle", ""), f"{f} references the oracle"


def test_unit_sharding_partitions_exactly():
    from quantizedmha_b200.sharding import shard_slabs, unit_range
    for B, H, W in [(8, 32, 8), (8, 32, 3), (1, 8, 2), (5, 7, 4), (2, 3, 8)]:
        seen = []
        for r in range(W):
            lo, hi = unit_range(B * H, W, r)
            got = [b * H + h for (b, h0, h1) in shard_slabs(B, H, W, r) for h in range(h0, h1)]
            assert got == list(range(lo, hi))
            seen += got
        assert seen == list(range(B * H))


_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import numpy as np, torch, torch.distributed as dist
from oracle import load_oracle
from quantizedmha_b200.sharding import shard_slabs, slab_view, gather_outputs, forward_sharded
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
orc = load_oracle()
B, N, H, d = 2, 48, 3, 8
q, k, v = orc.profile_inputs(B * N, H * d)
q, k, v = (a.reshape(B, N, H * d) for a in (q, k, v))
full = orc.mha(q, k, v, H, threads=1)
mine = np.zeros_like(full)
for (b, h0, h1) in shard_slabs(B, H, world, rank):
    o = orc.mha(np.ascontiguousarray(slab_view(q, b, h0, h1, H)), np.ascontiguousarray(slab_view(k, b, h0, h1, H)),
                np.ascontiguousarray(slab_view(v, b, h0, h1, H)), h1 - h0, threads=1)
    mine[b, :, h0 * d:h1 * d] = o
t = torch.from_numpy(mine)
dist.all_reduce(t)  # disjoint shards: the sum is the concatenation (test-side gather only)
times = torch.tensor([float(rank + 1)])
dist.all_reduce(times, op=dist.ReduceOp.MAX)
assert times.item() == float(world)
assert np.array_equal(t.numpy(), full), "sharded result differs from the unsharded oracle"
# the optional output gather (SURVEY 8f row 4): every rank ends up with the whole tensor
g = gather_outputs(torch.from_numpy(mine.copy()), B, H)
assert np.array_equal(g.numpy(), full), "gathered result differs from the unsharded oracle"
# sharded forward with the chunked, overlapped gather (same code path as on NCCL; the CPU oracle computes the slabs)
fwd = lambda a, b, c, heads: torch.from_numpy(orc.mha(a.numpy(), b.numpy(), c.numpy(), heads, threads=1))
for chunks in (1, 2, 5):
    o = forward_sharded(torch.from_numpy(q), torch.from_numpy(k), torch.from_numpy(v), H, chunks=chunks, forward_fn=fwd)
    assert np.array_equal(o.numpy(), full), f"forward_sharded(chunks={{chunks}}) differs from the unsharded oracle"
# fused gather (sharding.forward_fused_gather): every rank's launch plan writes its slabs into ALL replicas.  The
# replicas are file-backed shared memory here and a "peer address" is (rank + 1) << 40 | byte offset, so the slab
# offsets and strides the kernel would get are exercised across real processes; the fences run over gloo.
from quantizedmha_b200.sharding import forward_fused_gather
shape = (B, N, H * d)
paths = [os.path.join(os.environ["REP_DIR"], "rep%d.bin" % r) for r in range(world)]
np.memmap(paths[rank], dtype=np.float32, mode="w+", shape=shape)[...] = np.nan
dist.barrier()
mm = [np.memmap(p, dtype=np.float32, mode="r+", shape=shape) for p in paths]
class Rep:
    pass
rep = Rep()
rep.world, rep.rank, rep.local = world, rank, torch.from_numpy(mm[rank])
rep.peer_base = dict((r, (r + 1) << 40) for r in range(world) if r != rank)
rep.fence = dist.barrier
def fused(qs, ks, vs, heads, out_view, peers):
    o = orc.mha(qs.numpy(), ks.numpy(), vs.numpy(), heads, threads=1)
    out_view.copy_(torch.from_numpy(o))
    assert len(peers) == world - 1
    for addr in peers:
        r, off = (addr >> 40) - 1, (addr & ((1 << 40) - 1)) // 4
        dst = np.lib.stride_tricks.as_strided(mm[r].reshape(-1)[off:], shape=o.shape, strides=(N * H * d * 4, H * d * 4, 4))
        dst[...] = o
for Bx in (B,):
    o = forward_fused_gather(torch.from_numpy(q), torch.from_numpy(k), torch.from_numpy(v), H, rep, forward_fn=fused)
    assert np.array_equal(np.asarray(mm[rank]), full), "fused gather: this rank's replica differs from the unsharded oracle"
dist.barrier()
dist.destroy_process_group()
print("ok", rank)
"""


def test_slab_stride_detection_of_the_ctypes_mirror(qm):
    """binding._out_strides: (row pitch, batch pitch) in elements of a tensor that may be a (batch, head-range) view;
    what forward() hands to qmha_args.o_row_stride / in_row_stride.  Pure host logic: CPU tensors suffice."""
    import torch
    f = qm.binding._out_strides
    full = torch.zeros((4, 100, 640))
    assert f(full, 4, 100, 640) == (640, 100 * 640)
    assert f(full[1:3, :, 128:384], 2, 100, 256) == (640, 100 * 640)            # heads 1..2 of batch entries 1..2
    assert f(full[2, :, 128:384], 1, 100, 256) == (640, 100 * 640)              # 2-D view of one batch entry
    assert f(full[:, :60, :], 4, 60, 640) == (640, 100 * 640)                   # fewer rows than the parent
    assert f(torch.zeros((100, 256)), 1, 100, 256) == (256, 100 * 256)
    assert f(torch.zeros((1, 1, 64)), 1, 1, 64) == (64, 64)                     # N = 1: nothing to stride over
    for bad, args in ((torch.zeros((256, 100)).t(), (1, 100, 256)),              # last dimension not contiguous
                      (torch.zeros((2, 100, 256)), (1, 100, 256)),               # wrong shape
                      (torch.zeros((100, 256)), (2, 100, 256))):                 # 2-D tensor for a batched call
        with pytest.raises(qm.QmhaError):
            f(bad, *args)
    a = qm.QmhaArgs()
    qm.lib().qmha_args_init(ctypes.byref(a))
    assert (a.o_row_stride, a.o_batch_stride, a.in_row_stride, a.in_batch_stride, a.n_peers, a.device) == (0, 0, 0, 0, 0, -1)


def test_fused_gather_launch_plan_covers_every_unit_once():
    from quantizedmha_b200.sharding import launch_plan, slab_offset, unit_range
    for B, H, W in [(8, 32, 8), (32, 32, 8), (8, 32, 3), (1, 8, 2), (5, 7, 4), (2, 3, 8)]:
        seen = []
        for r in range(W):
            plan = launch_plan(B, H, W, r)
            got = [b * H + h for (b0, b1, h0, h1) in plan for b in range(b0, b1) for h in range(h0, h1)]
            assert got == list(range(*unit_range(B * H, W, r)))
            assert all(b1 - b0 == 1 or (h0, h1) == (0, H) for (b0, b1, h0, h1) in plan)   # only whole entries are merged
            seen += got
        assert seen == list(range(B * H))
    assert launch_plan(32, 32, 8, 7) == [(28, 32, 0, 32)]          # C5 over 8 GPUs: one launch per rank
    assert slab_offset(2, 3, 100, 8, 16) == (2 * 100 * 8 + 3) * 16


def test_two_rank_sharding_over_gloo(tmp_path):
    """The multi-GPU path is 'each rank runs its own units, no collective on the data path'.
    Two CPU ranks each compute their shard with the oracle; the union must equal the whole."""
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT))
    port = 29500 + (os.getpid() % 400)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   REP_DIR=str(tmp_path))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT))
    for p in procs:
        out, _ = p.communicate(timeout=180)
        assert p.returncode == 0, out.decode()


def test_bench_reference_arm_prints_contract_line():
    """bench.py --impl reference runs the reference's CPU path on the host cores only."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--cpu-sample-n", "256", "--cpu-sample-heads", "2"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr
    import json
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "TFLOP/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["e2e"]["h2d_bytes_per_step"] == 0


def test_compare_ncu_report_on_the_committed_summaries():
    """tools/compare_ncu.py (the reference's tools/compare_ncu.py does the same on text
    tables): metric union, unit normalisation and delta columns, on the ncu summaries under profiles/."""
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("compare_ncu", os.path.join(root, "tools", "compare_ncu.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    a = mod.load(os.path.join(root, "profiles", "r01", "ncu_attn_fwd_int8_block_c4.txt"))
    b = mod.load(os.path.join(root, "profiles", "r01", "ncu_block_quantize_c4.txt"))
    assert a["gpu__time_duration.sum"][1] == "s" and 1e-3 < a["gpu__time_duration.sum"][0] < 1e-2   # ms -> s
    assert b["gpu__time_duration.sum"][0] < 1e-3                                                     # us -> s
    assert a["dram__bytes_read.sum"][1] == "byte" and a["dram__bytes_read.sum"][0] > 1e9
    md = mod.report([a, b], ["attn", "quant"])
    row = next(l for l in md.splitlines() if "gpu__time_duration.sum" in l)
    assert row.count("|") == 6 and row.rstrip().endswith("% |") and "-8" in row   # quant is ~86 % shorter
    raw = ('"ID","Kernel Name","gpu__time_duration.sum","dram__bytes_read.sum"\n"","","us","Mbyte"\n'
           '"0","k_a","10","2"\n"1","k_b","30","4"\n"2","k_a","20","6"\n')
    c = mod.parse_raw_csv(raw, kernel="k_a")
    assert abs(c["gpu__time_duration.sum"][0] - 15e-6) < 1e-12 and c["dram__bytes_read.sum"] == (4e6, "byte")
