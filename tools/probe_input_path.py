#!/usr/bin/env python
"""Stand-alone timing of the candidate host->HBM row-gather paths (one GPU).

    python tools/probe_input_path.py [rows_per_batch] [row_elems]

Prints GB/s of: contiguous cudaMemcpyAsync (PCIe reference), frl_gather_rows at several grid
sizes, frl_gather_rows_tma, the native host gather pool per thread count (plain and with the
fp32 -> bf16 wire conversion), and the host cost of drawing one index batch from the DataLoader
machinery.
"""
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch  # noqa: E402
import frl_b200  # noqa: E402,F401
from frl_b200 import _native  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
W = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
NB = 16
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
try:
    from frl_b200.solver import bind_to_gpu_numa_node
    bind_to_gpu_numa_node(0)
except Exception as e:  # noqa: BLE001
    print("numa bind failed:", e)
src = torch.randn(NB * B, W).pin_memory()
dst = torch.empty(B, W, device=dev)
nbytes = B * W * 4
print("rows %d x %d B = %.1f MB per batch" % (B, W * 4, nbytes / 1e6))


def timed(fn, reps=6, sync_host=False):
    best = 1e9
    for r in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        fn(r)
        e1.record()
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        ms = e0.elapsed_time(e1)
        if sync_host:
            ms = (t1 - t0) * 1e3
        if r > 0:
            best = min(best, ms)
    return best


def report(name, ms):
    print("%-42s %8.3f ms  %7.1f GB/s" % (name, ms, nbytes / ms / 1e6), flush=True)


perms = [torch.randperm(NB * B)[:B].contiguous() for _ in range(8)]
perms_dev = [p.to(dev) for p in perms]
perms_pin = [p.pin_memory() for p in perms]

report("cudaMemcpyAsync contiguous", timed(lambda r: dst.copy_(src[r * B:(r + 1) * B], non_blocking=True)))
for blocks in (32, 64, 148, 296, 592, 1184):
    report("frl_gather_rows blocks=%d" % blocks,
           timed(lambda r: _native.gather_rows(src, perms_dev[r], dst, max_blocks=blocks)))
ok = torch.equal(dst.cpu(), src[perms[5]])
print("gather_rows correct:", ok)
for blocks in (16, 37, 74, 148, 296):
    try:
        dst.zero_()
        report("frl_gather_rows_tma blocks=%d" % blocks,
               timed(lambda r: _native.gather_rows_tma(src, perms_dev[r], dst, max_blocks=blocks)))
        print("   correct:", torch.equal(dst.cpu(), src[perms[5]]))
    except Exception as e:  # noqa: BLE001
        print("tma variant failed:", e)
        break
side = torch.cuda.Stream()
torch.cuda.set_stream(side)          # batched copies are not allowed on the legacy default stream
for blocks in (1, 2, 3, 4, 8):
    report("frl_gather_rows_tma blocks=%d" % blocks,
           timed(lambda r: _native.gather_rows_tma(src, perms_dev[r], dst, max_blocks=blocks)))
for blocks in (2, 4, 8, 16):
    report("frl_gather_rows blocks=%d" % blocks,
           timed(lambda r: _native.gather_rows(src, perms_dev[r], dst, max_blocks=blocks)))
# (cudaMemcpyBatchAsync with one descriptor per row was measured once and rejected: 7.4 GB/s,
#  19 ms of host time per 4096-row submission; see profiles/r1b_probe_input_path_b.log)
stage = torch.empty(B, W).pin_memory()
for th in (1, 4, 8, 16, 32):
    pool = _native.HostGatherPool(th)
    t = []
    for r in range(5):
        t0 = time.perf_counter()
        pool.wait(pool.submit(src, perms[r], stage))
        t.append(time.perf_counter() - t0)
    pool.close()
    report("frl_gather_pool threads=%d (host)" % th, min(t[1:]) * 1e3)
print("   correct:", torch.equal(stage, src[perms[4]]))
stage16 = torch.empty(B, W, dtype=torch.bfloat16).pin_memory()
pool = _native.HostGatherPool(16)
t = []
for r in range(5):
    t0 = time.perf_counter()
    pool.wait(pool.submit_f32_to_bf16(src, perms[r], stage16))
    t.append(time.perf_counter() - t0)
pool.close()
report("frl_gather_pool f32->bf16 wire, 16 threads", min(t[1:]) * 1e3)
print("   correct:", torch.equal(stage16, src[perms[4]].to(torch.bfloat16)))

# host cost of the index stream (DataLoader + sampler machinery, as DeviceBatchLoader uses it)
from frl_b200.device_loader import _IndexOnly, _collate_indices  # noqa: E402
import torch.utils.data as tud  # noqa: E402
ld = tud.DataLoader(_IndexOnly(NB * B), batch_size=B, shuffle=True, num_workers=0, collate_fn=_collate_indices)
t0 = time.perf_counter()
n = 0
for idx in ld:
    n += 1
t1 = time.perf_counter()
print("index DataLoader: %.3f ms per batch of %d (host)" % ((t1 - t0) * 1e3 / n, B))
samp = tud.RandomSampler(_IndexOnly(NB * B))
t0 = time.perf_counter()
allidx = torch.tensor(list(iter(samp)), dtype=torch.int64)
t1 = time.perf_counter()
print("list(sampler) -> tensor once per epoch: %.3f ms per batch-equivalent" % ((t1 - t0) * 1e3 / NB))
