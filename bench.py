#!/usr/bin/env python
"""Headline benchmark: samples/sec of the data-parallel training step on the synthetic 2-task
MLP Problem (BASELINE.json configs[1] at N=1, configs[2] at N>1; batch 4096 per rank).

    python bench.py --gpus N --steps K --warmup W            # this repo's B200 path
    python bench.py --impl reference --gpus N --steps K ...  # reference algorithm on host CPU

One JSON line on stdout (rank 0).  Everything else goes to stderr.
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

METRIC = "samples/sec (whole box) on synthetic 2-task Problem"
WIDTH, N_CLASSES, REG_DIM, DEPTH = 4096, 1000, 64, 3


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--algo", default="sgd", choices=["sgd", "adam", "rmsprop"])
    ap.add_argument("--batch", type=int, default=4096, help="samples per rank per step")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--cpu-batch", type=int, default=1024, help="CPU arm: samples per step (bounded sample)")
    ap.add_argument("--cpu-steps", type=int, default=3, help="cpu_baseline leg: timed steps")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--bucket-mb", type=float, default=None, help="gradient bucket cap (MiB)")
    ap.add_argument("--graph", type=int, default=1, help="replay the step from a CUDA graph (1) or issue it eagerly (0)")
    ap.add_argument("--profile", default=None, help="write a torch.profiler chrome trace of 5 steps here")
    return ap.parse_args()


# --------------------------------------------------------------------------------------------------
# shared: the Problem
# --------------------------------------------------------------------------------------------------

def workload_name(batch, algo):
    return ("2-task MLP Problem: 4096-d in, 3x[Linear(4096,4096)+ReLU] trunk, heads 4096->1000 CE + "
            "4096->64 MSE (54.70M params), batch %d/rank, %s wd=1e-5" % (
                batch, {"sgd": "SGD momentum 0.9 lr 0.01", "adam": "Adam lr 1e-3",
                        "rmsprop": "RMSprop momentum 0.9 lr 1e-3"}[algo]))


def build_problem(ns, save_dir):
    from frl_b200 import synthetic
    return synthetic.make_mlp_problem(ns, save_dir, n_train=64, width=WIDTH, n_classes=N_CLASSES,
                                      reg_dim=REG_DIM, depth=DEPTH)


def run_opts_for(ns, algo, batch):
    t = ns.types
    lr = 0.01 if algo == "sgd" else 1e-3
    return t.RunOpts(optim=t.OptimOpts(algo=t.OptAlgorithm(algo), lr=lr), batchSize=batch,
                     nEpochs=1, numThreads=0, singleThreaded=True, numVisualizedSamples=0)


# --------------------------------------------------------------------------------------------------
# CPU arm: the reference's algorithm (oracle restatement) on the host cores
# --------------------------------------------------------------------------------------------------

def time_cpu_reference(algo, batch, steps, warmup):
    """Reference `_pass_one_minibatch` (oracle/ref_loop.reference_minibatch: stock fp32 torch on
    the CPU, torch.optim, un-fused criterion) on a bounded sample of the workload."""
    import torch
    import frl_b200  # noqa: F401  (only the synthetic Problem definition)
    from frl_b200 import synthetic
    from oracle import ref_loop
    # all the host threads this process may use (torchrun exports OMP_NUM_THREADS=1 by default)
    try:
        torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    except AttributeError:
        torch.set_num_threads(os.cpu_count() or 1)
    ns = synthetic.api_namespace("frl_b200")
    torch.manual_seed(0)
    problem = build_problem(ns, "/tmp/frl_b200_bench_cpu")
    model = problem.get_model()
    crit = problem.get_criterion()
    mods, weights, names = list(crit.loss_modules), list(crit.loss_weights), list(crit.loss_names)
    spec = ref_loop.OptimSpec(algo=algo, lr=0.01 if algo == "sgd" else 1e-3)
    params = list(model.parameters())
    opt = ref_loop.make_optimizer(params, spec)

    def criterion_fn(outputs, targets):
        return ref_loop.parallel_criterion(mods, weights, names, outputs, targets)

    g = torch.Generator().manual_seed(1234)
    x = torch.randn(batch, WIDTH, generator=g)
    y = torch.randint(0, N_CLASSES, (batch,), generator=g)
    r = torch.randn(batch, REG_DIM, generator=g)
    model.train()
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        ref_loop.reference_minibatch(model, criterion_fn, opt, params, 0.0, [x], [(y,), (r,)])
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    total = sum(times)
    return {"value": batch * steps / total, "unit": "samples/s", "cores": torch.get_num_threads(),
            "host_cpus": os.cpu_count(), "kind": "port",
            "sample": "%d timed steps (+%d warm-up) of the same 2-task MLP at batch %d, fp32, "
                      "oracle/ref_loop.reference_minibatch (stock torch CPU ops + torch.optim)" % (
                          steps, warmup, batch),
            "ms_per_step": 1e3 * total / steps, "step_p50_ms": 1e3 * statistics.median(times)}


def main_reference(args, rank):
    if rank != 0:
        return
    res = time_cpu_reference(args.algo, args.cpu_batch, args.steps, args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": res["value"], "unit": "samples/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": res["ms_per_step"], "step_p50_ms": res["step_p50_ms"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": workload_name(args.cpu_batch, args.algo),
                       "note": "reference algorithm on the host CPU; bounded sample: batch %d "
                               "per step instead of %d" % (args.cpu_batch, args.batch)},
            "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": res["value"], "unit": "samples/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------------

class ClockSampler:
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
               0x4: "sw_power_cap", 0x80: "hw_power_brake", 0x2: "applications_clocks_setting"}

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception as e:                      # noqa: BLE001
            log("clock sampling unavailable:", e)
            self._nv = None

    def _run(self):
        nv = self._nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:                       # noqa: BLE001
                pass
            self._stop.wait(0.004)

    def __enter__(self):
        if self._nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()

    def summary(self):
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# --------------------------------------------------------------------------------------------------
# B200 arm
# --------------------------------------------------------------------------------------------------

BYTES_PER_PARAM = {  # algorithmic, fp32 master + state, bf16 gradient read, bf16 shadow written
    ("sgd", "bf16"): 2 + 4 + 4 + 4 + 4 + 2, ("sgd", "fp32"): 4 + 4 + 4 + 4 + 4,
    ("adam", "bf16"): 2 + 4 * 3 + 4 * 3 + 2, ("adam", "fp32"): 4 * 4 + 4 * 3,
    ("rmsprop", "bf16"): 2 + 4 * 3 + 4 * 3 + 2, ("rmsprop", "fp32"): 4 * 4 + 4 * 3}


def main_b200(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    import frl_b200  # noqa: F401
    from frl_b200 import _native, graph_step, synthetic
    from frl_b200.solver import Solver, SolverWorkerArgs, bind_to_gpu_numa_node
    from frl_b200.solver_worker import LossLog
    from frl_b200.types import Device, Precision

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    bind_to_gpu_numa_node(local_rank)         # pinned staging buffers on the GPU's NUMA node
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ns = synthetic.api_namespace("frl_b200")
    t = ns.types
    precision = Precision(args.precision)
    B = args.batch

    torch.manual_seed(0)
    save_dir = "/tmp/frl_b200_bench_%d" % rank
    os.makedirs(save_dir, exist_ok=True)
    problem = build_problem(ns, save_dir)
    wargs = SolverWorkerArgs(run_opts=run_opts_for(ns, args.algo, B), problem=problem,
                             save_dir=save_dir, run_device=Device.GPU, node_idx=0, node_count=1,
                             rank=rank, local_rank=local_rank, world_size=world, group_name=None,
                             init_method="env://", precision=precision)
    if args.bucket_mb is not None:
        os.environ["FRL_B200_BUCKET_MB"] = str(args.bucket_mb)
    os.environ["FRL_B200_CUDA_GRAPH"] = "1" if args.graph else "0"
    worker, _, _ = Solver.build_worker(wargs)
    worker.model.train()
    worker.criterion.train()
    arena = worker.arena
    n_tasks = len(worker.criterion.loss_names)
    log("rank %d: arena %d elements, grad dtype %s, buckets %d" % (
        rank, arena.numel, arena.grad_dtype, len(worker.pipeline.buckets)))

    # ---- synthetic batches (generated on the device; seed 1234 + rank) ----
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    POOL = 4
    pool = []
    for _ in range(POOL):
        x = torch.randn(B, WIDTH, device=dev, generator=gen)
        y = torch.randint(0, N_CLASSES, (B,), device=dev, generator=gen)
        r = torch.randn(B, REG_DIM, device=dev, generator=gen)
        pool.append(([x], [(y,), (r,)]))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        tv = torch.tensor([v], device=dev, dtype=torch.float64)
        dist.all_reduce(tv, op=dist.ReduceOp.MAX)
        return tv.item()

    W, K = args.warmup, args.steps
    log_ring = LossLog(n_tasks, W + K + 32, dev)

    def step_resident(i):
        data, target = pool[i % POOL]
        worker.criterion.set_step_sink(log_ring.row(i), log_ring.nan_flag)
        worker._pass_one_minibatch(i, t.Split.TRAIN, data, target)

    # ======================= leg 1: inputs resident in HBM =======================
    for i in range(W):
        step_resident(i)
    barrier()
    worker.pipeline.update_events.clear()
    worker.pipeline.record_update_events = True
    launches0 = _native.launch_count() + graph_step.REPLAYED_LAUNCHES
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
    with ClockSampler(local_rank) as clocks:
        marks[0].record()
        host_t0 = time.perf_counter()
        for i in range(K):
            step_resident(W + i)
            marks[i + 1].record()
        host_issue_ms = 1e3 * (time.perf_counter() - host_t0) / K     # CPU time to ISSUE a step
        barrier()
    launches = _native.launch_count() + graph_step.REPLAYED_LAUNCHES - launches0
    worker.pipeline.record_update_events = False
    total_ms = marks[0].elapsed_time(marks[K])
    step_ms = [marks[i].elapsed_time(marks[i + 1]) for i in range(K)]
    total_ms = max_over_ranks(total_ms)
    value = world * B * K / (total_ms / 1e3)
    losses = log_ring.rows[W:W + K, 0]
    assert torch.isfinite(losses).all(), "non-finite loss in the timed region"

    # roofline of the dominant kernel of OUR path: the fused update (K2), timed live by events
    # recorded on the launching stream around every launch inside the timed region.  When the
    # update launches live inside the replayed graph (multi-GPU) they cannot carry timing events,
    # so a few eager steps right after the timed region supply them.
    roofline_from = "events around every update launch inside the timed region"
    if not worker.pipeline.update_events:
        graphed, worker.graphed = worker.graphed, None
        worker.pipeline.record_update_events = True
        for i in range(6):
            step_resident(W + K + i)
        barrier()
        worker.pipeline.record_update_events = False
        worker.graphed = graphed
        # drop the first two (cold) eager steps
        per_step = max(len(worker.pipeline.update_events) // 6, 1)
        worker.pipeline.update_events = worker.pipeline.update_events[2 * per_step:]
        roofline_from = "events around the update launches of 4 eager steps run right after the timed region (in the timed region they are nodes of the replayed CUDA graph)"
    upd = worker.pipeline.update_events
    upd_ms = [e0.elapsed_time(e1) for e0, e1, _, _ in upd]
    upd_elems = sum(hi - lo for _, _, lo, hi in upd)
    bpp = BYTES_PER_PARAM[(args.algo, args.precision)]
    peaks_path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    traffic = None
    tpath = os.path.join(REPO, "profiles", "k2_traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get("%s_%s" % (args.algo, args.precision))
    nv_link = worker.pipeline.nvls
    kernel_name = "frl::update_kernel (fused grad-bucket + optimizer, K2)"
    nvlink = None
    if nv_link is not None and upd_ms:
        # K7: per bucket element this GPU's HBM serves its gradient copy to the switch (2 B, bf16),
        # receives the new shadow weight (2 B) and streams master + state for its 1/world shard
        state_bytes = bpp - 2 - 2 if args.precision == "bf16" else bpp - 4
        bpp = (4 if args.precision == "bf16" else 8) + state_bytes / world
        kernel_name = ("frl::nvls_update (K7: multimem.ld_reduce + sharded update + multimem.st; "
                       "launch time includes its two cross-GPU barriers)")
        g_b = 2 if args.precision == "bf16" else 4
        link_bytes = upd_elems * g_b * (1.0 + 1.0 / world)          # per direction, per GPU
        nvlink = {"bytes_per_direction_per_launch": link_bytes / len(upd_ms),
                  "achieved": link_bytes / (sum(upd_ms) / 1e3) / 1e9, "peak": 770.0, "unit": "GB/s",
                  "peak_source": "measured peer copy per direction (B200_PROFILING.md); 900 nominal",
                  "note": "NVLink 5 per-direction payload of the fused step: out = own gradient copy "
                          "read by the switch + multicast of the shard's new weights, in = reduced shard "
                          "+ every shard's new weights; this, the barriers and the SMs left over by the "
                          "overlapped backward GEMMs bound K7, not HBM"}
    achieved = bpp * upd_elems / (sum(upd_ms) / 1e3) / 1e9 if upd_ms else None
    roofline = {"bound": "hbm", "kernel": kernel_name,
                "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": (achieved / peak) if achieved else None, "traffic": traffic,
                "peak_source": peak_src, "bytes_per_param": bpp, "timed_by": roofline_from,
                "elems_per_launch": upd_elems / max(len(upd), 1),
                "avg_launch_ms": (sum(upd_ms) / len(upd_ms)) if upd_ms else None,
                "update_ms_per_step": (sum(upd_ms) / len(upd_ms)) * (arena.numel / (upd_elems / len(upd_ms))) if upd_ms else None}
    if roofline["update_ms_per_step"]:
        roofline["share_of_step"] = roofline["update_ms_per_step"] / (total_ms / K)
    if nvlink is not None:
        roofline["nvlink"] = nvlink
        roofline["traffic"] = None
        roofline["note"] = ("multi-GPU: the update is sharded 1/world per rank and overlapped with "
                            "backward; the single-GPU run carries the HBM roofline of the update kernel (K2)")

    if args.profile:
        # every rank runs the steps (collectives!); only rank 0 records
        from contextlib import nullcontext
        from torch.profiler import ProfilerActivity, profile
        ctx = profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) if rank == 0 else nullcontext()
        with ctx as prof:
            for i in range(5):
                step_resident(W + K + i)
            torch.cuda.synchronize()
        if rank == 0:
            prof.export_chrome_trace(args.profile)
            log(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25))
    barrier()

    # ======================= leg 2: end to end through the public loop =======================
    e2e = None
    if not args.no_e2e:
        # What a user's Solver.solve() runs per epoch: SolverWorker._pass_one_epoch over the
        # Problem's own dataset.  The dataset (fp32) lives in pinned HOST memory behind
        # `problem.datasets`; the loop's DeviceBatchLoader shuffles it with the reference's sampler
        # machinery, moves the rows of the next batches to HBM while the current one trains
        # (FRL_B200_INPUT_PATH: native host gather threads + one DMA per field, or GPU-pulled over
        # PCIe) and runs the transform on the device (frl_preproc_affine -> bf16).  Inside the
        # timed region, every step: H2D of that step's inputs, the step, the loss row landing in
        # pinned host memory (read by the host 2 steps late for the NaN guard), and the loop's
        # bookkeeping — retained batches, per-sample metrics of the Problem's
        # compute_batch_metrics hook read back every metricAmortizationSchedule (10) steps, epoch
        # summary.  One epoch of K steps is timed, after one warm-up epoch over the same loader.
        from frl_b200.device_loader import DeviceBatchLoader
        from frl_b200 import synthetic as syn
        from frl_b200.types import Mode

        from frl_b200.sampler import ScaffoldSampler

        class LocalShardSampler(ScaffoldSampler):
            """world > 1: every rank owns a node-local shard of the dataset and reshuffles it
            per epoch (the ScaffoldSampler contract — set_epoch, permutation from a generator
            seeded per epoch — on local indices; a global index space would need world x the
            pinned memory per rank)."""

            def __init__(self, n, seed):                    # no DistributedSampler bookkeeping
                torch.utils.data.Sampler.__init__(self)
                self.n, self.seed, self.epoch = n, seed, 0

            def __len__(self):
                return self.n

            def rank_index_tensor(self):
                g = torch.Generator().manual_seed(self.seed + self.epoch)
                return torch.randperm(self.n, generator=g)

        L = K if K <= 128 else K // ((K + 127) // 128)        # steps per epoch
        n_epochs_timed = max(1, K // L)
        n_host = L * B
        host_problem = syn.make_mlp_problem(ns, save_dir, n_train=n_host, width=WIDTH,
                                            n_classes=N_CLASSES, reg_dim=REG_DIM, depth=DEPTH,
                                            pinned=True, fast_fields=True)
        host_ds = host_problem.datasets[0]
        out_dtype = torch.bfloat16 if precision == Precision.BF16 else torch.float32
        loader = DeviceBatchLoader(host_ds, batch_size=B, device=dev, out_dtype=out_dtype,
                                   sampler=LocalShardSampler(n_host, 1000 * rank) if world > 1 else None)
        loaders = {t.Split.TRAIN: loader}
        h2d_bytes = loader.h2d_bytes_per_batch
        worker.cur_epoch = 1
        worker._pass_one_epoch(host_problem, loaders, Mode.TRAIN)        # warm-up epoch (graph capture)
        barrier()
        m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches_e0 = _native.launch_count() + graph_step.REPLAYED_LAUNCHES
        m0.record()
        for ep in range(n_epochs_timed):
            worker.cur_epoch = 2 + ep
            stats = worker._pass_one_epoch(host_problem, loaders, Mode.TRAIN)
        m1.record()
        barrier()
        e2e_launches = _native.launch_count() + graph_step.REPLAYED_LAUNCHES - launches_e0
        e2e_steps = n_epochs_timed * L
        e2e_ms = max_over_ranks(m0.elapsed_time(m1))
        ep_losses = stats[t.Split.TRAIN].losses
        assert all(v == v for v in ep_losses.values()), "NaN loss in the e2e leg"
        if args.profile and rank == 0:
            from torch.profiler import ProfilerActivity, profile
            with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
                worker.cur_epoch += 1
                worker._pass_one_epoch(host_problem, loaders, Mode.TRAIN)
                torch.cuda.synchronize()
            prof.export_chrome_trace(args.profile.replace(".json", "") + "_e2e.json")
            log(prof.key_averages().table(sort_by="cuda_time_total", row_limit=20))
        n_metrics = len(stats[t.Split.TRAIN].metrics)
        e2e = {"value": world * B * e2e_steps / (e2e_ms / 1e3), "unit": "samples/s",
               "h2d_bytes_per_step": h2d_bytes,
               "d2h_bytes_per_step": 4 * (1 + n_tasks) + 4 * n_metrics * B,
               "ms_per_step": e2e_ms / e2e_steps, "steps": e2e_steps, "input_path": loader.path,
               "input_threads": loader.threads, "input_blocks": loader.blocks,
               "input_wire": {k: str(v).replace("torch.", "") for k, v in loader._wire_dtype.items()},
               "gpu_launches": e2e_launches,
               "epoch_losses": {k: float(v) for k, v in ep_losses.items()},
               "how": "SolverWorker._pass_one_epoch (the loop Solver.solve runs per epoch) over the "
                      "Problem's dataset (fp32, %d batches) in pinned host memory: reference sampler "
                      "indices -> DeviceBatchLoader moves the rows of the next batches to HBM while "
                      "the current step runs (%s) -> transform + bf16 cast on device "
                      "(frl_preproc_affine) -> _pass_one_minibatch (CUDA-graph replay) -> loss row "
                      "written to pinned host memory by the criterion kernel, read 2 steps late; "
                      "includes the loop's retained-batch bookkeeping, the Problem's per-sample "
                      "metric hook every 10 steps (D2H) and the epoch summary" % (
                          L, "native host gather threads into pinned staging + one DMA per field"
                          if loader.path == "host" else
                          "rows pulled over PCIe by frl_gather_rows%s on a copy stream" % (
                              "_tma" if loader.path == "tma" else ""))}

    # ======================= CPU baseline (rank 0, N=1) =======================
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        res = time_cpu_reference(args.algo, args.cpu_batch, args.cpu_steps, 1)
        cpu = {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")}

    nv = worker.pipeline.nvls
    if world == 1:
        grad_sync_desc = "none (1 GPU)"
    elif nv is not None:
        grad_sync_desc = ("fused per-bucket NVLS kernel over NVSwitch multicast (multimem.ld_reduce of the "
                          "bf16 grads + sharded update + multimem.st of the new weights), %d buckets, "
                          "%d CTAs, barriers %s" % (len(worker.pipeline.buckets), nv.max_blocks,
                                                    "as separate 1-CTA launches" if nv.flags & 1 else "in-kernel"))
    else:
        grad_sync_desc = "NCCL all-reduce in place on bf16 arena buckets + fused update per bucket"
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": K,
                "warmup": W, "ms_per_step": total_ms / K, "step_p50_ms": statistics.median(step_ms),
                "host_issue_ms_per_step": host_issue_ms,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16" if precision == Precision.BF16 else "f32", "data": "synthetic",
                "config": {"workload": workload_name(B, args.algo), "global_batch": B * world,
                           "step_issue": "CUDA graph replay" if args.graph else "eager",
                           "parallelism": "dp%d" % world,
                           "precision": "bf16 forward/backward + bf16 grads, fp32 master weights and "
                                        "optimizer state" if precision == Precision.BF16 else "fp32",
                           "l2": "no flush needed: each step streams the 54.7M-element arena "
                                 "(>= 1.1 GB) and 4 rotating input batches, far larger than the 126 MB L2",
                           "grad_allreduce": grad_sync_desc},
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches,
                "clocks": clocks.summary(), "final_loss": float(losses[-1])}
        print(json.dumps(line), flush=True)
    if world > 1:
        # Release the captured graphs (they hold NCCL work) before anything NCCL is torn down,
        # line the ranks up, and leave without running communicator destructors: a rank that
        # exits early while another still tears down captured collectives can hang the job.
        import gc
        worker.graphed = None
        gc.collect()
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        main_reference(args, rank)
        return
    if world != args.gpus:
        log("note: --gpus %d but WORLD_SIZE %d; using WORLD_SIZE" % (args.gpus, world))
    main_b200(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
