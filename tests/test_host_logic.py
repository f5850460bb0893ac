"""Host-side logic of the arena, the fused optimizers and the bucket pipeline on CPU tensors.
The kernel entry points are replaced by ``oracle.optim_np.KernelDouble`` (a test double with the
same call signatures) — this exercises layout, bucketing, state (de)serialisation and the
multi-rank protocol, not the kernels."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn

import frl_b200  # noqa: F401
from frl_b200 import fused_optim, grad_sync
from frl_b200.arena import ALIGN_ELEMS, ParamArena
from frl_b200.types import OptAlgorithm, OptimOpts, Precision
from oracle.optim_np import KernelDouble


@pytest.fixture()
def double(monkeypatch):
    d = KernelDouble()
    monkeypatch.setattr(fused_optim, "KERNELS", d)
    monkeypatch.setattr(grad_sync, "KERNELS", d)
    return d


def _net(seed=0):
    torch.manual_seed(seed)
    return nn.Sequential(nn.Linear(7, 5), nn.ReLU(), nn.Linear(5, 3), nn.ReLU(), nn.Linear(3, 2))


def test_arena_layout_and_views():
    net = _net()
    before = [p.detach().clone() for p in net.parameters()]
    extra = nn.Parameter(torch.tensor([0.5, -0.25]))
    arena = ParamArena(net.parameters(), [extra], device="cpu")
    assert arena.numel % ALIGN_ELEMS == 0 and arena.model_end % ALIGN_ELEMS == 0
    for s in arena.slots:
        assert s.offset % ALIGN_ELEMS == 0
    for p, b in zip(net.parameters(), before):
        assert torch.equal(p, b)
        s = arena.slot_of(p)
        assert p.data_ptr() == arena.master[s.offset:].data_ptr()     # a view, not a copy
    assert arena.slots[-1].is_model is False and arena.slots[-1].offset == arena.model_end
    # padding stays zero
    used = torch.zeros(arena.numel, dtype=torch.bool)
    for s in arena.slots:
        used[s.offset:s.end] = True
    assert torch.all(arena.master[~used] == 0)
    # buckets cover the arena back to front
    buckets = arena.buckets(cap_bytes=64)
    assert buckets[0][1] == arena.numel and buckets[-1][0] == 0
    assert all(a[0] == b[1] for a, b in zip(buckets[:-1], buckets[1:]))
    with arena.exported():
        sd = {k: v.clone() for k, v in net.state_dict().items()}
        assert all(v.data_ptr() != arena.master.data_ptr() for v in sd.values())
    for p in net.parameters():
        assert p.data_ptr() >= arena.master.data_ptr()


def test_bf16_arena_keeps_fp32_master_and_bf16_views():
    net = _net()
    arena = ParamArena(net.parameters(), device="cpu", precision=Precision.BF16)
    assert arena.grad.dtype == torch.bfloat16 and arena.lp.dtype == torch.bfloat16
    for p in net.parameters():
        assert p.dtype == torch.bfloat16
    assert arena.master.dtype == torch.float32
    assert torch.equal(arena.lp.float(), arena.master.to(torch.bfloat16).float())


@pytest.mark.parametrize("algo,kw", [
    (OptAlgorithm.SGD, {}), (OptAlgorithm.ADAM, {}), (OptAlgorithm.ADAM, {"amsgrad": True}),
    (OptAlgorithm.RMSPROP, {})])
def test_fused_optimizer_follows_torch_optim_and_round_trips_state(double, algo, kw):
    from oracle.ref_loop import OptimSpec, make_optimizer
    net, ref = _net(1), _net(1)
    opts = OptimOpts(algo=algo, lr=0.01, **kw)
    arena = ParamArena(net.parameters(), device="cpu")
    opt = fused_optim.create_fused_optimizer(arena, opts)
    pipe = grad_sync.GradBucketPipeline(arena, opt)
    ref_opt = make_optimizer(ref.parameters(), OptimSpec(algo=algo.value, lr=0.01,
                                                         amsgrad=kw.get("amsgrad", False)))
    x = torch.randn(16, 7)
    for step in range(4):
        for m, o in ((net, None), (ref, ref_opt)):
            loss = m(x).square().mean()
            if o is None:
                pipe.begin_step()
                loss.backward()
                pipe.finish_step()
            else:
                o.zero_grad()
                loss.backward()
                o.step()
    for a, b in zip(net.parameters(), ref.parameters()):
        np.testing.assert_allclose(a.detach().numpy(), b.detach().numpy(), rtol=1e-5, atol=1e-7)
    # state dict is torch-format and interchangeable
    sd = opt.state_dict()
    ref_sd = ref_opt.state_dict()
    assert sorted(sd["state"].keys()) == sorted(ref_sd["state"].keys())
    for k, entry in ref_sd["state"].items():
        for name, val in entry.items():
            if torch.is_tensor(val) and val.dim() > 0:
                np.testing.assert_allclose(sd["state"][k][name].numpy(), val.numpy(),
                                           rtol=1e-4, atol=1e-6)
    # load the torch optimizer's state into a fresh fused optimizer and keep stepping in sync
    net2 = _net(1)
    with torch.no_grad():
        for a, b in zip(net2.parameters(), ref.parameters()):
            a.copy_(b)
    arena2 = ParamArena(net2.parameters(), device="cpu")
    opt2 = fused_optim.create_fused_optimizer(arena2, opts)
    opt2.load_state_dict(ref_sd)
    pipe2 = grad_sync.GradBucketPipeline(arena2, opt2)
    pipe2.begin_step(); net2(x).square().mean().backward(); pipe2.finish_step()
    ref_opt.zero_grad(); ref(x).square().mean().backward(); ref_opt.step()
    for a, b in zip(net2.parameters(), ref.parameters()):
        np.testing.assert_allclose(a.detach().numpy(), b.detach().numpy(), rtol=1e-5, atol=1e-7)


def test_clip_applies_to_model_parameters_only(double):
    net, ref = _net(2), _net(2)
    extra, ref_extra = nn.Parameter(torch.tensor([1.0])), nn.Parameter(torch.tensor([1.0]))
    arena = ParamArena(net.parameters(), [extra], device="cpu")
    opt = fused_optim.create_fused_optimizer(arena, OptimOpts(algo=OptAlgorithm.SGD, lr=0.1))
    pipe = grad_sync.GradBucketPipeline(arena, opt, clip_norm=0.05)
    ref_opt = torch.optim.SGD(list(ref.parameters()) + [ref_extra], lr=0.1, momentum=0.9,
                              weight_decay=1e-5)
    x = torch.randn(8, 7)
    pipe.begin_step(); (net(x).square().mean() * extra.sum() * 3).backward(); pipe.finish_step()
    (ref(x).square().mean() * ref_extra.sum() * 3).backward()
    torch.nn.utils.clip_grad_norm_(ref.parameters(), 0.05)
    ref_opt.step()
    for a, b in zip(list(net.parameters()) + [extra], list(ref.parameters()) + [ref_extra]):
        np.testing.assert_allclose(a.detach().numpy(), b.detach().numpy(), rtol=1e-5, atol=1e-7)


def test_unused_parameter_is_skipped_like_torch_optim(double):
    net, ref = _net(3), _net(3)
    arena = ParamArena(net.parameters(), device="cpu")
    opt = fused_optim.create_fused_optimizer(arena, OptimOpts(algo=OptAlgorithm.SGD, lr=0.1))
    pipe = grad_sync.GradBucketPipeline(arena, opt)
    ref_opt = torch.optim.SGD(ref.parameters(), lr=0.1, momentum=0.9, weight_decay=1e-5)
    x = torch.randn(8, 7)
    pipe.begin_step(); net[0](x).square().mean().backward(); pipe.finish_step()   # only layer 0
    ref[0](x).square().mean().backward(); ref_opt.step()
    for a, b in zip(net.parameters(), ref.parameters()):
        np.testing.assert_allclose(a.detach().numpy(), b.detach().numpy(), rtol=1e-6, atol=1e-8)


# ---- world_size 2 over gloo ---------------------------------------------------------------------

def _rank_main(rank, world, port, algo_value, clip, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    d = KernelDouble()
    fused_optim.KERNELS = d
    grad_sync.KERNELS = d
    net = _net(10 + rank)                      # deliberately different: broadcast must fix it
    arena = ParamArena(net.parameters(), device="cpu")
    opt = fused_optim.create_fused_optimizer(arena, OptimOpts(algo=OptAlgorithm(algo_value), lr=0.05))
    pipe = grad_sync.GradBucketPipeline(arena, opt, world_size=world, clip_norm=clip,
                                        bucket_cap_mb=0.0001, first_bucket_mb=0.00005)
    pipe.broadcast_parameters(src=0)
    assert len(pipe.buckets) > 1
    g = torch.Generator().manual_seed(99)
    for step in range(3):
        x = torch.randn(8 * world, 7, generator=g)
        pipe.begin_step()
        net(x[rank::world]).square().mean().backward()
        pipe.finish_step()
    torch.save([p.detach().clone() for p in net.parameters()], os.path.join(out_dir, f"r{rank}.pt"))
    dist.destroy_process_group()


@pytest.mark.parametrize("algo,clip", [("sgd", 0.0), ("adam", 0.0), ("sgd", 0.01)])
def test_two_rank_pipeline_equals_global_batch_training(tmp_path, algo, clip):
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_rank_main, args=(world, port, algo, clip, str(tmp_path)), nprocs=world, join=True)
    r0 = torch.load(tmp_path / "r0.pt")
    r1 = torch.load(tmp_path / "r1.pt")
    for a, b in zip(r0, r1):
        assert torch.equal(a, b)                               # replicas stay identical
    # single-process oracle on the concatenated batch (mean of per-rank means == global mean)
    from oracle.ref_loop import OptimSpec, make_optimizer
    ref = _net(10)
    ref_opt = make_optimizer(ref.parameters(), OptimSpec(algo=algo, lr=0.05))
    g = torch.Generator().manual_seed(99)
    for step in range(3):
        x = torch.randn(8 * world, 7, generator=g)
        ref_opt.zero_grad()
        ref(x).square().mean().backward()
        if clip:
            torch.nn.utils.clip_grad_norm_(ref.parameters(), clip)
        ref_opt.step()
    for a, b in zip(r0, ref.parameters()):
        np.testing.assert_allclose(a.numpy(), b.detach().numpy(), rtol=2e-5, atol=1e-7)


# ------------------------------------------------------------------------------------------------
# host gather pool (native threads; no GPU involved)
# ------------------------------------------------------------------------------------------------

def test_host_gather_pool_matches_index_select_and_orders_tickets():
    from frl_b200 import _native
    pool = _native.HostGatherPool(3)
    assert pool.n_threads == 3
    src = torch.randn(5000, 1024)
    jobs = []
    for k in range(6):
        idx = torch.randperm(5000)[:700 + k].contiguous()
        dst = torch.zeros(800, 1024)
        jobs.append((pool.submit(src, idx, dst), idx, dst))
    assert [t for t, _, _ in jobs] == list(range(1, 7))
    pool.wait(jobs[-1][0])                    # FIFO: waiting for the last covers all
    for _, idx, dst in jobs:
        assert torch.equal(dst[:len(idx)], src[idx])
        assert torch.count_nonzero(dst[len(idx):]) == 0
    # narrow rows (labels), repeated indices, unaligned row sizes
    lab = torch.arange(1000, dtype=torch.int64)
    out = torch.zeros(6, dtype=torch.int64)
    pool.wait(pool.submit(lab, torch.tensor([5, 999, 0, 3, 3, 3]), out))
    assert out.tolist() == [5, 999, 0, 3, 3, 3]
    odd = torch.randn(100, 7)
    out = torch.zeros(9, 7)
    idx = torch.randint(0, 100, (9,))
    pool.wait(pool.submit(odd, idx, out))
    assert torch.equal(out, odd[idx])
    # empty job completes; a bad index rejects the whole job and nothing is written
    pool.wait(pool.submit(lab, torch.zeros(0, dtype=torch.int64), out))
    keep = out.clone()
    with pytest.raises(_native.NativeLibraryError):
        pool.submit(odd, torch.tensor([1, 100]), out)
    with pytest.raises(_native.NativeLibraryError):
        pool.submit(odd, torch.tensor([-1]), out)
    assert torch.equal(out, keep)
    with pytest.raises(_native.NativeLibraryError):
        pool.wait(10_000)                     # never issued
    pool.close()


# ------------------------------------------------------------------------------------------------
# index stream of the batched input path: same batches, same global-RNG consumption as the
# reference's enumerate(DataLoader(shuffle=True)) / list(iter(sampler))
# ------------------------------------------------------------------------------------------------

@pytest.mark.parametrize("n,batch", [(1000, 64), (64, 64), (65, 64), (7, 3), (1, 4)])
def test_fast_index_stream_equals_dataloader_stream_and_rng_draws(n, batch):
    import torch.utils.data as tud
    from frl_b200.device_loader import DeviceBatchLoader, _IndexOnly, _collate_indices
    torch.manual_seed(5)
    ld = tud.DataLoader(_IndexOnly(n), batch_size=batch, shuffle=True, num_workers=0,
                        collate_fn=_collate_indices)
    want = [b.clone() for b in ld]
    rng_after = torch.get_rng_state()

    class Shell:                       # the three attributes _index_batches reads
        sampler, batch_size, _index_loader = ld.sampler, batch, ld

    torch.manual_seed(5)
    got = list(DeviceBatchLoader._index_batches(Shell))
    assert len(got) == len(want) and all(torch.equal(a, b) for a, b in zip(got, want))
    assert torch.equal(torch.get_rng_state(), rng_after)


def test_planned_order_for_the_null_accessor_only_draws_the_seed():
    import torch.utils.data as tud
    from frl_b200.solver_worker import _planned_order
    from frl_b200.storage_layers.dataset import NullAccessor
    sampler = tud.RandomSampler(range(1234))
    torch.manual_seed(9)
    full = list(iter(sampler))
    rng_after = torch.get_rng_state()
    torch.manual_seed(9)
    assert _planned_order(sampler, NullAccessor()) == []
    assert torch.equal(torch.get_rng_state(), rng_after)
    torch.manual_seed(9)
    assert _planned_order(sampler, object()) == full           # a real cache accessor gets the order
    seq = tud.SequentialSampler(range(5))
    assert _planned_order(seq, NullAccessor()) == [0, 1, 2, 3, 4]


def test_host_pool_bf16_wire_conversion_is_bit_identical_to_torch():
    from frl_b200 import _native
    pool = _native.HostGatherPool(2)
    rs = np.random.RandomState(0)
    vals = (rs.randn(300, 1037) * 10.0 ** rs.randint(-30, 30, size=(300, 1))).astype(np.float32)
    src = torch.from_numpy(vals)
    # exact ties, subnormals, infinities, NaNs, the largest finite value, signed zeros
    special = torch.tensor([1.00390625, 1.01171875, -1.00390625, 3.3895313892515355e38, 1e-40, -1e-45,
                            float("inf"), float("-inf"), float("nan"), 0.0, -0.0, 65504.0, 1.0 + 2 ** -8,
                            1.0 + 2 ** -8 + 2 ** -20], dtype=torch.float32)
    src[0, :len(special)] = special
    nan_payload = torch.tensor([0x7f800001, 0xffc12345], dtype=torch.int64).to(torch.int32).view(torch.float32)
    src[1, :2] = nan_payload
    idx = torch.cat([torch.tensor([0, 1]), torch.randint(0, 300, (70,))])
    out = torch.zeros(len(idx), 1037, dtype=torch.bfloat16)
    pool.wait(pool.submit_f32_to_bf16(src, idx, out))
    want = src[idx].to(torch.bfloat16)
    got_bits, want_bits = out.view(torch.int16), want.view(torch.int16)
    nan = torch.isnan(want)
    assert torch.equal(got_bits[~nan], want_bits[~nan])
    assert torch.isnan(out[nan]).all() and nan.sum() >= 3
    # unaligned destination rows (odd row length) and a row shorter than one vector
    small = torch.randn(50, 5)
    out2 = torch.zeros(9, 5, dtype=torch.bfloat16)
    i2 = torch.randint(0, 50, (9,))
    pool.wait(pool.submit_f32_to_bf16(small, i2, out2))
    assert torch.equal(out2, small[i2].to(torch.bfloat16))
    with pytest.raises(_native.NativeLibraryError):
        pool.submit_f32_to_bf16(small, torch.tensor([50]), out2)
    pool.close()


# ------------------------------------------------------------------------------------------------
# hang guard (reference tests/test_watchdog_timer.py restated for the one-thread-per-loop guard)
# ------------------------------------------------------------------------------------------------

def test_watchdog_completes_quietly_and_stops_its_thread():
    import threading
    import time
    from frl_b200.watchdog import StepWatchdog
    before = threading.active_count()
    with StepWatchdog(1_000_000) as dog:                    # reference: test_timer_completes
        assert dog._thread is not None and dog._thread.is_alive()
        for _ in range(5):
            dog.kick()
    dog._thread.join(timeout=3)
    assert not dog._thread.is_alive() and not dog.fired
    assert threading.active_count() <= before


def test_watchdog_expires_with_timeout_error_in_the_guarded_thread():
    import time
    from frl_b200.watchdog import StepWatchdog
    with pytest.raises(TimeoutError):                       # reference: test_timer_expires_with_exception
        with StepWatchdog(0):
            time.sleep(1)
            for _ in range(100):                            # async exceptions land between bytecodes
                time.sleep(0.01)


def test_watchdog_kicks_postpone_the_deadline_and_a_stall_fires_it():
    import time
    from frl_b200.watchdog import StepWatchdog
    with StepWatchdog(300) as dog:
        for _ in range(8):                                  # 0.8 s of work, never 300 ms without a kick
            time.sleep(0.1)
            dog.kick()
        assert not dog.fired
    with pytest.raises(TimeoutError):
        with StepWatchdog(200) as dog:
            dog.kick()
            for _ in range(300):                            # a "hung minibatch": no kick for > 200 ms
                time.sleep(0.01)
    assert dog.fired


# ---- a Linear applied twice per forward (weight reuse) under the eager per-bucket pipeline --------

class _Reuse(nn.Module):
    """`shared` runs twice per forward and `tied` borrows `first`'s weight: both parameters get
    two gradient contributions per backward."""

    def __init__(self, seed):
        super().__init__()
        torch.manual_seed(seed)
        self.first = nn.Linear(7, 7)
        self.shared = nn.Sequential(nn.Linear(7, 7), nn.ReLU())
        self.tied = nn.Linear(7, 7)
        self.tied.weight = self.first.weight
        self.head = nn.Linear(7, 2)

    def forward(self, x):
        h = torch.relu(self.first(x))
        h = self.shared(self.shared(h))
        return self.head(torch.relu(self.tied(h)))


def _reuse_rank_main(rank, world, port, patched, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from frl_b200 import arena_linear
    d = KernelDouble()
    fused_optim.KERNELS = d
    grad_sync.KERNELS = d
    arena_linear.KERNELS = d
    net = _Reuse(3)
    arena = ParamArena(net.parameters(), device="cpu")
    opt = fused_optim.create_fused_optimizer(arena, OptimOpts(algo=OptAlgorithm.SGD, lr=0.05))
    pipe = grad_sync.GradBucketPipeline(arena, opt, world_size=world, bucket_cap_mb=0.0001,
                                        first_bucket_mb=0.00005, eager_update=True)
    assert pipe.eager and len(pipe.buckets) > 2
    if patched:
        assert pipe.patch_linears(net) == 4
        assert sum(s.relu is not None for s in pipe.linear_sites) == 1
    pipe.broadcast_parameters(src=0)
    g = torch.Generator().manual_seed(5)
    net.train()
    for step in range(3):
        x = torch.randn(8 * world, 7, generator=g)
        if step == 1:                       # a forward that is never back-propagated (eval split)
            net.eval()
            net(x[rank::world])
            net.train()
        out = net(x[rank::world])
        pipe.begin_step()
        out.square().mean().backward()
        pipe.finish_step()
    torch.save([p.detach().clone() for p in net.parameters()], os.path.join(out_dir, f"r{rank}.pt"))
    dist.destroy_process_group()


@pytest.mark.parametrize("patched", [True, False])
def test_reused_and_tied_linears_wait_for_their_last_gradient(tmp_path, patched):
    """with gradients written from inside the layer's backward, a bucket must
    not be reduced/updated after the FIRST of several backward passes through a reused Linear."""
    world = 2
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_reuse_rank_main, args=(world, port, patched, str(tmp_path)), nprocs=world, join=True)
    r0, r1 = torch.load(tmp_path / "r0.pt"), torch.load(tmp_path / "r1.pt")
    for a, b in zip(r0, r1):
        assert torch.equal(a, b)
    ref = _Reuse(3)
    ref_opt = torch.optim.SGD(ref.parameters(), lr=0.05, momentum=0.9, weight_decay=1e-5)
    g = torch.Generator().manual_seed(5)
    for step in range(3):
        x = torch.randn(8 * world, 7, generator=g)
        ref_opt.zero_grad()
        ref(x).square().mean().backward()
        ref_opt.step()
    for a, b in zip(r0, ref.parameters()):
        np.testing.assert_allclose(a.numpy(), b.detach().numpy(), rtol=2e-5, atol=2e-7)


# ---- tail split: the weight whose dW ends backward is exchanged in two row blocks --------------------

class _TailNet(nn.Module):
    def __init__(self, seed, twice):
        super().__init__()
        torch.manual_seed(seed)
        self.first = nn.Sequential(nn.Linear(8, 16), nn.ReLU())
        self.mid = nn.Linear(16, 8)
        self.head = nn.Linear(8, 2)
        self.twice = twice

    def forward(self, x):
        h = self.first(x)
        if self.twice:                       # second application of `first`: no early hand-over
            h = self.first(torch.relu(self.mid(h)))
        return self.head(torch.relu(self.mid(h)))


def _tail_rank_main(rank, world, port, mode, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["FRL_B200_TAIL_SPLIT_MIN_BYTES"] = "0"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from frl_b200 import arena_linear
    d = KernelDouble()
    fused_optim.KERNELS = d
    grad_sync.KERNELS = d
    arena_linear.KERNELS = d
    net = _TailNet(4, twice=(mode == "twice"))
    arena = ParamArena(net.parameters(), device="cpu")
    opt = fused_optim.create_fused_optimizer(arena, OptimOpts(algo=OptAlgorithm.SGD, lr=0.05))
    pipe = grad_sync.GradBucketPipeline(arena, opt, world_size=world, bucket_cap_mb=0.0001,
                                        first_bucket_mb=0.00005, eager_update=True)
    first_w = arena.slots[0]
    assert tuple(first_w.shape) == (16, 8) and pipe.row_split(first_w) == 8
    a, b = pipe.buckets[-2], pipe.buckets[-1]
    assert (a.lo, a.hi) == (0, 64) and b.lo == 64 and a.slots == [first_w] and first_w in b.slots
    if mode != "hooks":
        pipe.patch_linears(net)
    launches = []
    inner = pipe._launch_bucket
    pipe._launch_bucket = lambda bk: (launches.append((bk.lo, bk.hi, pipe._ready)), inner(bk))[1]
    pipe.broadcast_parameters(src=0)
    g = torch.Generator().manual_seed(5)
    net.train()
    for step in range(3):
        x = torch.randn(8 * world, 8, generator=g)
        out = net(x[rank::world])
        pipe.begin_step()
        del launches[:]
        out.square().mean().backward()
        pipe.finish_step()
        order = [l[:2] for l in launches]
        ia, ib = order.index((0, 64)), order.index((64, b.hi))
        assert ia < ib and len(order) == len(pipe.buckets)
        early = launches[ia][2] < launches[ib][2]        # first half went before the slot was ready
        assert early == (mode == "once"), (mode, launches)
    torch.save([p.detach().clone() for p in net.parameters()], os.path.join(out_dir, f"r{rank}.pt"))
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["once", "twice", "hooks"])
def test_tail_split_hands_over_the_first_row_block_early(tmp_path, mode):
    world = 2
    port = 33500 + (os.getpid() % 2000)
    mp.spawn(_tail_rank_main, args=(world, port, mode, str(tmp_path)), nprocs=world, join=True)
    r0, r1 = torch.load(tmp_path / "r0.pt"), torch.load(tmp_path / "r1.pt")
    for a, b in zip(r0, r1):
        assert torch.equal(a, b)
    ref = _TailNet(4, twice=(mode == "twice"))
    ref_opt = torch.optim.SGD(ref.parameters(), lr=0.05, momentum=0.9, weight_decay=1e-5)
    g = torch.Generator().manual_seed(5)
    for step in range(3):
        x = torch.randn(8 * world, 8, generator=g)
        ref_opt.zero_grad()
        ref(x).square().mean().backward()
        ref_opt.step()
    for a, b in zip(r0, ref.parameters()):
        np.testing.assert_allclose(a.numpy(), b.detach().numpy(), rtol=2e-5, atol=2e-7)


# ---- round-2 host logic: arena layout groups, bucket merging, quiet randperm, multi-head unit ------

def test_adjacent_groups_keep_parameter_positions_and_sorted_offsets():
    torch.manual_seed(0)
    trunk = nn.Linear(16, 24)
    heads = [nn.Linear(24, 8), nn.Linear(24, 16), nn.Linear(24, 5)]
    params = list(trunk.parameters()) + [p for h in heads for p in h.parameters()]
    groups = [[h.weight for h in heads], [h.bias for h in heads]]
    arena = ParamArena(params, device="cpu", adjacent=groups)
    assert len(arena.adjacent_groups) == 2
    offs = [s.offset for s in arena.slots]
    assert offs == sorted(offs)                                  # slot list stays in layout order
    by_param = {id(s.param): s for s in arena.slots}
    for s in arena.slots:                                        # optimizer position = parameter order
        assert params[s.index] is s.param
    w = [by_param[id(h.weight)] for h in heads]
    assert w[0].end == w[1].offset and w[1].end == w[2].offset   # back to back, no padding
    b = [by_param[id(h.bias)] for h in heads]
    assert b[0].end == b[1].offset and b[1].end == b[2].offset
    for p, before in zip(params, [p.detach().clone() for p in params]):
        assert torch.equal(p, before)
    # a group whose inner members are not multiples of 8 elements is ignored, not half-applied
    odd = [nn.Linear(3, 5), nn.Linear(3, 7)]
    plain = ParamArena([p for m in odd for p in m.parameters()], device="cpu",
                       adjacent=[[m.weight for m in odd]])
    assert plain.adjacent_groups == [] and [s.index for s in plain.slots] == [0, 1, 2, 3]


def test_lone_bias_bucket_absorbs_its_weight():
    """6 -> 4 buckets on the MLP layout: a tiny open bucket takes the next oversize tensor in."""
    layers = []
    for _ in range(3):
        layers += [nn.Linear(4096, 4096), nn.ReLU()]
    net = nn.Sequential(*layers, nn.Linear(4096, 1000))
    arena = ParamArena(net.parameters(), device="cpu", precision=Precision.BF16)
    buckets = arena.buckets(24 << 20)
    assert len(buckets) == 4
    assert buckets[0][1] == arena.numel and buckets[-1][0] == 0
    assert all(a[0] == b[1] for a, b in zip(buckets[:-1], buckets[1:]))
    slot_of = {s.offset: s for s in arena.slots}
    for lo, hi in buckets:                                       # whole tensors only
        assert lo in slot_of and any(s.end == hi or arena.numel == hi for s in arena.slots)


def _random_net(rng):
    dims = [int(rng.choice([8, 16, 24, 40, 64])) for _ in range(int(rng.integers(2, 6)))]
    layers = []
    for a, b in zip(dims[:-1], dims[1:]):
        layers += [nn.Linear(a, b, bias=bool(rng.integers(0, 2))), nn.ReLU()]
    return nn.Sequential(*layers[:-1])


@pytest.mark.parametrize("seed", range(12))
@pytest.mark.parametrize("world", [2, 3, 8])
def test_bucket_and_shard_partition_invariants(double, monkeypatch, seed, world):
    """Whatever the layer sizes, the cap and the tail split do: launch ranges tile the arena
    exactly once in reverse order on 8-element boundaries, every parameter knows the buckets it
    overlaps, and the per-rank shards of every launch range tile that range (what makes
    ``sync_sharded_state`` and the K7 ownership rule well defined)."""
    rng = np.random.default_rng(seed)
    monkeypatch.setenv("FRL_B200_TAIL_SPLIT", "1")
    monkeypatch.setenv("FRL_B200_TAIL_SPLIT_MIN_BYTES", str(int(rng.choice([0, 256, 1 << 20]))))
    net = _random_net(rng)
    arena = ParamArena(net.parameters(), device="cpu",
                       precision=Precision.BF16 if seed % 2 else Precision.FP32)
    opt = fused_optim.create_fused_optimizer(arena, OptimOpts(algo=OptAlgorithm.SGD, lr=0.1))
    pipe = grad_sync.GradBucketPipeline(arena, opt, world_size=world, eager_update=True,
                                        bucket_cap_mb=float(rng.choice([1e-4, 1e-3, 1.0])),
                                        first_bucket_mb=None)
    try:
        ranges = [(b.lo, b.hi) for b in pipe.buckets]
        tiled = sorted(ranges)
        assert tiled[0][0] == 0 and tiled[-1][1] == arena.numel
        assert all(a[1] == b[0] for a, b in zip(tiled[:-1], tiled[1:]))            # gap-free, no overlap
        # launch order: last parameters first; the two row blocks of a split weight in row order
        whole = ranges if not pipe._row_split else ranges[:-2] + [(ranges[-2][0], ranges[-1][1])]
        assert whole == sorted(whole, reverse=True) and (not pipe._row_split or ranges[-2][1] == ranges[-1][0])
        assert all(lo % ALIGN_ELEMS == 0 and hi % ALIGN_ELEMS == 0 and lo < hi for lo, hi in ranges)
        for s in arena.slots:
            mine = pipe._buckets_of[id(s.param)]
            assert mine == [b for b in pipe.buckets if s.offset < b.hi and s.end > b.lo]
            rows = pipe.row_split(s)
            assert len(mine) == (2 if rows else 1)
            if rows:                                    # only the first tensor, cut at a row boundary
                assert s.offset == 0 and mine[0].hi == mine[1].lo == rows * s.shape[1]
        assert sum(1 for s in arena.slots if pipe.row_split(s)) <= 1

        class _Link:                                    # what shard_of reads of an NvlsLink
            pass
        for lo, hi in ranges:
            cover = []
            for r in range(world):
                opt.nvls = _Link()
                opt.nvls.rank, opt.nvls.world = r, world
                a, z = opt.shard_of(lo, hi)
                assert lo <= a <= z <= hi and (a - lo) % ALIGN_ELEMS == 0
                cover.append((a, z))
            assert cover[0][0] == lo and max(z for _, z in cover) == hi
            assert all(x[1] == y[0] or y[0] == y[1] for x, y in zip(cover[:-1], cover[1:]))
    finally:
        opt.nvls = None
        pipe.remove_hooks()


def test_quiet_randperm_is_torch_randperm():
    from frl_b200.device_loader import randperm_quiet
    for n in (1, 7, 1000, 100_000):
        g1, g2 = torch.Generator().manual_seed(n), torch.Generator().manual_seed(n)
        threads = torch.get_num_threads()
        assert torch.equal(torch.randperm(n, generator=g1), randperm_quiet(n, g2))
        assert torch.equal(g1.get_state(), g2.get_state()) and torch.get_num_threads() == threads


def test_multihead_unit_matches_per_head_autograd(double):
    """The task heads as one backward unit (one dX / one dW over adjacent arena weights) against
    stock autograd + torch SGD, three heads with a bias size that is not a multiple of 8."""
    from frl_b200 import arena_linear
    from frl_b200.model import ListSelect, MultiTaskModel
    arena_linear.KERNELS, saved = double, arena_linear.KERNELS
    try:
        def build():
            torch.manual_seed(1)
            trunk = nn.Sequential(ListSelect(sel_index=0, num_elements=1), nn.Linear(16, 24), nn.ReLU())
            return MultiTaskModel(trunk, [nn.Linear(24, 8), nn.Linear(24, 16), nn.Linear(24, 5)])
        net, ref = build(), build()
        arena = ParamArena(net.parameters(), device="cpu", adjacent=arena_linear.head_layout_groups(net))
        opt = fused_optim.create_fused_optimizer(arena, OptimOpts(algo=OptAlgorithm.SGD, lr=0.05))
        pipe = grad_sync.GradBucketPipeline(arena, opt, world_size=1, eager_update=False)
        assert pipe.patch_linears(net) == 4
        assert sum(s.multihead is not None for s in pipe.linear_sites) == 1 and "forward" in net.__dict__
        ref_opt = torch.optim.SGD(ref.parameters(), lr=0.05, momentum=0.9, weight_decay=1e-5)
        net.train(), ref.train()
        g = torch.Generator().manual_seed(3)
        for _ in range(3):
            x = torch.randn(12, 16, generator=g)
            outs = net([x])
            pipe.begin_step()
            sum(o.square().mean() for o in outs).backward()
            pipe.finish_step()
            ref_opt.zero_grad()
            sum(o.square().mean() for o in ref([x])).backward()
            ref_opt.step()
        for a, b in zip(net.parameters(), ref.parameters()):
            np.testing.assert_allclose(a.detach().numpy(), b.detach().numpy(), rtol=1e-5, atol=1e-7)
        pipe.unpatch_linears()
        assert "forward" not in net.__dict__                        # pickles as a plain module
        with torch.no_grad():
            assert len(net([torch.randn(2, 16)])) == 3
    finally:
        arena_linear.KERNELS = saved


def test_metric_hooks_run_synchronously_unless_the_problem_opts_in(monkeypatch):
    """Round-1 ADVICE (medium): host-returning metric hooks are folded on the training thread,
    as the reference does; the worker thread is opt-in."""
    import frl_b200.solver_worker as sw
    from frl_b200.problem import Ordering

    class P:
        def get_rankable_metric(self):
            return "m", Ordering.DESC

    monkeypatch.delenv("FRL_B200_ASYNC_METRICS", raising=False)
    s = sw.SamplerState(P(), 10, 10, torch.device("cpu"), 2)
    assert s._runner is None
