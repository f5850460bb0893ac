"""Per-bucket gradient pipeline: flatten -> all-reduce -> fused update.

Stands where the reference wraps the model in ``DistributedDataParallel`` (reference
solver.py:265-294) and later calls ``clip_grad_norm_`` / ``optimizer.step()`` (reference
solver_worker.py:585-592).  Differences that matter on B200:

* ``nn.Linear`` gradients are written straight into the flat ``grad`` arena by ``arena_linear``;
  every other gradient (convolutions, normalisation layers: cuDNN allocates its own output) is
  left where autograd put it and handed to the kernels through a segment table
  (``multi_tensor.GradSegTable``): on one GPU the fused update reads it in place
  (``frl_*_mt``: no flatten pass at all), with several GPUs ONE ``frl_flatten_grads`` launch per
  bucket gathers the bucket's stragglers into the slice NCCL / the NVLS kernel reduce **in place**
  — no per-tensor copy kernels, no copy-out;
* the 1/world mean and the clip coefficient are folded into the update kernel's gradient read;
* each bucket's all-reduce is issued on a side stream the moment its last gradient is ready and
  its fused update is chained right behind it, overlapping the rest of backward;
* with clipping enabled the updates wait for the global norm (two-phase tail), computed by one
  reduction kernel over the model range with no host sync.
"""
import os
from typing import Callable, Dict, List, Optional, Tuple

import torch
import torch.distributed as dist
import torch.nn as nn

from . import _native
from .arena import ParamArena
from .fused_optim import FusedArenaOptimizer

KERNELS = _native     # swapped by CPU tests of the host logic


class _TableSet:
    """Segment tables of one pipeline, built on first use: the whole arena (1-GPU tail update that
    reads gradients in place) and, per (bucket, set of straggler slots), the flatten tables.
    A CUDA-graph capture gets a set of its own: captured uploads read the set's pinned rows at
    every replay, so nothing else may ever rewrite them."""

    def __init__(self, arena) -> None:
        self.arena = arena
        self._whole = None
        self._flatten: Dict[Tuple, object] = {}

    def whole(self):
        if self._whole is None:
            from .multi_tensor import GradSegTable
            self._whole = GradSegTable(self.arena.slots, self.arena.device)
        return self._whole

    def flatten(self, key, slots):
        t = self._flatten.get(key)
        if t is None:
            from .multi_tensor import GradSegTable
            t = self._flatten[key] = GradSegTable(slots, self.arena.device)
        return t


class _Bucket:
    __slots__ = ("lo", "hi", "slots", "pending", "work", "launched", "replicated")

    def __init__(self, lo: int, hi: int):
        self.lo, self.hi = lo, hi
        self.slots = []
        self.pending = 0
        self.work = None
        self.launched = False
        self.replicated = False      # fused-NVLS runs: this bucket keeps the NCCL + K2 path


class GradBucketPipeline:
    def __init__(self, arena: ParamArena, optimizer: FusedArenaOptimizer, *,
                 process_group=None, world_size: int = 1, clip_norm: float = 0.0,
                 bucket_cap_mb: float = 25.0, first_bucket_mb: Optional[float] = 1.0,
                 eager_update: bool = True, nvls_link=None) -> None:
        self.arena = arena
        self.optimizer = optimizer
        self.pg = process_group
        self.world = world_size
        self.clip_norm = float(clip_norm or 0.0)
        self.grad_scale = 1.0 / world_size
        self.distributed = world_size > 1
        self.on_cuda = arena.device.type == "cuda"
        # eager: update a bucket on the side stream as soon as it is complete (and reduced) while
        # backward is still running.  Needs no global norm, so clipping turns it off.  On one GPU
        # the caller decides (measured on B200: the persistent cuBLAS GEMMs leave the update no
        # SMs to overlap on, so a single tail launch is faster and cheaper to issue).
        self.eager = eager_update and self.clip_norm == 0.0 and (self.distributed or self.on_cuda)

        cap = int(bucket_cap_mb * 1024 * 1024)
        first = int(first_bucket_mb * 1024 * 1024) if (first_bucket_mb and self.distributed) else None
        ranges = arena.buckets(cap, first) if (self.distributed or self.eager) else [(0, arena.numel)]
        use_nvls = nvls_link is not None and self.distributed and self.eager
        if use_nvls and arena.lp is not None and arena.model_end < arena.numel:
            # BF16 mode: criterion parameters read the fp32 master, which the fused NVLS step keeps
            # current only on the owning rank -> give them their own, replicated (NCCL) bucket
            cut = arena.model_end
            split = []
            for lo, hi in ranges:
                split += [(cut, hi), (lo, cut)] if lo < cut < hi else [(lo, hi)]
            ranges = split
        # Tail split.  The bucket that becomes ready LAST (the first layer's weight: its dW GEMM is
        # the final kernel of backward) is the only one whose exchange cannot hide behind backward
        # work.  When that bucket starts with a large 2-D weight, the weight's rows are cut in two
        # buckets: the layer computes dW as two row-block GEMMs and hands over the first half
        # (``rows_ready``) while the second is still running, so only half of the exchange is
        # exposed.  The cut is STATIC (decided here, not per step): with the fused NVLS step the
        # launch ranges define which rank owns which shard of the master weights and state.
        self._row_split: Dict[int, int] = {}            # slot.index -> rows in the first half
        min_bytes = int(os.environ.get("FRL_B200_TAIL_SPLIT_MIN_BYTES", str(4 << 20)))
        # measured (MLP, 24 MiB buckets): 2 GPUs 1.167 -> 1.141 ms/step; 8 GPUs 1.019 -> 1.032 (p50):
        # there a half-bucket exchange is mostly its two cross-GPU barriers, and one more launch
        # costs more than the exposed half saves -> on by default at world 2 only
        want_split = os.environ.get("FRL_B200_TAIL_SPLIT", "1" if world_size == 2 else "0") != "0"
        if self.distributed and self.eager and ranges and want_split:
            lo, hi = ranges[-1]
            esz = 2 if arena.grad_dtype == torch.bfloat16 else 4
            head = next((s for s in arena.slots if s.offset == lo), None)
            if (head is not None and len(head.shape) == 2 and head.end <= hi
                    and head.shape[0] >= 2 and (head.end - lo) * esz >= min_bytes
                    and 2 * (head.end - lo) >= hi - lo):
                rows = head.shape[0] // 2
                if (rows * head.shape[1]) % 8 == 0:       # launch ranges stay 8-element aligned
                    mid = lo + rows * head.shape[1]
                    ranges = ranges[:-1] + [(lo, mid), (mid, hi)]
                    self._row_split[head.index] = rows
        self.buckets: List[_Bucket] = [_Bucket(lo, hi) for lo, hi in ranges]
        if use_nvls and arena.lp is not None:
            for b in self.buckets:
                b.replicated = b.lo >= arena.model_end
        self._last_bucket = self.buckets[-1] if self.buckets else None      # last in launch order
        self._buckets_of: Dict[int, List[_Bucket]] = {}
        for s in arena.slots:
            for b in self.buckets:                   # launch order; a row-split slot is in two
                inside = (s.offset < b.hi and s.end > b.lo) if s.end > s.offset else (b.lo <= s.offset < b.hi)
                if inside:
                    b.slots.append(s)
                    self._buckets_of.setdefault(id(s.param), []).append(b)
        self._n_slots = len(arena.slots)
        self._ready = 0
        self._ready_ids = set()
        self._handles = []
        for s in arena.slots:
            self._handles.append(s.param.register_post_accumulate_grad_hook(self._make_hook(s)))

        # high priority: the (short) reduce/update kernels should not queue behind GEMM CTAs
        self.side_stream = torch.cuda.Stream(device=arena.device, priority=-1) if self.on_cuda else None
        self.clip_out = None
        self.clip_scratch = None
        if self.clip_norm > 0.0:
            self.clip_out = torch.zeros(3, dtype=torch.float32, device=arena.device)
            nbytes = KERNELS.reduce_scratch_bytes()
            self.clip_scratch = torch.zeros((nbytes + 3) // 4, dtype=torch.int32, device=arena.device)
        # fused NVLS step: all-reduce + update + weight broadcast are ONE kernel per bucket
        # (csrc/nvls.cu); needs per-bucket updates, so clipping keeps the NCCL + K2/K3 path
        self.nvls = nvls_link if use_nvls else None
        if self.nvls is not None:
            optimizer.nvls = self.nvls
        self._step_open = False
        self.step_id = 0
        self.linear_sites = []
        # arena_linear bookkeeping: per-parameter touch/application counters, the generation the
        # forward-application counts belong to (advanced when a step ends) and slots whose layer
        # ran a backward pass but expects more (module applied several times per forward)
        self.slot_states = {}
        self.forward_gen = 0
        self._deferred = {}
        # gradients autograd allocated itself, by slot index, kept alive until the kernels that
        # read them in place have been enqueued (multi-tensor path; CUDA only)
        self.mt_enabled = (self.on_cuda and hasattr(KERNELS, "flatten_grads")
                           and os.environ.get("FRL_B200_MT_GRADS", "1") != "0")
        self._ext: Dict[int, torch.Tensor] = {}
        self.tables = _TableSet(arena)
        self._keep_ext = False       # replay of a captured step: the capture owns the references
        # timing taps (bench): list of (start_event, end_event, lo, hi) for update launches
        self.record_update_events = False
        self.update_events: List[Tuple[torch.cuda.Event, torch.cuda.Event, int, int]] = []

    # -- step protocol ---------------------------------------------------------------------------
    def begin_step(self) -> None:
        """Call before ``backward()``."""
        self._ready = 0
        self._ready_ids.clear()
        self._deferred.clear()
        for b in self.buckets:
            b.pending = len(b.slots)
            b.work = None
            b.launched = False
        self.optimizer.begin_step()
        self.step_id += 1
        self._step_open = True

    @property
    def step_open(self) -> bool:
        return self._step_open

    def _make_hook(self, slot) -> Callable[[nn.Parameter], None]:
        from .multi_tensor import grad_usable_in_place

        def hook(param: nn.Parameter) -> None:
            g = param.grad
            if g is not None:
                dst = self.arena.grad_view(slot)
                if g.data_ptr() != dst.data_ptr():
                    if self.mt_enabled and self._step_open and grad_usable_in_place(g, slot):
                        self._ext[slot.index] = g     # read in place / gathered per bucket later
                    else:
                        dst.copy_(g)          # odd layouts, CPU tensors: flatten this one now
                param.grad = None             # next backward steals again instead of accumulating
            self.mark_ready(slot)
        return hook

    # -- multi-tensor plumbing ---------------------------------------------------------------------
    def _point(self, table) -> None:
        """This step's gradient locations into ``table`` (arena slice unless a straggler)."""
        g_arena = self.arena.grad
        esz = g_arena.element_size()
        base = g_arena.data_ptr()
        for s in table.slots:
            g = self._ext.get(s.index)
            if g is None:
                table.point(s, base + s.offset * esz, g_arena.dtype)
            else:
                table.point(s, g.data_ptr(), g.dtype)

    def _flatten_stragglers(self, slots, key, side: bool) -> None:
        """ONE launch: gather the listed slots' out-of-arena gradients into their arena slices
        (cast to the arena's gradient dtype) on the current stream."""
        ext = [s for s in slots if s.index in self._ext]
        if not ext:
            return
        table = self.tables.flatten((key, tuple(s.index for s in ext)), ext)
        self._point(table)
        table.upload()
        KERNELS.flatten_grads(table, self.arena.grad, scale=1.0)
        for s in ext:
            g = self._ext.pop(s.index) if not self._keep_ext else self._ext[s.index]
            if side:
                g.record_stream(torch.cuda.current_stream())

    def materialize_grads(self) -> None:
        """Debug/inspection aid: make ``arena.grad`` hold every gradient of the open step."""
        self._flatten_stragglers(self.arena.slots, "all", side=False)

    def detach_grad_refs(self):
        """Hand the straggler references and the table set to a captured graph (which replays
        kernels that read exactly these addresses); the pipeline continues with fresh ones."""
        refs, tables = self._ext, self.tables
        self._ext, self.tables = {}, _TableSet(self.arena)
        return refs, tables

    def mark_ready(self, slot) -> None:
        if not self._step_open or id(slot.param) in self._ready_ids:
            return
        self._ready_ids.add(id(slot.param))
        self._ready += 1
        for b in self._buckets_of[id(slot.param)]:
            if b.launched:                    # the first half of a row-split weight went ahead
                continue
            b.pending -= 1
            if b.pending == 0 and (self.distributed or self.eager):
                self._launch_bucket(b)

    def row_split(self, slot) -> int:
        """Rows of the slot's first half if its weight gradient is exchanged in two row blocks
        (tail split, see ``__init__``), else 0."""
        return self._row_split.get(slot.index, 0)

    def rows_ready(self, slot, rows: int) -> None:
        """Rows ``[0, rows)`` of the slot's gradient are final and the rest is still being
        computed: launch every bucket that lies inside them and waits for nothing else."""
        if not self._step_open:
            return
        done = slot.offset + rows * slot.shape[1]
        for b in self._buckets_of[id(slot.param)]:
            if not b.launched and b.hi <= done and b.pending == 1:
                b.pending = 0
                self._launch_bucket(b)

    def defer_ready(self, slot) -> None:
        """The slot's gradient was written but more contributions are expected in this backward;
        ``mark_ready`` follows from the last one, or from ``finish_step`` at the latest."""
        self._deferred[slot.index] = slot

    def _launch_bucket(self, b: _Bucket) -> None:
        """Bucket complete: (all-reduce it and) update it on the side stream, behind everything
        the compute stream has issued so far.  Runs in the autograd thread, so it is kept lean:
        no context managers, one event."""
        if self.on_cuda:
            cur = torch.cuda.current_stream()
            self.side_stream.wait_stream(cur)
            torch.cuda.set_stream(self.side_stream)
        try:
            if self._ext:
                self._flatten_stragglers(b.slots, id(b), side=self.on_cuda)
            if self.nvls is not None and not b.replicated:
                nv = self.nvls
                grid = nv.max_blocks
                if b is self._last_bucket:            # runs alone: backward has nothing left to issue
                    nv.max_blocks = nv.tail_blocks
                try:
                    self._update(b.lo, b.hi, None)    # K7: reduce + update + broadcast
                finally:
                    nv.max_blocks = grid
            elif self.distributed:
                b.work = dist.all_reduce(self.arena.grad[b.lo:b.hi], op=dist.ReduceOp.SUM,
                                         group=self.pg, async_op=True)
                if self.eager:
                    b.work.wait()             # stream-level wait, the host does not block
                    self._update(b.lo, b.hi, None)
            elif self.eager:
                self._update(b.lo, b.hi, None)
        finally:
            if self.on_cuda:
                torch.cuda.set_stream(cur)
        b.launched = True

    def _apply(self, lo: int, hi: int, coef) -> None:
        if self.nvls is not None and lo < self.arena.model_end or \
                (self.nvls is not None and self.arena.lp is None):
            self.optimizer.apply_range_nvls(lo, hi, grad_scale=self.grad_scale)
        else:
            self.optimizer.apply_range(lo, hi, grad_scale=self.grad_scale, clip_coef_dev=coef)

    def _update(self, lo: int, hi: int, coef) -> None:
        if self.record_update_events and self.on_cuda:
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            self._apply(lo, hi, coef)
            e1.record()
            self.update_events.append((e0, e1, lo, hi))
        else:
            self._apply(lo, hi, coef)

    def finish_step(self, defer_tail: bool = False) -> None:
        """Call after ``backward()`` returned: reduces/updates whatever is still outstanding and
        re-joins the side stream.  ``defer_tail=True`` (CUDA-graph capture) leaves the tail
        launches (norm + update of whatever was not updated eagerly) to ``run_tail()``."""
        if not self._step_open:
            raise RuntimeError("finish_step() without begin_step()")
        for slot in list(self._deferred.values()):    # backward() has returned: nothing is missing
            self.mark_ready(slot)
        self._deferred.clear()
        self.forward_gen += 1
        self._step_open = False
        self._tail_deferred = False
        if self._ready != self._n_slots:
            missing = [s.index for s in self.arena.slots if id(s.param) not in self._ready_ids]
            if self.distributed:
                # same contract as DDP with find_unused_parameters=False (the reference's setting)
                raise RuntimeError(
                    "Expected to have finished reduction for every parameter, but parameters at "
                    f"indices {missing} did not receive a gradient in this step")
            if self.eager and self.on_cuda:
                torch.cuda.current_stream().wait_stream(self.side_stream)
            self._flatten_stragglers(self.arena.slots, "all", side=False)
            self._finish_partial(missing)
            return
        if self.distributed or self.eager:
            if self.on_cuda:
                cur = torch.cuda.current_stream()
                if self.distributed and not self.eager:
                    torch.cuda.set_stream(self.side_stream)
                    for b in self.buckets:
                        b.work.wait()
                    torch.cuda.set_stream(cur)
                cur.wait_stream(self.side_stream)
            elif self.distributed and not self.eager:
                for b in self.buckets:
                    b.work.wait()
            if defer_tail:
                self._tail_deferred = True
                return
            if not self.eager:
                self._tail_update()
        else:
            if defer_tail:
                self._tail_deferred = True
                return
            self._tail_update()
        self._ext.clear()
        self.optimizer.end_step()

    @property
    def has_tail(self) -> bool:
        """True if a step ends with tail launches (not everything is updated eagerly)."""
        return not self.eager

    def run_tail(self, grad_refs=None, tables=None) -> None:
        """The deferred part of ``finish_step(defer_tail=True)``; also what a CUDA-graph replay
        of the captured step is followed by (then with the capture's gradient references and
        segment tables: the replayed backward wrote to exactly those addresses)."""
        if grad_refs is not None:
            mine = (self._ext, self.tables)
            self._ext, self.tables, self._keep_ext = grad_refs, tables, True
        try:
            if self.has_tail:
                self._tail_update()
        finally:
            if grad_refs is not None:
                (self._ext, self.tables), self._keep_ext = mine, False
        self.optimizer.end_step()

    def _tail_update(self) -> None:
        if self.clip_norm > 0.0:
            # the global norm needs every gradient first: gather the stragglers into the arena,
            # then K3 + K2 over the arena (the clip coefficient applies to model parameters only)
            self._flatten_stragglers(self.arena.slots, "all", side=False)
            n_model = self.arena.model_end
            KERNELS.grad_sumsq_clip(self.arena.grad[:n_model], n_model, pre_scale=self.grad_scale,
                                    max_norm=self.clip_norm, out3=self.clip_out,
                                    scratch=self.clip_scratch)
            self._update(0, self.arena.numel, self.clip_out[2:3])
            return
        if self._ext and not self.distributed:
            # one GPU: the update reads every gradient where it lies — no flatten pass
            table = self.tables.whole()
            self._point(table)
            table.upload()
            self._update_table(table)
            if not self._keep_ext:
                self._ext.clear()
            return
        self._update(0, self.arena.numel, None)

    def _update_table(self, table) -> None:
        if self.record_update_events and self.on_cuda:
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            self.optimizer.apply_table(table, grad_scale=self.grad_scale)
            e1.record()
            self.update_events.append((e0, e1, 0, self.arena.numel))
        else:
            self.optimizer.apply_table(table, grad_scale=self.grad_scale)

    def _finish_partial(self, missing) -> None:
        """world_size == 1 and some parameters got no gradient: torch.optim skips those (no
        weight decay, no momentum decay), so update only the contiguous runs that did."""
        skip = set(missing)
        done = [(b.lo, b.hi) for b in self.buckets if b.launched and self.eager]
        coef = None
        if self.clip_norm > 0.0:
            for s in self.arena.slots:
                if s.index in skip and s.is_model:
                    self.arena.grad[s.offset:s.end].zero_()
            n_model = self.arena.model_end
            KERNELS.grad_sumsq_clip(self.arena.grad[:n_model], n_model, pre_scale=self.grad_scale,
                                    max_norm=self.clip_norm, out3=self.clip_out,
                                    scratch=self.clip_scratch)
            coef = self.clip_out[2:3]
        run_lo = None
        prev_end = None
        for s in self.arena.slots:
            if s.index in skip or any(lo <= s.offset < hi for lo, hi in done):
                if run_lo is not None:
                    self._update(run_lo, prev_end, coef)
                    run_lo = None
                continue
            if run_lo is None:
                run_lo = s.offset
            prev_end = s.end
        if run_lo is not None:
            self._update(run_lo, prev_end, coef)
        self.optimizer.end_step()

    # -- one-time synchronisation ----------------------------------------------------------------
    def broadcast_parameters(self, src: int = 0) -> None:
        """Replicas start identical to rank ``src`` (DDP does this in its constructor)."""
        if not self.distributed:
            return
        dist.broadcast(self.arena.master, src=src, group=self.pg)
        self.arena.refresh_shadow()

    def sync_sharded_state(self) -> None:
        """Fused NVLS step only: every rank holds the current fp32 master / optimizer state of its
        own shards; before a checkpoint export make them whole everywhere (collective call)."""
        if self.nvls is None:
            return
        opt = self.optimizer
        vecs = list(opt._vec.values())
        if self.arena.lp is not None:
            vecs.append(self.arena.master)         # FP32 mode multicasts the master itself
        for vec in vecs:
            whole = torch.zeros_like(vec)
            for b in self.buckets:
                if b.replicated:
                    if self.nvls.rank == 0:
                        whole[b.lo:b.hi] = vec[b.lo:b.hi]
                    continue
                a, z = opt.shard_of(b.lo, b.hi)
                whole[a:z] = vec[a:z]
            dist.all_reduce(whole, op=dist.ReduceOp.SUM, group=self.pg)
            vec.copy_(whole)

    def remove_hooks(self) -> None:
        for h in self._handles:
            h.remove()
        self._handles = []
        self.unpatch_linears()

    # -- nn.Linear gradients written directly into the arena ---------------------------------------
    def patch_linears(self, model: nn.Module) -> int:
        from .arena_linear import patch_linears
        self.linear_sites = patch_linears(model, self)
        return len(self.linear_sites)

    def unpatch_linears(self) -> None:
        from .arena_linear import unpatch_linears
        unpatch_linears(self.linear_sites)

    def repatch_linears(self) -> None:
        from .arena_linear import repatch_linears
        repatch_linears(self.linear_sites)


class BufferBroadcaster:
    """Module buffers (BatchNorm statistics) follow rank 0 before every forward, as DDP's
    ``broadcast_buffers=True`` default does for the reference — but as ONE broadcast per dtype
    over a flat buffer arena the buffers are re-pointed into."""

    def __init__(self, module: nn.Module, *, process_group=None, world_size: int = 1) -> None:
        self.pg = process_group
        self.enabled = world_size > 1
        self.flat: List[torch.Tensor] = []
        if not self.enabled:
            return
        by_dtype = {}
        for buf in module.buffers():
            by_dtype.setdefault(buf.dtype, []).append(buf)
        for dtype, bufs in by_dtype.items():
            total = sum(b.numel() for b in bufs)
            if total == 0:
                continue
            flat = torch.empty(total, dtype=dtype, device=bufs[0].device)
            off = 0
            for b in bufs:
                n = b.numel()
                flat[off:off + n].copy_(b.reshape(-1))
                b.data = flat[off:off + n].view(b.shape)
                off += n
            self.flat.append(flat)

    def sync(self, src: int = 0) -> None:
        if self.enabled:
            for flat in self.flat:
                dist.broadcast(flat, src=src, group=self.pg)
