"""The plugin API keeps the reference's names, signatures, defaults and error behaviour
(SURVEY §8b); includes the reference's own unit tests for this path, restated."""
import inspect
import json
import os

import pytest
import torch

import frl_b200
from frl_b200 import (criteria, local_solver, lr_scheduler, model, multitask_problem, problem,
                      sampler, solver, solver_worker, task, transform, types)


def test_run_opts_and_optim_opts_fields_and_defaults():
    assert types.OptimOpts._fields == ("algo", "lr", "lr_scheduler", "weightDecay", "momentum",
                                       "epsilon", "amsgrad", "gradientClip")
    o = types.OptimOpts(algo=types.OptAlgorithm.SGD)
    assert (o.lr, o.weightDecay, o.momentum, o.epsilon, o.amsgrad, o.gradientClip) == (
        0.001, 0.00001, 0.9, 1e-8, False, 0.0)
    assert types.RunOpts._fields == (
        "optim", "batchSize", "cpuonly", "nEpochs", "maxEpochImages", "numThreads", "numIOThreads",
        "metricAmortizationSchedule", "initialModelPath", "mode", "numVisualizedSamples",
        "singleThreaded", "outputTTL", "lossLoggingFreq", "debugGrad", "shuffleType",
        "minibatchTimeoutMs")
    r = types.RunOpts(optim=o, batchSize=8)
    assert (r.nEpochs, r.numThreads, r.metricAmortizationSchedule, r.mode, r.shuffleType,
            r.minibatchTimeoutMs) == (75, 4, 10, types.Mode.TRAIN, types.ShuffleType.RANDPERM,
                                      3600000)
    assert [e.value for e in types.Split] == ["training", "testing", "heldOut"]
    assert [e.value for e in types.OptAlgorithm] == ["rmsprop", "sgd", "adam"]
    b = types.RunOptsBase(o, 4, nEpochs=3)
    assert b.batchSize == 4 and b.nEpochs == 3 and b.optim is o


def test_entry_point_signatures():
    sig = inspect.signature(solver.Solver.solve)
    names = list(sig.parameters)
    assert names[:2] == ["run_opts", "problem"]
    for kw in ("group_name", "init_method", "node_idx", "node_count", "memory_quota"):
        assert sig.parameters[kw].kind == inspect.Parameter.KEYWORD_ONLY
    assert sig.parameters["node_idx"].default == 0 and sig.parameters["node_count"].default == 1
    lsig = inspect.signature(local_solver.LocalSolver.solve)
    assert list(lsig.parameters)[:3] == ["run_opts", "problem", "save_notebook"]
    assert solver.PerformanceSummary._fields == ("epoch", "performance", "save_dir")
    assert solver_worker.FractionalPerformanceSummary._fields == (
        "epoch", "modelBuffer", "optimizerStateBuffer", "performance")
    wsig = inspect.signature(solver_worker.SolverWorker.train)
    assert list(wsig.parameters)[1:] == ["problem", "startEpoch", "nEpochs", "batchSize", "scheduler"]


def test_abstract_contracts():
    for name in ("datasets", "save_dir", "anno_param", "get_model", "get_criterion",
                 "refine_batch_meta", "compute_batch_metrics", "get_rankable_metric",
                 "summarize_epoch_samples", "summarize_epoch_metrics"):
        assert name in problem.Problem.__abstractmethods__
    assert problem.Problem.get_solver_buck_target() is None
    assert "get_model_base" in multitask_problem.MultiTaskProblem.__abstractmethods__
    for name in ("network_head", "criterion", "criterion_weight", "get_target",
                 "compute_batch_metrics", "rankable_metrics", "summarize_epoch_metrics",
                 "summarize_epoch_samples"):
        assert hasattr(task.Task, name)
    assert transform.Sample._fields == ("data", "target")
    assert [o.value for o in problem.Ordering] == ["asc", "desc"]


def test_solver_refuses_cpu():
    o = types.OptimOpts(algo=types.OptAlgorithm.SGD)
    r = types.RunOpts(optim=o, batchSize=8, cpuonly=True)
    with pytest.raises(RuntimeError, match="no CPU path"):
        next(solver.Solver.solve(r, problem=None, group_name=None, init_method="file:///tmp/x"))


def test_unknown_scheduler_and_optimizer_raise_value_error():
    class FakeAlgo:
        pass
    o = types.OptimOpts(algo=types.OptAlgorithm.SGD,
                        lr_scheduler=types.LRSchedulerOpts(algo=FakeAlgo()))
    r = types.RunOpts(optim=o, batchSize=1)
    opt = torch.optim.SGD([torch.nn.Parameter(torch.zeros(1))], lr=0.1)
    with pytest.raises(ValueError):
        solver.create_lr_scheduler(r, opt)


# ---- the reference's own tests for this path --------------------------------------------------

def test_create_lr_scheduler_T47589710():
    """reference tests/test_solver.py:17-34 — resume at epoch 60 of 75 must give lr/10."""
    lr = 0.01
    optim_opts = types.OptimOpts(lr=lr, algo=types.OptAlgorithm.ADAM)
    run_opts = types.RunOpts(nEpochs=75, mode=types.Mode.TRAIN, batchSize=16, optim=optim_opts)
    optimizer = torch.optim.Adam({torch.Tensor()}, lr=lr, weight_decay=0.0001, eps=1e-8,
                                 amsgrad=False)
    optimizer.param_groups[0]["initial_lr"] = lr
    solver.create_lr_scheduler(run_opts, optimizer, 60)
    assert optimizer.param_groups[0]["lr"] == 0.001


@pytest.mark.parametrize("n,nodes", [(12, 4), (11, 4), (12, 1), (11, 1)])
def test_per_node_randperm_covers_range(n, nodes):
    """reference tests/test_sampler.py:15-41."""
    g = torch.Generator()
    shuffle = []
    for i in range(nodes):
        shuffle += sampler.per_node_randperm(max=n, node_idx=i, node_count=nodes, generator=g)
    if n % nodes:
        shuffle = shuffle[:-1]          # the last element is a recycled pad
    assert sorted(shuffle) == list(range(n))


# ---- bit-exact index streams and LR tables vs the fixtures recorded from the reference --------

def test_scaffold_sampler_matches_reference_lists(golden_dir, monkeypatch):
    import re
    import torch.distributed as dist
    table = json.load(open(os.path.join(golden_dir, "samplers.json")))
    for key, per_rank in table.items():
        m = re.fullmatch(r"(\w+)_n(\d+)_w(\d+)_nodes(\d+)_e(\d+)", key)
        kind = m.group(1)
        n, w, nodes, e = (int(m.group(i)) for i in (2, 3, 4, 5))
        for rank, expect in enumerate(per_rank):
            monkeypatch.setattr(dist, "get_world_size", lambda *a, **k: w)
            monkeypatch.setattr(dist, "get_rank", lambda *a, **k: rank)
            s = sampler.ScaffoldSampler(list(range(n)), shuffle_type=types.ShuffleType(kind),
                                        node_idx=rank // (w // nodes), node_count=nodes)
            s.set_epoch(e)
            assert list(iter(s)) == expect, key


def test_lr_schedules_match_reference_tables(golden_dir):
    table = json.load(open(os.path.join(golden_dir, "lr_schedules.json")))
    for key, lrs in table.items():
        if key.startswith("kat_"):
            continue
        sched, n = key.split("_n")
        n = int(n)
        opt = torch.optim.SGD([torch.nn.Parameter(torch.zeros(1))], lr=0.1)
        ro = types.RunOpts(optim=types.OptimOpts(
            algo=types.OptAlgorithm.SGD, lr=0.1,
            lr_scheduler=types.LRSchedulerOpts(algo=types.LRSchedulerAlgorithm(sched))),
            batchSize=1, nEpochs=n)
        sch = solver.create_lr_scheduler(ro, opt)
        got = []
        for _ in range(n):
            got.append(opt.param_groups[0]["lr"])
            opt.step()
            sch.step()
        assert got == lrs, key          # same closed form, same float ops: exact


def test_reference_alias_makes_reference_imports_resolve():
    import sys
    saved = {k: v for k, v in sys.modules.items() if k == "frldistml" or k.startswith("frldistml.")}
    for k in saved:
        del sys.modules[k]
    try:
        frl_b200.install_reference_alias()
        from frldistml.scaffold.sampler import per_node_randperm   # reference tests/test_sampler.py:11
        from frldistml.scaffold.types import RunOpts
        assert per_node_randperm is sampler.per_node_randperm and RunOpts is types.RunOpts
    finally:       # the oracle shim imports the REAL reference under the same name: leave no alias
        for k in [k for k in sys.modules if k == "frldistml" or k.startswith("frldistml.")]:
            del sys.modules[k]
        sys.modules.update(saved)


def test_model_helpers():
    base = torch.nn.Sequential(model.ListSelect(sel_index=0, num_elements=1), torch.nn.Linear(4, 3))
    m = model.MultiTaskModel(base, [torch.nn.Linear(3, 2), torch.nn.Linear(3, 1)])
    out = m([torch.randn(5, 4)])
    assert [tuple(o.shape) for o in out] == [(5, 2), (5, 1)]
    assert m.final_shared_params(out) is base[1].weight or m.final_shared_params(out) is base[1].bias
    assert model.View((2, 6))(torch.zeros(3, 4)).shape == (2, 6)
    assert torch.equal(model.MulConstant(2.0)(torch.ones(2)), torch.full((2,), 2.0))
    with pytest.raises(AssertionError):
        model.ListSelect(sel_index=0, num_elements=2)([torch.zeros(1)])


# ---- SamplerState against the fixtures recorded from the live reference class ---------------------




@pytest.mark.parametrize("config", ["err_MSE_DESC", "score_ASC", "score_DESC"])
def test_sampler_state_host_path_matches_reference_fixture(golden_dir, config):
    """tests/golden/sampler_state.json = what the unmodified reference's SamplerState produced for
    the shared scenario (oracle/make_sampler_state_golden.py): per-sample metrics, random picks
    and worst-k set of this repo's class must be the same (host arrays from the hook)."""
    import json
    import random
    import numpy as np
    import torch
    import frl_b200.solver_worker as sw
    from frl_b200.problem import Ordering
    from oracle import make_sampler_state_golden as gen
    want = json.load(open(os.path.join(golden_dir, "sampler_state.json")))[config]
    name, ordering = config.rsplit("_", 1)
    batches, total = gen.scenario()
    random.seed(gen.PY_SEED)
    mine = sw.SamplerState(gen.make_problem(Ordering, name, ordering), total, total, torch.device("cpu"),
                           gen.N_VIS)
    gen.drive(mine, batches)
    mine.finish()
    for k, v in want["metrics"].items():
        np.testing.assert_array_equal(np.asarray(mine.data_metric[k], dtype=np.float64), np.asarray(v))
    assert [int(s.meta["index"]) for s in mine.random_samples] == want["random_ids"]
    assert sorted(int(s.meta["index"]) for s in mine.worst_samples) == want["worst_ids"]


@pytest.mark.parametrize("config", ["err_MSE_DESC", "score_ASC", "score_DESC"])
def test_sampler_state_device_fold_logic_on_host_tensors(golden_dir, config):
    """The device-side fold (per-sample columns, random picks by position, running top-k merged per
    window, ONE packed read-back) is tensor logic that does not care where the tensors live: driven
    here with host tensors (the hook returns tensors, ``_dev_mode`` forced) against the same
    reference fixtures; ``tests/test_gpu_solver.py`` runs it on the device."""
    import json
    import random
    import numpy as np
    import torch
    import frl_b200.solver_worker as sw
    from frl_b200.problem import Ordering
    from oracle import make_sampler_state_golden as gen
    want = json.load(open(os.path.join(golden_dir, "sampler_state.json")))[config]
    name, ordering = config.rsplit("_", 1)
    batches, total = gen.scenario()
    random.seed(gen.PY_SEED)
    mine = sw.SamplerState(gen.make_problem(Ordering, name, ordering, as_numpy=False), total, total,
                           torch.device("cpu"), gen.N_VIS)
    mine._dev_mode = True
    gen.drive(mine, batches)
    assert not mine.random_samples and not mine.worst_samples and mine._dev_worst is not None
    mine.finish()
    for k, v in want["metrics"].items():
        np.testing.assert_allclose(np.asarray(mine.data_metric[k], dtype=np.float64), np.asarray(v),
                                   rtol=1e-6, atol=1e-7)
    assert [int(s.meta["index"]) for s in mine.random_samples] == want["random_ids"]
    assert sorted(int(s.meta["index"]) for s in mine.worst_samples) == want["worst_ids"]
    rows = {int(i): (b["data"][0][j], b["outputs"][0][j], b["targets"][1][0][j])
            for b in batches for j, i in enumerate(b["meta"]["index"])}
    for smp in mine.random_samples + mine.worst_samples:
        d, o, t = rows[int(smp.meta["index"])]
        assert torch.equal(smp.data[0], d) and torch.equal(smp.output[0], o) and torch.equal(smp.target[1][0], t)

