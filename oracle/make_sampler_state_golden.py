"""TEST INFRASTRUCTURE — golden fixtures for ``SamplerState`` (reference solver_worker.py:189-374)
from the LIVE, unmodified reference (build container only):

    python -m oracle.make_sampler_state_golden

For three (rankable metric, ordering) configurations the reference's own class folds five
synthetic minibatches (ragged last one, fold every second minibatch); recorded are the per-sample
metric arrays, the ids of its random picks and of its worst-k set.  ``scenario()`` is shared with
the tests so both sides see the very same tensors."""
import json
import os
import random
import sys
from typing import NamedTuple

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
OUT = os.path.join(REPO, "tests", "golden", "sampler_state.json")
CONFIGS = [("err_MSE", "DESC"), ("score", "ASC"), ("score", "DESC")]
SIZES = [16, 16, 16, 16, 9]
N_VIS = 6
PY_SEED = 5


class Meta(NamedTuple):
    index: object = None


def scenario():
    g = torch.Generator().manual_seed(11)
    batches, start = [], 0
    for n in SIZES:
        batches.append(dict(
            meta={"index": torch.arange(start, start + n)},
            data=[torch.randn(n, 5, generator=g)],
            outputs=[torch.randn(n, 4, generator=g), torch.randn(n, 3, generator=g)],
            targets=[(torch.randn(n, 4, generator=g),), (torch.randint(0, 3, (n,), generator=g),)]))
        start += n
    return batches, start


def metrics(output, target):
    """The per-sample metrics of the scenario's Problem as torch expressions (device-agnostic)."""
    err = ((output[0] - target[0][0]) ** 2).mean(1)
    score = output[1][:, 0] - target[1][0].float()           # signed, tie-free
    return {"err_MSE": err, "score": score}


def make_problem(Ordering, metric_name, ordering, as_numpy=True):
    class P:
        def refine_batch_meta(self, meta):
            return Meta(**meta)

        def compute_batch_metrics(self, meta, target, output, device):
            m = metrics(output, target)
            return {k: v.numpy() for k, v in m.items()} if as_numpy else m

        def get_rankable_metric(self):
            return metric_name, Ordering[ordering]
    return P()


def drive(state, batches):
    for k, b in enumerate(batches):
        if k % 2 == 0:                                # amortisation: fold every second minibatch
            state.compute_metrics()
        state.append_sample(b["meta"], b["data"], outputs=b["outputs"], targets=b["targets"])
    state.compute_metrics()


def main():
    from oracle.ref_shim import import_reference
    import_reference()
    import frldistml.scaffold.solver_worker as ref_sw
    from frldistml.scaffold.problem import Ordering
    batches, total = scenario()

    class FakeLoader:
        sampler = list(range(total))

    out = {}
    for name, ordering in CONFIGS:
        random.seed(PY_SEED)
        ref = ref_sw.SamplerState(make_problem(Ordering, name, ordering), FakeLoader, list(range(total)),
                                  torch.device("cpu"), N_VIS)
        drive(ref, batches)
        out["%s_%s" % (name, ordering)] = {
            "metrics": {k: np.asarray(v, dtype=np.float64).tolist() for k, v in ref.data_metric.items()},
            "random_ids": [int(s.meta["index"]) for s in ref.random_samples],
            "worst_ids": sorted(int(s.meta["index"]) for s in ref.worst_samples)}
    with open(OUT, "w") as f:
        json.dump(out, f)
    print("wrote", OUT, {k: (v["random_ids"], v["worst_ids"]) for k, v in out.items()})


if __name__ == "__main__":
    main()
