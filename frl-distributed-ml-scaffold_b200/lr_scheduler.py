"""Per-epoch learning-rate schedules (reference lr_scheduler.py:16-78, solver.py:191-218).

Both schedules are closed forms of the epoch counter, so besides the ``_LRScheduler`` classes
the reference API exposes, the pure functions ``drop_epochs_lr`` / ``warmup_multistep_lr`` give
the scalar that is handed to the fused update kernel as its ``lr`` argument.  The scheduler is
stepped once per EPOCH (reference solver_worker.py:790).
"""
from bisect import bisect_right
from typing import List, Sequence

import numpy as np
import torch
from torch.optim.optimizer import Optimizer


def drop_epochs_lr(base_lr: float, last_epoch: int, drop_epochs: Sequence[float],
                   gamma: float = 0.1) -> float:
    n_drops = int(np.sum([last_epoch + 1 >= d for d in drop_epochs]))
    return base_lr * gamma ** n_drops


def warmup_multistep_lr(base_lr: float, last_epoch: int, milestones: Sequence[float],
                        gamma: float = 0.1, warmup_factor: float = 1.0 / 3,
                        warmup_iters: int = 500, warmup_method: str = "linear") -> float:
    factor = 1
    if last_epoch < warmup_iters:
        if warmup_method == "constant":
            factor = warmup_factor
        elif warmup_method == "linear":
            alpha = last_epoch / warmup_iters
            factor = warmup_factor * (1 - alpha) + alpha
    return base_lr * factor * gamma ** bisect_right(milestones, last_epoch)


class DropEpochsScheduler(torch.optim.lr_scheduler._LRScheduler):
    def __init__(self, optimizer: Optimizer, drop_epochs: List[int], *,
                 gamma: float = 0.1, last_epoch: int = -1) -> None:
        self._drop_epochs = drop_epochs
        self._gamma = gamma
        super().__init__(optimizer, last_epoch)

    def get_lr(self) -> List[float]:
        return [drop_epochs_lr(b, self.last_epoch, self._drop_epochs, self._gamma)
                for b in self.base_lrs]


class WarmupMultiStepLR(torch.optim.lr_scheduler._LRScheduler):
    def __init__(self, optimizer, milestones, gamma=0.1, warmup_factor=1.0 / 3,
                 warmup_iters=500, warmup_method="linear", last_epoch=-1):
        if list(milestones) != sorted(milestones):
            raise ValueError(
                "Milestones should be a list of increasing integers. Got {}", milestones)
        if warmup_method not in ("constant", "linear"):
            raise ValueError(
                "Only 'constant' or 'linear' warmup_method accepted, got {}".format(warmup_method))
        self.milestones = milestones
        self.gamma = gamma
        self.warmup_factor = warmup_factor
        self.warmup_iters = warmup_iters
        self.warmup_method = warmup_method
        super().__init__(optimizer, last_epoch)

    def get_lr(self):
        return [warmup_multistep_lr(b, self.last_epoch, self.milestones, self.gamma,
                                    self.warmup_factor, self.warmup_iters, self.warmup_method)
                for b in self.base_lrs]
