"""One task of a multitask problem: head, loss, targets, metrics (reference task.py:34-80)."""
from abc import abstractmethod
from typing import Dict, Generic, List, NamedTuple, Optional, Sequence, Set, Tuple, TypeVar

import torch
import torch.nn as nn
import torch.nn.modules.loss as L

from .types import SampleSummary

SampleMetaT = TypeVar("SampleMetaT")
BatchMetaT = TypeVar("BatchMetaT")
TransformT = TypeVar("TransformT", bound=NamedTuple)


class Task(Generic[TransformT, SampleMetaT, BatchMetaT]):
    # -- network / loss ------------------------------------------------------------------
    @property
    @abstractmethod
    def network_head(self) -> nn.Module:
        ...

    @property
    @abstractmethod
    def criterion(self) -> L._Loss:
        ...

    @property
    @abstractmethod
    def criterion_weight(self) -> float:
        ...

    # -- data ----------------------------------------------------------------------------
    @abstractmethod
    def get_target(self, tensors: Dict[str, torch.Tensor], transform: TransformT
                   ) -> Tuple[Sequence[torch.Tensor], SampleMetaT]:
        ...

    # -- metrics -------------------------------------------------------------------------
    @abstractmethod
    def compute_batch_metrics(self, meta: BatchMetaT, target: Tuple[torch.Tensor, ...],
                              output: torch.Tensor) -> "Dict[str, object]":
        ...

    @property
    @abstractmethod
    def rankable_metrics(self) -> "Set[Tuple[str, object]]":
        ...

    @abstractmethod
    def summarize_epoch_metrics(self, batch_metrics) -> Dict[str, float]:
        ...

    @abstractmethod
    def summarize_epoch_samples(self, data: List[torch.Tensor],
                                target: Tuple[torch.Tensor, ...], meta: BatchMetaT,
                                output: torch.Tensor, metric: Optional[dict]
                                ) -> List[SampleSummary]:
        ...
