"""Multitask specialisation of ``Problem`` (reference multitask_problem.py:38-132)."""
from abc import abstractmethod
from itertools import chain
from typing import Any, Dict, Generic, List, NamedTuple, Optional, Sequence, Tuple, Type, TypeVar

import torch

from .model import MultiTaskModel
from .problem import BatchMetrics, EpochMetrics, Ordering, Problem
from .task import Task
from .transform import MultifieldTransform, Sample
from .types import SampleSummary, Split

SampleMetaT = TypeVar("SampleMetaT", bound=NamedTuple)
BatchMetaT = TypeVar("BatchMetaT")
AnnoParamT = TypeVar("AnnoParamT")
TransformT = TypeVar("TransformT", bound=NamedTuple)


class MultiTaskTransform(Generic[TransformT, SampleMetaT], MultifieldTransform[SampleMetaT]):
    """Shared source transform followed by one ``get_target`` per task."""

    SampleMetaType: Type[SampleMetaT]
    _tasks: Sequence[Task]

    def __init__(self, tasks: Sequence[Task], SampleMetaType: Type[SampleMetaT]) -> None:
        self._tasks = tasks
        self.SampleMetaType = SampleMetaType

    @abstractmethod
    def transform_source_data(self, tensors: Dict[str, torch.Tensor], split: Split
                              ) -> Tuple[Sequence[torch.Tensor], TransformT]:
        ...

    def transform(self, data: Dict[str, Any], split: Split) -> Tuple[Sample, SampleMetaT]:
        tensors = {name: torch.from_numpy(arr) for name, arr in data.items()}
        source, applied = self.transform_source_data(tensors, split)
        targets, meta_fields = [], {}
        for task in self._tasks:
            tgt, meta = task.get_target(tensors, applied)
            targets.append(tgt)
            meta_fields.update(meta._asdict())
        return Sample(data=source, target=targets), self.SampleMetaType(**meta_fields)


class MultiTaskProblem(Generic[BatchMetaT, AnnoParamT], Problem[BatchMetaT, AnnoParamT]):
    _tasks: Sequence[Task]
    BatchMetaType: Type[BatchMetaT]

    @abstractmethod
    def get_model_base(self) -> torch.nn.Module:
        ...

    def get_model(self) -> torch.nn.Module:
        return MultiTaskModel(model_base=self.get_model_base(),
                              additional_layers=[t.network_head for t in self._tasks])

    def refine_batch_meta(self, meta: Dict[str, Any]) -> BatchMetaT:
        return self.BatchMetaType(**meta)

    def compute_batch_metrics(self, meta: BatchMetaT, target: List[Tuple[torch.Tensor, ...]],
                              output: List[torch.Tensor], device: torch.device) -> BatchMetrics:
        merged: BatchMetrics = {}
        for i, task in enumerate(self._tasks):
            merged.update(task.compute_batch_metrics(meta, target[i], output[i]))
        return merged

    def get_rankable_metric(self) -> Tuple[str, Ordering]:
        return list(chain.from_iterable(t.rankable_metrics for t in self._tasks))[0]

    def summarize_epoch_metrics(self, batch_metrics: BatchMetrics) -> EpochMetrics:
        merged: EpochMetrics = {}
        for task in self._tasks:
            merged.update(task.summarize_epoch_metrics(batch_metrics))
        return merged

    def summarize_epoch_samples(self, data: List[torch.Tensor],
                                target: List[Tuple[torch.Tensor, ...]], meta: BatchMetaT,
                                output: List[torch.Tensor],
                                metric: Optional[BatchMetrics] = None) -> List[SampleSummary]:
        per_task = (task.summarize_epoch_samples(data, target[i], meta, output[i], metric)
                    for i, task in enumerate(self._tasks))
        return list(chain.from_iterable(per_task))
