"""Per-sample transform API (reference transform.py:20-44) plus its batched device-side twin.

``MultifieldTransform`` is the reference contract: a raw record (dict of ndarrays) becomes
``(Sample(data, target), meta)`` and ``__call__`` flattens that to the triple the DataLoader
collates.  ``DeviceBatchTransform`` is the B200 addition: the same arithmetic applied to a
whole batch after the raw bytes reached HBM, through the ``frl_preproc_affine`` kernel.
"""
from abc import ABC, abstractmethod
from typing import Any, Dict, Generic, List, NamedTuple, Sequence, Tuple, TypeVar, Union

import numpy as np
from torch import Tensor

from .types import Split

SampleMetaT = TypeVar("SampleMetaT", bound=NamedTuple)


class Sample(NamedTuple):
    data: Sequence[Tensor]
    target: Sequence[Union[Tensor, Tuple[Tensor, ...]]]


class MultifieldTransform(ABC, Generic[SampleMetaT]):
    def __call__(self, data: Dict[str, np.ndarray], split: Split
                 ) -> Tuple[Sequence[Tensor], Sequence[Tensor], Dict[str, Any]]:
        sample, meta = self.transform(data, split)
        # default_collate handles dicts/lists/tensors only: drop unset meta fields
        kept = {k: v for k, v in meta._asdict().items() if v is not None}
        return sample.data, sample.target, kept

    @abstractmethod
    def transform(self, data: Dict[str, np.ndarray], split: Split
                  ) -> Tuple[Sample, SampleMetaT]:
        ...


class DeviceBatchTransform(ABC):
    """Batched, on-device counterpart of ``MultifieldTransform`` (extension).

    A dataset that can hand out whole raw batches (``get_raw_batch``) pairs with one of these:
    ``stage`` lists the raw host tensors to ship (pinned, any dtype), ``apply`` turns the
    device copies into ``(data, target)`` exactly as the per-sample transform would.
    """

    @abstractmethod
    def apply(self, raw: Dict[str, Tensor], split: Split
              ) -> Tuple[List[Tensor], List[Tuple[Tensor, ...]]]:
        ...
