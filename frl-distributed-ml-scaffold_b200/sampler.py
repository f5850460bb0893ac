"""Data-parallel partitioning of the sample indices (reference sampler.py:17-87).

The index stream must be bit-exact with the reference, so the permutation is drawn with the
same CPU ``torch.randperm`` from a ``torch.Generator`` seeded with the epoch number (epochs
start at 1, reference solver_worker.py:429,785) and then padded and strided exactly as there.
"""
import math
from typing import Iterator, List

import torch
import torch.utils.data.distributed

from .types import ShuffleType


def per_node_randperm(max: int, *, node_idx: int, node_count: int,
                      generator: torch.Generator) -> List[int]:
    """Permutation of this node's contiguous chunk of ``range(max)``.

    Every node gets ``ceil(max/node_count)`` indices; the last node's short chunk is padded by
    recycling its own leading indices so all ranks do the same amount of work.
    """
    chunk = math.ceil(max / node_count)
    first = node_idx * chunk
    have = min(max - first, chunk)
    order = (torch.randperm(have, generator=generator) + chunk * node_idx).tolist()
    return order + order[: chunk - have]


class ScaffoldSampler(torch.utils.data.distributed.DistributedSampler):
    def __init__(self, dataset, *, shuffle_type: ShuffleType, node_idx: int,
                 node_count: int) -> None:
        super().__init__(dataset,
                         num_replicas=torch.distributed.get_world_size(),
                         rank=torch.distributed.get_rank())
        self._shuffle_type = shuffle_type
        self._node_idx = node_idx
        self._node_count = node_count

    def rank_index_tensor(self) -> torch.Tensor:
        """This rank's sample ids of the current epoch as an int64 tensor — what ``__iter__``
        yields, without materialising ``len(dataset)`` Python ints (the batched input path
        consumes it directly)."""
        gen = torch.Generator()
        gen.manual_seed(self.epoch)
        n = len(self.dataset)
        if self._shuffle_type == ShuffleType.PER_NODE_RANDPERM:
            ranks_per_node = self.num_replicas // self._node_count
            chunk = math.ceil(n / self._node_count)
            first = self._node_idx * chunk
            have = min(n - first, chunk)
            order = torch.randperm(have, generator=gen) + chunk * self._node_idx
            order = torch.cat([order, order[: chunk - have]])
            return order[self.rank % ranks_per_node:: ranks_per_node].contiguous()
        if self._shuffle_type != ShuffleType.RANDPERM:
            raise ValueError("Unhandled shuffle type %s", self._shuffle_type)
        from .device_loader import randperm_quiet
        order = randperm_quiet(n, gen) if self.shuffle else torch.arange(n)
        order = torch.cat([order, order[: self.total_size - n]])          # pad with the head
        assert len(order) == self.total_size
        mine = order[self.rank: self.total_size: self.num_replicas].contiguous()
        assert len(mine) == self.num_samples
        return mine

    def rank_indices(self) -> List[int]:
        return self.rank_index_tensor().tolist()

    def __iter__(self) -> Iterator[int]:
        return iter(self.rank_indices())
