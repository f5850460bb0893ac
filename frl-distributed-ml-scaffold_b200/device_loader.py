"""Batched, device-side input pipeline for datasets that live in pinned host memory.

The reference feeds the loop through ``DataLoader`` -> per-sample ``__getitem__`` -> Python
``MultifieldTransform`` -> ``default_collate`` -> pageable H2D copy (reference
solver_worker.py:462-469, 805-832; transform.py:25-38): ~0.15 ms of host Python per sample, which
starves a B200 at batch 4096.  A dataset that exposes

    pinned_fields     Dict[str, Tensor]   whole raw dataset, one pinned host tensor per field
    device_transform  DeviceBatchTransform

is consumed here instead: the index stream comes from the very same sampler/``DataLoader``
machinery (so sample order and global-RNG consumption stay bit-identical to the reference), the
rows of batch *k+1* are pulled over PCIe by ``frl_gather_rows`` on a copy stream while batch *k*
trains, and the per-sample arithmetic runs once per batch on the device (``frl_preproc_affine``).
"""
from typing import Dict, Iterator, List, Optional, Tuple

import torch
import torch.utils.data

from . import _native
from .transform import DeviceBatchTransform
from .types import Split


class _IndexOnly(torch.utils.data.Dataset):
    def __init__(self, n: int) -> None:
        self._n = n

    def __len__(self) -> int:
        return self._n

    def __getitem__(self, i: int) -> int:
        return i


def _collate_indices(items: List[int]) -> torch.Tensor:
    return torch.tensor(items, dtype=torch.int64)


def supports_device_batches(dataset) -> bool:
    return (isinstance(getattr(dataset, "pinned_fields", None), dict)
            and isinstance(getattr(dataset, "device_transform", None), DeviceBatchTransform))


class DeviceBatchLoader:
    """Iterates ``(data, target, raw_meta)`` like the reference's DataLoader, already on device."""

    def __init__(self, dataset, *, batch_size: int, sampler, device: torch.device,
                 out_dtype: torch.dtype = torch.float32, depth: int = 2) -> None:
        self.dataset = dataset
        self.device = device
        self.batch_size = batch_size
        self.out_dtype = out_dtype
        self.depth = max(depth, 2)
        # same construction as the reference's loader -> same sampler classes, same RNG draws
        self._index_loader = torch.utils.data.DataLoader(
            _IndexOnly(len(dataset)), batch_size=batch_size, shuffle=sampler is None,
            sampler=sampler, num_workers=0, collate_fn=_collate_indices)
        self.sampler = self._index_loader.sampler
        self._fields: Dict[str, torch.Tensor] = dataset.pinned_fields
        for name, t in self._fields.items():
            if not (t.is_pinned() and t.is_contiguous()):
                raise ValueError(f"field {name!r} must be a contiguous pinned host tensor")
        self._copy_stream = torch.cuda.Stream(device=device)
        self._slots = []
        for _ in range(self.depth):
            slot = {name: torch.empty((batch_size,) + tuple(t.shape[1:]), dtype=t.dtype, device=device)
                    for name, t in self._fields.items()}
            slot["__idx_host"] = torch.empty(batch_size, dtype=torch.int64, pin_memory=True)
            slot["__idx_dev"] = torch.empty(batch_size, dtype=torch.int64, device=device)
            self._slots.append(slot)
        self._ready = [torch.cuda.Event() for _ in range(self.depth)]
        self._freed = [torch.cuda.Event() for _ in range(self.depth)]
        self.h2d_bytes_per_batch = sum(t[0].numel() * t.element_size() for t in self._fields.values()
                                       ) * batch_size + 8 * batch_size

    def __len__(self) -> int:
        return len(self._index_loader)

    def _upload(self, k: int, idx: torch.Tensor) -> int:
        s = k % self.depth
        slot = self._slots[s]
        n = idx.numel()
        slot["__idx_host"][:n].copy_(idx)
        with torch.cuda.stream(self._copy_stream):
            self._copy_stream.wait_event(self._freed[s])          # previous user of the slot is done
            slot["__idx_dev"][:n].copy_(slot["__idx_host"][:n], non_blocking=True)
            for name, src in self._fields.items():
                _native.gather_rows(src, slot["__idx_dev"][:n], slot[name][:n])
            self._ready[s].record()
        return n

    def __iter__(self) -> Iterator[Tuple[List[torch.Tensor], List[Tuple[torch.Tensor, ...]], dict]]:
        split: Split = self.dataset.data_type
        transform: DeviceBatchTransform = self.dataset.device_transform
        for ev in self._freed:
            ev.record()
        batches = iter(self._index_loader)
        pending: List[int] = []
        k_up = 0
        try:
            pending.append(self._upload(k_up, next(batches)))
            k_up += 1
        except StopIteration:
            return
        k = 0
        while pending:
            try:                                  # keep one batch in flight behind the current one
                pending.append(self._upload(k_up, next(batches)))
                k_up += 1
            except StopIteration:
                pass
            n = pending.pop(0)
            s = k % self.depth
            slot = self._slots[s]
            torch.cuda.current_stream().wait_event(self._ready[s])
            raw = {name: slot[name][:n] for name in self._fields}
            data, target = transform.apply(raw, split, self.out_dtype)
            meta = transform.meta(raw, slot["__idx_dev"][:n])
            yield data, target, meta
            # the consumer has issued everything that reads this slot: let the copy stream reuse it
            self._freed[s].record()
            k += 1
