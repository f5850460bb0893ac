"""Multitask model container and small layer helpers (reference model.py:16-88).

``MultiTaskModel`` runs a shared trunk once and feeds every task head from the same trunk
output.  The trunk receives the minibatch as a ``List[Tensor]`` (the loop hands the whole
data list to ``model(data)``, reference solver_worker.py:465,551) — ``ListSelect`` picks one.
"""
from collections import deque
from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn as nn


class MultiTaskModel(nn.Module):
    def __init__(self, model_base: nn.Module, additional_layers: Sequence[nn.Module],
                 additional_layer_names: Optional[List[str]] = None) -> None:
        super().__init__()
        self.model_base = model_base
        self.additional_layers = nn.ModuleList(additional_layers)
        if additional_layer_names is None:
            additional_layer_names = list(range(len(self.additional_layers)))
        elif len(additional_layer_names) != len(self.additional_layers):
            raise AssertionError("one name per head expected")
        self.additional_layer_names = additional_layer_names

    def forward(self, x):
        shared = self.model_base(x)
        return [head(shared) for head in self.additional_layers]

    def final_shared_params(self, outputs: List[torch.Tensor]) -> torch.Tensor:
        """Last trunk parameter on the autograd path of the first head's output.

        Breadth-first walk of ``grad_fn.next_functions`` (reference model.py:32-50); the first
        AccumulateGrad node whose variable is a trunk parameter wins.  GradNorm measures the
        per-task gradient norms there.
        """
        trunk_ids = {id(p) for p in self.model_base.parameters()}
        frontier = deque([outputs[0].grad_fn])
        while frontier:
            node = frontier.popleft()
            for nxt, _ in node.next_functions:
                if nxt is None:
                    continue
                var = getattr(nxt, "variable", None)
                if var is not None and id(var) in trunk_ids:
                    return var
                frontier.append(nxt)
        raise RuntimeError("Unable to find any shared parameters in the model")


class View(nn.Module):
    def __init__(self, dims: Tuple[int, ...]) -> None:
        super().__init__()
        self.dims = dims

    def forward(self, *inputs) -> torch.Tensor:
        return inputs[0].view(self.dims)


class MulConstant(nn.Module):
    def __init__(self, constant: float) -> None:
        super().__init__()
        self.constant = constant

    def forward(self, *inputs) -> torch.Tensor:
        return inputs[0] * self.constant


class ListSelect(nn.Module):
    def __init__(self, *, sel_index: int, num_elements: int):
        super().__init__()
        self._sel_index = sel_index
        self._num_elements = num_elements

    def forward(self, items):
        assert len(items) == self._num_elements, "number of elements does not match!"
        return items[self._sel_index]
