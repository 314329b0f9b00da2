"""Connector pytrees, mirroring jumanji.environments.routing.connector.types
(jumanji==0.2.2, UPSTREAM) and the reference's copy of `Agent`
(board_generation/types_addition.py:16-42).  Leaves are torch CUDA tensors;
a leading batch axis plays the role of jax.vmap's mapped axis.
"""
from __future__ import annotations

from dataclasses import dataclass, field, fields, replace
from typing import Any, Dict

import torch

# jumanji.types.StepType
FIRST, MID, LAST = 0, 1, 2


class _Tree:
    def replace(self, **kw):
        return replace(self, **kw)

    def map(self, fn):
        """tree_map over tensor leaves."""
        out = {}
        for f in fields(self):
            v = getattr(self, f.name)
            if isinstance(v, _Tree):
                out[f.name] = v.map(fn)
            elif isinstance(v, dict):
                out[f.name] = {k: fn(x) for k, x in v.items()}
            else:
                out[f.name] = fn(v)
        return type(self)(**out)

    def __getitem__(self, idx):
        return self.map(lambda x: x[idx])


@dataclass
class Agent(_Tree):
    id: torch.Tensor        # int32 [..., N]
    start: torch.Tensor     # int32 [..., N, 2]
    target: torch.Tensor    # int32 [..., N, 2]
    position: torch.Tensor  # int32 [..., N, 2]

    @property
    def connected(self) -> torch.Tensor:
        """types_addition.py:29-33: all(position == target)."""
        return (self.position == self.target).all(dim=-1)


@dataclass
class State(_Tree):
    key: torch.Tensor         # uint32 [..., 2]
    grid: torch.Tensor        # int32 [..., G, G]
    step_count: torch.Tensor  # int32 [...]
    agents: Agent


@dataclass
class Observation(_Tree):
    grid: torch.Tensor         # int32 [..., N, G, G]
    action_mask: torch.Tensor  # bool  [..., N, 5]
    step_count: torch.Tensor   # int32 [...]


@dataclass
class TimeStep(_Tree):
    step_type: torch.Tensor  # int8 [...]
    reward: torch.Tensor     # float32 [..., N]
    discount: torch.Tensor   # float32 [..., N]
    observation: Observation
    extras: Dict[str, torch.Tensor] = field(default_factory=dict)

    def first(self) -> torch.Tensor:
        return self.step_type == FIRST

    def mid(self) -> torch.Tensor:
        return self.step_type == MID

    def last(self) -> torch.Tensor:
        return self.step_type == LAST
