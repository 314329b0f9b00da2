"""Generator plugins: `Generator(grid_size, num_agents)`, `__call__(key) -> State`.

Mirrors routing_board_generation/rl_training/online_generators/
  uniform_generator.py:26-109                 Generator, UniformRandomGenerator
  parallel_random_walk_generator.py:28-77     ParallelRandomWalkGenerator
  random_seed_generator.py:15-57              SeedExtensionGenerator
  sequential_random_walk_generator.py:19-62   SequentialRandomWalkGenerator
so that `Connector(generator=...)` (rl_training/setup_train.py:112-161) takes them unchanged.
"""
from __future__ import annotations

import abc

from . import engine
from .board_generation import ParallelRandomWalkBoard, SeedExtensionBoard, SequentialRandomWalkBoard
from .types import State


class Generator(abc.ABC):
    """Base class for generators for the connector environment (uniform_generator.py:26-53)."""

    kind: str = ""

    def __init__(self, grid_size: int, num_agents: int) -> None:
        self._grid_size = grid_size
        self._num_agents = num_agents

    @property
    def grid_size(self) -> int:
        return self._grid_size

    @property
    def num_agents(self) -> int:
        return self._num_agents

    @abc.abstractmethod
    def __call__(self, key) -> State:
        """Generates a `Connector` state that contains the grid and the agents' layout."""


class _KernelGenerator(Generator):
    def __call__(self, key) -> State:
        keys, batched = engine.as_keys(key)
        st = engine.generator_state(self.kind, keys, self.grid_size, self.num_agents)
        return st if batched else st[0]


class UniformRandomGenerator(_KernelGenerator):
    """Start and target cells uniformly at random, may be unsolvable (uniform_generator.py:56-109)."""

    kind = "uniform"


class ParallelRandomWalkGenerator(_KernelGenerator):
    """Solvable boards from the parallel random walk (parallel_random_walk_generator.py:28-77)."""

    kind = "parallel_random_walk"

    def __init__(self, grid_size: int, num_agents: int) -> None:
        super().__init__(grid_size, num_agents)
        self.board_generator = ParallelRandomWalkBoard(grid_size, grid_size, num_agents)


class SeedExtensionGenerator(_KernelGenerator):
    """Solvable boards from seed extension with the default settings (random_seed_generator.py:15-57)."""

    kind = "seed_extension"

    def __init__(self, grid_size: int, num_agents: int) -> None:
        super().__init__(grid_size, num_agents)
        self.board_generator = SeedExtensionBoard(grid_size, grid_size, num_agents)


class SequentialRandomWalkGenerator(_KernelGenerator):
    """Boards from the sequential random walk (sequential_random_walk_generator.py:19-62; `online_seq_rw`,
    rl_training/setup_train.py:137-141).  A generation whose every attempt fails leaves all pins at (0, 0), as the
    reference's argwhere(size=2) fill value does."""

    kind = "sequential_random_walk"

    def __init__(self, grid_size: int, num_agents: int) -> None:
        super().__init__(grid_size, num_agents)
        self.board_generator = SequentialRandomWalkBoard(grid_size, grid_size, num_agents)
