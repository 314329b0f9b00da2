"""Board algorithms behind the reference's board API.

Mirrors (same class names, constructor arguments, method names, return layouts):
  routing_board_generation/board_generation_methods/jax_implementation/board_generation/
    parallel_random_walk.py:49-90   ParallelRandomWalkBoard.generate_board
    seed_extension.py:36-304        SeedExtensionBoard.return_solved_board / return_training_board /
                                    generate_starts_ends / return_seeded_board
    sequential_random_walk.py:23-431 SequentialRandomWalkBoard.generate / generate_starts_ends
A key of shape (2,) gives the reference's single-board result; keys of shape
(B, 2) give the jax.vmap'd result (leading axis B).  All work happens in the
CUDA kernels of librbg_b200.so.
"""
from __future__ import annotations

from typing import Tuple

import torch

from . import engine


def _check_square(rows: int, cols: int) -> int:
    if rows != cols:
        # parallel_random_walk.py:276,284 strides by rows and divmods by cols: only square grids are meaningful
        raise ValueError(f"only square boards are supported (rows={rows}, cols={cols})")
    return int(rows)


class ParallelRandomWalkBoard:
    def __init__(self, rows: int, cols: int, num_agents: int):
        self.rows = rows
        self.cols = cols
        self.num_agents = num_agents
        self._G = _check_square(rows, cols)

    def generate_board(self, key) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """key -> (heads[2,N], targets[2,N], solved_grid[R,C]) (parallel_random_walk.py:60-90)."""
        keys, batched = engine.as_keys(key)
        heads, targets, solved = engine.prw_generate(keys, self._G, self.num_agents)
        if batched:
            return heads, targets, solved
        return heads[0], targets[0], solved[0]

    def generate_board_with_stats(self, key):
        """As generate_board plus stats[..., 2] = (while-loop trips, collided moves)."""
        keys, batched = engine.as_keys(key)
        out = engine.prw_generate(keys, self._G, self.num_agents, with_stats=True)
        return out if batched else tuple(o[0] for o in out)


class SeedExtensionBoard:
    def __init__(self, rows: int, cols: int, num_agents: int = 0):
        self._rows = rows
        self._cols = cols
        self.grid_size = max(rows, cols)
        self._num_agents = num_agents
        self._wires_on_board = min(num_agents, rows * cols // 3)  # seed_extension.py:60-63
        self._G = _check_square(rows, cols)

    def return_solved_board(self, key, randomness: float = 0.0, two_sided: bool = True, extension_iterations: int = 1, extension_steps: float = 1e23) -> torch.Tensor:
        keys, batched = engine.as_keys(key)
        out = engine.seedext_solved(keys, self._G, self._num_agents, randomness, two_sided, extension_iterations, extension_steps)
        return out if batched else out[0]

    def return_training_board(self, key, randomness: float = 0.0, two_sided: bool = True, extension_iterations: int = 1, extension_steps: float = 1e23) -> torch.Tensor:
        """Solved board with the PATH cells zeroed (post_processor_utils_jax.py:433-444)."""
        board = self.return_solved_board(key, randomness, two_sided, extension_iterations, extension_steps)
        return board * ((board % 3) != 1).to(board.dtype)

    def generate_starts_ends(self, key, randomness: float = 0.0, two_sided: bool = True, extension_iterations: int = 1, extension_steps: float = 1e23):
        """-> ((start_rows[N], start_cols[N]), (end_rows[N], end_cols[N])) (seed_extension.py:257-304)."""
        keys, batched = engine.as_keys(key)
        starts, ends = engine.seedext_starts_ends(keys, self._G, self._num_agents, randomness, two_sided, extension_iterations, extension_steps)
        if batched:
            return (starts[:, 0], starts[:, 1]), (ends[:, 0], ends[:, 1])
        return (starts[0, 0], starts[0, 1]), (ends[0, 0], ends[0, 1])


class SequentialRandomWalkBoard:
    """One wire after the other, each a self-avoiding random walk from a random empty cell; the whole board is retried
    with a shorter maximum walk until every wire could start and move (sequential_random_walk.py:23-431).

    The reference class cannot be instantiated as shipped: it derives from the NumPy AbstractBoard without
    implementing `return_training_board` / `return_solved_board` (abstract_board.py:47-53), so
    `SequentialRandomWalkBoard(rows, cols, n)` raises TypeError there.  This mirror IS instantiable and reproduces what
    the method bodies compute (pinned by running them with the abstract-method check lifted:
    tests/tools/make_seqrw_fixtures.py)."""

    def __init__(self, rows: int, cols: int, num_agents: int = 3):
        self._rows = rows
        self._cols = cols
        self._num_agents = num_agents
        self._G = _check_square(rows, cols)
        if self._G < 3:
            # available_cells pads with jnp.full(rows - len(available_cells) + 1, -1) (sequential_random_walk.py:138-140)
            raise ValueError(f"SequentialRandomWalkBoard needs rows >= 3 (rows={rows})")

    def return_blank_board(self) -> torch.Tensor:
        return torch.zeros((self._rows, self._cols), dtype=torch.int32, device=engine._device())

    def generate(self, key, as_float32: bool = True) -> torch.Tensor:
        """key -> board[R,C] with all wires present, or a zero board when every attempt fails (:324-392).  float32
        codes like the reference (jnp.zeros' default dtype, :350); as_float32=False gives int32."""
        keys, batched = engine.as_keys(key)
        out = engine.seqrw_generate(keys, self._G, self._num_agents, as_float32)
        return out if batched else out[0]

    def generate_with_stats(self, key, as_float32: bool = False):
        """As generate plus stats[..., 2] = (attempt that succeeded, 0 = none; steps of that attempt)."""
        keys, batched = engine.as_keys(key)
        board, stats = engine.seqrw_generate(keys, self._G, self._num_agents, as_float32, with_stats=True)
        return (board, stats) if batched else (board[0], stats[0])

    def generate_starts_ends(self, key):
        """-> ((start_rows[N], start_cols[N]), (end_rows[N], end_cols[N])) (:394-431)."""
        keys, batched = engine.as_keys(key)
        starts, ends = engine.seqrw_starts_ends(keys, self._G, self._num_agents)
        if batched:
            return (starts[:, 0], starts[:, 1]), (ends[:, 0], ends[:, 1])
        return (starts[0, 0], starts[0, 1]), (ends[0, 0], ends[0, 1])
