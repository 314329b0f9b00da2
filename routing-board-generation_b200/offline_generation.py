"""Dataset-backed generator: pre-generate K boards, `__call__(key)` picks one.

Mirrors routing_board_generation/rl_training/offline_generation/dataset_generator_jax.py:20-141
(`BoardDatasetGeneratorJAX`): same constructor arguments and defaults (note randomness=1,
two_sided=False here, unlike the online SeedExtensionGenerator), boards generated from
`jax.random.split(PRNGKey(0), number_of_boards)`, `__call__` = `key, _ = split(key)`;
`randint(key, (), 0, K)`.  The reference fills the dataset with a Python loop of jitted calls;
here it is one batched launch.
"""
from __future__ import annotations

from . import engine
from .board_generation import ParallelRandomWalkBoard, SeedExtensionBoard
from .online_generators import Generator
from .types import State


class BoardDatasetGeneratorJAX(Generator):
    kind = "dataset"

    def __init__(self, grid_size: int, num_agents: int, randomness: float = 1, two_sided: bool = False, extension_iterations: int = 1, extension_steps: float = 1e23,
                 board_name: str = "offline_seed_extension", number_of_boards: int = 10000, generate_solved_boards: bool = False) -> None:
        super().__init__(grid_size, num_agents)
        self.board_name = board_name
        self.randomness = randomness
        self.two_sided = two_sided
        self.extension_iterations = extension_iterations
        self.extension_steps = extension_steps
        if board_name == "offline_uniform":
            # dataset_generator_jax.py:42-49,99-104 unpacks a State into three values: the reference's own path raises
            raise NotImplementedError("board_name='offline_uniform' does not work in the reference either (generate_n_boards unpacks a State)")
        if board_name == "offline_seed_extension":
            self.board_generator = SeedExtensionBoard(grid_size, grid_size, num_agents)
        else:
            self.board_generator = ParallelRandomWalkBoard(grid_size, grid_size, num_agents)
        self.heads, self.targets, self.solved_boards = self.generate_n_boards(engine.PRNGKey(0), number_of_boards, generate_solved_boards)

    def generate_n_boards(self, key, n_boards: int = 10, generate_solved_boards: bool = False):
        """-> heads[K,2,N], targets[K,2,N], solved boards [K,G,G] (or None); keys = split(key, n_boards) (:76)."""
        keys = engine.split(key, n_boards)
        opts = (self.randomness, self.two_sided, self.extension_iterations, self.extension_steps)
        if self.board_name == "offline_seed_extension":
            solved = self.board_generator.return_solved_board(keys, *opts) if generate_solved_boards else None
            (sr, sc), (er, ec) = self.board_generator.generate_starts_ends(keys, *opts)
            import torch

            return torch.stack((sr, sc), dim=1).contiguous(), torch.stack((er, ec), dim=1).contiguous(), solved
        heads, targets, solved = self.board_generator.generate_board(keys)
        return heads.contiguous(), targets.contiguous(), solved

    def __call__(self, key) -> State:
        keys, batched = engine.as_keys(key)
        st = engine.dataset_state(keys, self.grid_size, self.num_agents, self.heads, self.targets)
        return st if batched else st[0]

    def print_board(self, which_board: int):
        """Pins-only grid of stored board `which_board` (dataset_generator_jax.py:143-155)."""
        import torch

        g = torch.zeros((self.grid_size, self.grid_size), dtype=torch.int32, device=self.heads.device)
        ids = torch.arange(self.num_agents, dtype=torch.int32, device=g.device)
        h, t = self.heads[which_board].long(), self.targets[which_board].long()
        g[h[0], h[1]] = 2 + 3 * ids
        g[t[0], t[1]] = 3 + 3 * ids
        return g
