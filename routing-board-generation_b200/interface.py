"""Generator registry subset: the names the reference registers for this path
(routing_board_generation/interface/board_generator_interface.py:32-71)."""
from __future__ import annotations

from enum import Enum

from .board_generation import ParallelRandomWalkBoard, SeedExtensionBoard


class BoardName(str, Enum):
    JAX_PARALLEL_RW = "offline_parallel_rw"
    JAX_SEED_EXTENSION = "offline_seed_extension"


class BoardGenerator:
    board_generator_dict = {
        BoardName.JAX_PARALLEL_RW: ParallelRandomWalkBoard,
        BoardName.JAX_SEED_EXTENSION: SeedExtensionBoard,
    }

    @classmethod
    def get_board_generator(cls, board_enum: BoardName):
        return cls.board_generator_dict[BoardName(board_enum)]
