"""Batched board statistics on the CUDA engine.

Mirrors `EvaluateEmptyBoard` (routing_board_generation/benchmarking/benchmarks/empty_board_evaluation.py:31-155),
the per-board consumer of solved boards in the reference's benchmark
(`evaluate_generator_outputs_averaged_on_n_boards`, :158-225): `scored_board`, `count_detours()` and the
`heatmap_score_diversity` entry of `board_statistics`.  A board [G,G] or a batch [B,G,G].

The reference's constructor first sends the board through `BoardProcessor`, which re-routes every wire along a
shortest path picked with an unseeded `random.shuffle` (board_processor.py:111-162); that preprocessing and the
wire-length / bend statistics derived from it have no deterministic result and are not reproduced: the board is
scored as given (for boards whose wires are already shortest paths the two coincide).
"""
from __future__ import annotations

from . import engine


class EvaluateEmptyBoard:
    def __init__(self, filled_training_board):
        self.filled_board = engine.as_tensor(filled_training_board)
        self.empty_slot_score = -2
        self.end_score = 3
        self.wire_score = 2
        self.scored_board, self._detours, self._diversity = engine.board_statistics(self.filled_board)
        self.board_statistics = {"count_detours": self._detours, "heatmap_score_diversity": self._diversity}

    def score_from_neighbours(self):
        return self.scored_board

    def count_detours(self, count_current_wire: bool = False):
        if not count_current_wire:
            return self._detours
        return engine.board_statistics(self.filled_board, count_current_wire=True, with_scores=False)[1]
