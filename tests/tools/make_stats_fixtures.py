#!/usr/bin/env python
"""Board statistics from the REFERENCE'S OWN class.

    python tests/tools/make_stats_fixtures.py      # writes tests/golden/board_stats_reference.npz

`EvaluateEmptyBoard` (benchmarking/benchmarks/empty_board_evaluation.py:31-155) is cut out of the reference file
with `ast` (the module itself imports matplotlib, which is not installed) and executed unmodified:
`assess_board`, `_change_heads_to_wire_ids`, `score_from_neighbours` (:41-88), `count_detours` (:99-137),
`get_wire_num` (:139-151) and `heatmap_score_diversity = len(np.unique(scored_board))` (:155).

Two sets of boards:
  * "raw": the class is constructed on the board as given (its `BoardProcessor` is replaced by a stand-in that
    returns the layout untouched), so the statistics are pure functions of the stored board;
  * "processed": the reference's real `BoardProcessor` (numpy_implementation/utils/board_processor.py) runs first,
    with `random.seed` fixed: it re-routes every wire along a shortest path through its own and EMPTY cells
    (remove_extraneous_path_cells) and the class scores THAT layout; the rewritten layout is stored as the input.
What is NOT reproduced: BoardProcessor.get_board_statistics (wire lengths / bends of those shuffled shortest paths).
"""
from __future__ import annotations

import ast
import contextlib
import io
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SHIM = os.path.join(HERE, "jax_shim")
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden", "board_stats_reference.npz")
GMAX = 14


def load_class(board_processor_cls):
    src = open(os.path.join(REF, "routing_board_generation/benchmarking/benchmarks/empty_board_evaluation.py")).read()
    cls = [n for n in ast.parse(src).body if isinstance(n, ast.ClassDef) and n.name == "EvaluateEmptyBoard"]
    assert len(cls) == 1
    ns = dict(np=np, BoardProcessor=board_processor_cls)
    exec(compile(ast.Module(body=cls, type_ignores=[]), "empty_board_evaluation.py", "exec"), ns)
    return ns["EvaluateEmptyBoard"]


class PassThroughProcessor:
    """Stand-in for BoardProcessor: the layout as given, and the two keys _get_board_statistics pops."""

    def __init__(self, board):
        self.board = np.array(board)

    def get_board_layout(self):
        return self.board

    def get_board_statistics(self):
        return dict(wire_lengths=[], wire_bends=[])


def main():
    sys.path.insert(0, SHIM)
    sys.path.insert(0, REF)
    sys.path.insert(0, ROOT)
    from routing_board_generation.board_generation_methods.numpy_implementation.utils.board_processor import BoardProcessor
    from oracle import oracle

    Raw, Processed = load_class(PassThroughProcessor), load_class(BoardProcessor)
    boards, Gs, Ns, kinds = [], [], [], []
    rng = np.random.default_rng(3)
    for (G, N, seed, n, src) in ((5, 3, 301, 60, "prw"), (10, 5, 302, 80, "prw"), (7, 12, 303, 40, "prw"), (14, 7, 304, 40, "prw"), (8, 4, 305, 60, "se"), (10, 5, 306, 60, "se"), (14, 7, 307, 30, "se")):
        keys = oracle.split(oracle.PRNGKey(seed), n)
        solved = oracle.prw_generate_batch(keys, G, N)[2] if src == "prw" else oracle.seedext_solved_batch(keys, G, N)[0]
        for b in solved:
            boards.append(np.asarray(b, dtype=np.int64))
            Gs.append(G), Ns.append(N), kinds.append(0)
            if rng.random() < 0.3:  # a few damaged boards: missing heads change the label map of _change_heads_to_wire_ids
                c = boards[-1].copy()
                c[c == 3 * int(rng.integers(N)) + 2] = 0
                boards.append(c)
                Gs.append(G), Ns.append(N), kinds.append(1)
    B = len(boards)
    inp = np.zeros((B, GMAX, GMAX), np.int16)
    scored = np.zeros((B, GMAX, GMAX), np.int32)
    det = np.zeros((B, 2), np.int32)
    div = np.zeros(B, np.int32)
    for i, (b, G) in enumerate(zip(boards, Gs)):
        ev = Raw(b.copy())
        inp[i, :G, :G] = b
        scored[i, :G, :G] = ev.scored_board
        det[i] = (ev.count_detours(False), ev.count_detours(True))
        div[i] = ev.board_statistics["heatmap_score_diversity"]
        assert ev.board_statistics["count_detours"] == det[i, 0]
    # the reference's own preprocessing in front (valid boards only: it raises on boards without a path)
    pin, pscored, pdet, pdiv, pG, pN = [], [], [], [], [], []
    sink = io.StringIO()
    for i, (b, G, N, k) in enumerate(zip(boards, Gs, Ns, kinds)):
        if k != 0 or i % 3:
            continue
        random.seed(1000 + i)
        try:
            with contextlib.redirect_stdout(sink):
                ev = Processed(b.copy().astype(np.int64))
        except Exception:  # noqa: BLE001  (zero-length wires make get_heads_and_targets / the BFS fail)
            continue
        x = np.zeros((GMAX, GMAX), np.int16)
        x[:G, :G] = ev.filled_board
        s = np.zeros((GMAX, GMAX), np.int32)
        s[:G, :G] = ev.scored_board
        pin.append(x), pscored.append(s), pdet.append(ev.board_statistics["count_detours"]), pdiv.append(ev.board_statistics["heatmap_score_diversity"]), pG.append(G), pN.append(N)
    np.savez_compressed(OUT, boards=inp, G=np.asarray(Gs, np.int8), N=np.asarray(Ns, np.int8), scored=scored, detours=det, diversity=div,
                        p_boards=np.stack(pin), p_G=np.asarray(pG, np.int8), p_N=np.asarray(pN, np.int8), p_scored=np.stack(pscored), p_detours=np.asarray(pdet, np.int32), p_diversity=np.asarray(pdiv, np.int32))
    print("wrote", OUT, os.path.getsize(OUT), "bytes:", B, "raw boards,", len(pin), "processed boards; detours mean", det[:, 0].mean(), "diversity mean", div.mean())


if __name__ == "__main__":
    main()
