#!/usr/bin/env python
"""Pin the board-validity rules to the REFERENCE'S OWN NumPy code.

    python tests/tools/make_validity_fixtures.py      # writes tests/golden/validity_reference.npz

The reference's validity code is plain NumPy and runs here as it is (imported unmodified from
/root/reference; only its `import jax` / `import jumanji...constants` lines need the stand-ins in
tests/tools/jax_shim):

  * numpy_implementation/utils/post_processor_utils_numpy.py:88-155  verify_wire_validity,
    num_wire_neighbors (module level, takes the bare layout)
  * numpy_implementation/utils/board_processor.py:341-440  BoardProcessor.is_valid_board,
    verify_encodings_range, verify_number_heads_tails, verify_wire_validity
  * numpy_implementation/utils/board_processor.py:111-162  BoardProcessor.get_path_from_head_and_target
    (BFS through the wire's own cells and EMPTY cells; raises PathNotFoundError)

Input boards (inputs only, any source would do): ParallelRandomWalk and SeedExtension boards from the C
oracle, as generated and after random corruptions (stray / missing / duplicated codes, cut wires,
out-of-range codes), plus hand-made zero-length wires.  For every board the script stores what the
reference says:

  outcome      BoardProcessor.is_valid_board(): 0 ok, 1 EncodingOutOfRangeError, 2 MissingHeadTailError,
               3 InvalidWireStructureError (first failing rule, in the reference's order)
  enc_ok       verify_encodings_range()
  missing      verify_number_heads_tails() raised MissingHeadTailError
  wire_valid   verify_wire_validity() -- the method and the module-level function must agree
  path_found   [N] per wire: get_path_from_head_and_target(first head, first target) on a FRESH copy of
               the board (the method rewrites the layout, so wires are asked independently; whether a
               path exists does not depend on its random.shuffle); -1 when the wire lacks a head or target
"""
from __future__ import annotations

import contextlib
import io
import os
import random
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SHIM = os.path.join(HERE, "jax_shim")
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden", "validity_reference.npz")

GMAX = 10  # boards are stored zero-padded to GMAX x GMAX


def corrupt(rng: np.random.Generator, board: np.ndarray, N: int) -> np.ndarray:
    b = board.copy()
    G = b.shape[0]
    kind = rng.integers(0, 11)
    cells = np.argwhere(b > 0)
    r, c = (int(x) for x in rng.integers(0, G, 2))
    if kind == 0:  # random code anywhere
        b[r, c] = rng.integers(0, 3 * N + 1)
    elif kind == 1 and len(cells):  # cut a wire
        rr, cc = cells[rng.integers(len(cells))]
        b[rr, cc] = 0
    elif kind == 2 and len(cells):  # duplicate some code next to itself / somewhere
        rr, cc = cells[rng.integers(len(cells))]
        b[r, c] = b[rr, cc]
    elif kind == 3:  # out-of-range high
        b[r, c] = 3 * N + int(rng.integers(1, 4))
    elif kind == 4:  # negative
        b[r, c] = -int(rng.integers(1, 4))
    elif kind == 5 and len(cells):  # swap two cells
        i, j = rng.integers(len(cells), size=2)
        (r1, c1), (r2, c2) = cells[i], cells[j]
        b[r1, c1], b[r2, c2] = b[r2, c2], b[r1, c1]
    elif kind == 6:  # drop a head
        w = int(rng.integers(N))
        b[b == 3 * w + 2] = 0
    elif kind == 7:  # drop a target
        w = int(rng.integers(N))
        b[b == 3 * w + 3] = 0
    elif kind == 8:  # head -> path (wire ends in a PATH cell)
        w = int(rng.integers(N))
        b[b == 3 * w + 2] = 3 * w + 1
    elif kind == 9:  # several cuts
        for _ in range(int(rng.integers(2, 5))):
            if len(cells):
                rr, cc = cells[rng.integers(len(cells))]
                b[rr, cc] = 0
    else:  # zero-length wire: the whole wire collapses into a lone TARGET
        w = int(rng.integers(N))
        t = np.argwhere(b == 3 * w + 3)
        b[(b >= 3 * w + 1) & (b <= 3 * w + 3)] = 0
        if len(t):
            b[t[0][0], t[0][1]] = 3 * w + 3
    return b


def main():
    sys.path.insert(0, SHIM)
    sys.path.insert(0, REF)
    sys.path.insert(0, ROOT)
    from routing_board_generation.board_generation_methods.numpy_implementation.utils import board_processor as bp
    from routing_board_generation.board_generation_methods.numpy_implementation.utils import exceptions as ex
    from routing_board_generation.board_generation_methods.numpy_implementation.utils import post_processor_utils_numpy as pp
    from oracle import oracle  # input boards only

    rng = np.random.default_rng(20261018)
    random.seed(7)
    boards, Gs, Ns, srcs = [], [], [], []
    for (G, N, seed, n, src) in ((5, 3, 101, 90, "prw"), (6, 6, 102, 90, "prw"), (8, 4, 103, 90, "prw"), (10, 5, 104, 90, "prw"), (7, 12, 105, 60, "prw"),
                                 (6, 3, 106, 60, "se"), (8, 4, 107, 60, "se"), (10, 5, 108, 60, "se")):
        keys = oracle.split(oracle.PRNGKey(seed), n)
        if src == "prw":
            _, _, solved, _ = oracle.prw_generate_batch(keys, G, N)
        else:
            solved, _ = oracle.seedext_solved_batch(keys, G, N)
        for b in solved:
            b = np.asarray(b, dtype=np.int64)
            boards.append(b)
            Gs.append(G), Ns.append(N), srcs.append(0 if src == "prw" else 1)
            for _ in range(3):  # three corruptions of every board, some of them compounded
                cb = corrupt(rng, b, N)
                if rng.random() < 0.25:
                    cb = corrupt(rng, cb, N)
                boards.append(cb)
                Gs.append(G), Ns.append(N), srcs.append(2)
    print(len(boards), "boards")

    names = {None: 0, "EncodingOutOfRangeError": 1, "MissingHeadTailError": 2, "InvalidWireStructureError": 3}
    B = len(boards)
    out_boards = np.zeros((B, GMAX, GMAX), np.int16)
    outcome = np.zeros(B, np.int8)
    enc_ok = np.zeros(B, np.int8)
    missing = np.zeros(B, np.int8)
    wire_valid = np.zeros(B, np.int8)
    path_found = np.full((B, 12), -1, np.int8)
    sink = io.StringIO()
    for i, (b, G, N) in enumerate(zip(boards, Gs, Ns)):
        out_boards[i, :G, :G] = b

        def proc(layout):
            p = bp.BoardProcessor.__new__(bp.BoardProcessor)  # the constructor would already run the BFS (and rewrite the board)
            p.board_layout = layout
            p.board = types.SimpleNamespace(wires_on_board=N, rows=G, cols=G)
            p.rows, p.cols = G, G
            return p

        with contextlib.redirect_stdout(sink):  # the reference prints on failures
            p = proc(b.copy())
            try:
                p.is_valid_board()
                outcome[i] = 0
            except Exception as e:  # noqa: BLE001
                outcome[i] = names[type(e).__name__]
            enc_ok[i] = bool(p.verify_encodings_range())
            try:
                p.verify_number_heads_tails()
            except ex.MissingHeadTailError:
                missing[i] = 1
            wv = bool(p.verify_wire_validity())
            assert wv == bool(pp.verify_wire_validity(b.copy())), "method and module-level verify_wire_validity disagree"
            wire_valid[i] = wv
            for w in range(N):
                h, t = np.argwhere(b == 3 * w + 2), np.argwhere(b == 3 * w + 3)
                if len(h) == 0 or len(t) == 0:
                    continue
                verdicts = set()
                for rep in range(2):  # two shuffles: existence of a path must not depend on the move order
                    q = proc(b.copy())
                    try:
                        q.get_path_from_head_and_target(tuple(int(x) for x in h[0]), tuple(int(x) for x in t[0]))
                        verdicts.add(1)
                    except ex.PathNotFoundError:
                        verdicts.add(0)
                assert len(verdicts) == 1
                path_found[i, w] = verdicts.pop()
    np.savez_compressed(OUT, boards=out_boards, G=np.asarray(Gs, np.int8), N=np.asarray(Ns, np.int8), source=np.asarray(srcs, np.int8),
                        outcome=outcome, enc_ok=enc_ok, missing=missing, wire_valid=wire_valid, path_found=path_found)
    print("wrote", OUT, os.path.getsize(OUT), "bytes;", "outcomes", np.bincount(outcome, minlength=4).tolist(),
          "wire_valid", int(wire_valid.sum()), "missing", int(missing.sum()), "path not found", int((path_found == 0).sum()))


if __name__ == "__main__":
    main()
