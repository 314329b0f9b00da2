#!/usr/bin/env python
"""Secondary CPU baseline named by BASELINE.json: the reference's NumPy BFSBoard generator
(routing_board_generation/board_generation_methods/numpy_implementation/board_generation/bfs_board.py,
used as in README.md:88-93 / benchmarking/utils/benchmark_utils.py:59-80), timed as-is from
/root/reference.  It only runs where the reference is mounted (the build container, not the GPU
box); its module imports jax.numpy for type hints, which tests/tools/jax_shim satisfies.
Unseeded python `random`: throughput only, no parity.
    python tools/time_bfs_board_numpy.py [n_boards]
"""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "tools", "jax_shim"))
sys.path.insert(0, "/root/reference")
import logging
logging.disable(logging.CRITICAL)
from routing_board_generation.board_generation_methods.numpy_implementation.board_generation.bfs_board import BFSBoard

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
ok = 0
t0 = time.perf_counter()
for _ in range(n):
    try:
        b = BFSBoard(rows=10, cols=10, num_agents=5)
        b.return_solved_board()
        ok += 1
    except Exception:
        pass
dt = time.perf_counter() - t0
print(json.dumps({"generator": "numpy BFSBoard 10x10/5 (reference, single process)", "boards": n, "ok": ok, "seconds": round(dt, 3), "boards_per_sec": round(n / dt, 1), "cpu_count": os.cpu_count()}))
