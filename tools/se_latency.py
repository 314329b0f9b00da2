"""Per-sweep latency of se_extend_kernel: one warp (32 boards), time / max sweeps (sweep counts from the oracle)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import routing_board_generation_b200 as rbg
from oracle import oracle as orc
G, N = 14, 7
for B in (1, 2, 4, 8, 32):
    kref = orc.split(orc.PRNGKey(0), 32)[1:1 + B] if B < 32 else orc.split(orc.PRNGKey(0), B)
    _, st = orc.seedext_solved_batch(kref, G, N)
    keys = rbg.engine.as_tensor(kref)
    board = rbg.SeedExtensionBoard(G, G, N)
    for _ in range(3):
        board.return_solved_board(keys)
    torch.cuda.synchronize()
    rbg._lib.kernel_timing(True); rbg._lib.kernel_time("seedext")
    reps = 10
    for _ in range(reps):
        board.return_solved_board(keys)
    torch.cuda.synchronize()
    n, ms = rbg._lib.kernel_time("seedext")
    rbg._lib.kernel_timing(False)
    print(f"B={B}: sweeps max {st[:,0].max()} mean {st[:,0].mean():.1f}; seedext kernels {ms/reps:.3f} ms per batch ({n//reps} launches); {ms/reps*1e3/st[:,0].max():.1f} us per sweep of the slowest board (incl. seed/optimise/finish)")
