"""Time bulk ParallelRandomWalk generation (prw_kernel) at the BASELINE shapes; also used for ncu captures.

    python tools/run_prw.py [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import routing_board_generation_b200 as rbg

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
for (G, N, B) in ((10, 5, 65536), (20, 10, 131072), (32, 16, 32768)):
    keys = rbg.split(rbg.PRNGKey(0), 1048576, 0, B)
    board = rbg.ParallelRandomWalkBoard(G, G, N)
    for _ in range(3):
        heads, targets, solved = board.generate_board(keys)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        heads, targets, solved = board.generate_board(keys)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"prw {G}x{G}/{N} B={B}: {ms:.4f} ms  {B / ms / 1e3:.1f} M boards/s  checksum {int(solved.sum())}")
