"""Time SeedExtension batches (BASELINE configs[3]: 14x14 / 7 agents, 65 536 boards, + rbg_validate) and other
shapes; per-kernel time from the library's event pairs.  Also used for ncu launch lists.

    python tools/run_seedext.py [reps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import routing_board_generation_b200 as rbg

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
for (G, N, B) in ((14, 7, 65536), (10, 5, 65536), (20, 10, 32768)):
    keys = rbg.split(rbg.PRNGKey(0), B)
    board = rbg.SeedExtensionBoard(G, G, N)
    for _ in range(2):
        solved = board.return_solved_board(keys)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        solved = board.return_solved_board(keys)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    flags = rbg.engine.validate(solved, N)
    print(f"seedext {G}x{G}/{N} B={B}: {ms:.3f} ms  {B / ms / 1e3:.2f} M boards/s  invalid {int((flags != 0).sum())}  checksum {int(solved.sum())}")
