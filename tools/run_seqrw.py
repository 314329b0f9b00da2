"""SequentialRandomWalkBoard.generate throughput on the device: python tools/run_seqrw.py [G N B]..."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import routing_board_generation_b200 as rbg  # noqa: E402

shapes = [(10, 5, 65536), (14, 7, 65536), (20, 10, 65536), (32, 16, 16384)]
if len(sys.argv) > 3:
    a = [int(x) for x in sys.argv[1:]]
    shapes = [tuple(a[i:i + 3]) for i in range(0, len(a), 3)]
for G, N, B in shapes:
    keys = rbg.split(rbg.PRNGKey(0), B)
    board = rbg.SequentialRandomWalkBoard(G, G, N)
    for _ in range(2):
        out, stats = board.generate_with_stats(keys)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    R = 20
    e0.record()
    for _ in range(R):
        out = board.generate(keys)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / R
    s = stats.cpu().numpy()
    print(f"seqrw {G}x{G}/{N} B={B}: {ms:.3f} ms  {B / ms / 1e3:.2f} M boards/s  failed {int((s[:, 0] == 0).sum())}  first-attempt {int((s[:, 0] == 1).sum())}  mean steps {s[:, 1].mean():.1f}", flush=True)
