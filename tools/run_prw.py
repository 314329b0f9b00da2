"""One bulk ParallelRandomWalk generation (BASELINE configs[2] per-GPU slice: 20x20 / 10 agents, 131 072 boards)
for ncu captures of prw_kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import routing_board_generation_b200 as rbg
G, N, B = 20, 10, 131072
keys = rbg.split(rbg.PRNGKey(0), 1048576, 0, B)
board = rbg.ParallelRandomWalkBoard(G, G, N)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    heads, targets, solved = board.generate_board(keys)
torch.cuda.synchronize()
print("ok", int(solved.sum()))
