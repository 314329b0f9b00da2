import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import routing_board_generation_b200 as rbg
from oracle import oracle as orc
G, N = 14, 7
B = int(sys.argv[1])
kref = orc.split(orc.PRNGKey(0), 32)[1:1 + B] if B < 32 else orc.split(orc.PRNGKey(0), B)
_, st = orc.seedext_solved_batch(kref, G, N)
print("B", B, "sweeps max", st[:, 0].max(), "mean", st[:, 0].mean())
keys = rbg.engine.as_tensor(kref)
board = rbg.SeedExtensionBoard(G, G, N)
for _ in range(2):
    board.return_solved_board(keys)
torch.cuda.synchronize()
