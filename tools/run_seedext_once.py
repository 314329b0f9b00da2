"""One SeedExtension batch (14x14/7, 65 536 boards unless given) for ncu launch lists: python tools/run_seedext_once.py [G N B]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import routing_board_generation_b200 as rbg

a = [int(x) for x in sys.argv[1:]]
G, N, B = (a + [14, 7, 65536][len(a):])[:3]
keys = rbg.split(rbg.PRNGKey(0), B)
board = rbg.SeedExtensionBoard(G, G, N)
for _ in range(2):
    solved = board.return_solved_board(keys)
flags = rbg.engine.validate(solved, N)
torch.cuda.synchronize()
print("ok", int((flags != 0).sum()))
