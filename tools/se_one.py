import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import routing_board_generation_b200 as rbg
G, N, B = 14, 7, 65536
keys = rbg.split(rbg.PRNGKey(0), B)
board = rbg.SeedExtensionBoard(G, G, N)
for _ in range(2):
    solved = board.return_solved_board(keys)
torch.cuda.synchronize()
