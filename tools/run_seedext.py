"""One SeedExtension batch (BASELINE configs[3]: 14x14 / 7 agents, 65 536 boards) for ncu launch lists."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import routing_board_generation_b200 as rbg
G, N, B = 14, 7, 65536
keys = rbg.split(rbg.PRNGKey(0), B)
board = rbg.SeedExtensionBoard(G, G, N)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    solved = board.return_solved_board(keys)
flags = rbg.engine.validate(solved, N)
torch.cuda.synchronize()
print("invalid", int((flags != 0).sum()))
