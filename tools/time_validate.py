"""rbg_validate on generated boards: ms per batch (SeedExtension 14x14/7 x 262 144, ParallelRandomWalk 20x20/10 x 131 072)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import routing_board_generation_b200 as rbg
for (G, N, B, gen) in ((14, 7, 262144, "se"), (20, 10, 131072, "prw")):
    keys = rbg.split(rbg.PRNGKey(0), B)
    solved = rbg.SeedExtensionBoard(G, G, N).return_solved_board(keys) if gen == "se" else rbg.ParallelRandomWalkBoard(G, G, N).generate_board(keys)[2]
    for _ in range(3):
        f = rbg.engine.validate(solved, N)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        f = rbg.engine.validate(solved, N)
    e1.record()
    torch.cuda.synchronize()
    print(gen, G, N, B, "validate %.3f ms" % (e0.elapsed_time(e1) / 10), "flags", int(f.abs().max()))
