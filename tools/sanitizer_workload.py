"""A small pass over every kernel for compute-sanitizer (memcheck / racecheck): generators, step,
auto-reset (cached and synchronous), fused rollout with in-kernel generation, SeedExtension, validate."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import routing_board_generation_b200 as rbg

B = 96
keys = rbg.split(rbg.PRNGKey(0), B)
for (G, N) in ((10, 5), (5, 3), (9, 9)):
    h, t, s = rbg.ParallelRandomWalkBoard(G, G, N).generate_board(keys)
    rbg.engine.validate(s, N)
    for kind, gen in (("prw", rbg.ParallelRandomWalkGenerator(G, N)), ("uni", rbg.UniformRandomGenerator(G, N))):
        env = rbg.VmapAutoResetWrapper(rbg.Connector(generator=gen, time_limit=3))
        st, ts = env.reset(keys)
        for _ in range(5):
            st, ts, act = env.step_random(st, inplace=True)
        st, ts, act = env.rollout_random(st, 9)
solved = rbg.SeedExtensionBoard(8, 8, 4).return_solved_board(keys)
rbg.engine.validate(solved, 4)
st = rbg.SeedExtensionGenerator(8, 4)(keys)
ds = rbg.BoardDatasetGeneratorJAX(6, 4, board_name="offline_parallel_rw", number_of_boards=7)(keys)
torch.cuda.synchronize()
print("sanitizer workload done")
