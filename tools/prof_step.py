import os, sys, time, cProfile, pstats
sys.path.insert(0, os.getcwd())
import torch
import routing_board_generation_b200 as rbg
from routing_board_generation_b200 import engine
G, N, B = 10, 5, 65536
env = rbg.VmapAutoResetWrapper(rbg.Connector(generator=rbg.ParallelRandomWalkGenerator(G, N), time_limit=50))
st, _ = env.reset(rbg.split(rbg.PRNGKey(0), B))
ts1 = engine.alloc_timestep(B, G, N)
def step():
    global st
    st, _, _ = engine.connector_step(st, None, 50, -0.03, 0.1, autoreset_kind="parallel_random_walk", inplace=True, random_policy=True, out=ts1, owner=env)
for _ in range(200): step()
torch.cuda.synchronize()
t = time.perf_counter()
for _ in range(300): step()
t_issue = time.perf_counter() - t
torch.cuda.synchronize()
t_all = time.perf_counter() - t
print(f"host issue {t_issue / 300 * 1e6:.1f} us/call, with sync {t_all / 300 * 1e6:.1f} us/call")
pr = cProfile.Profile(); pr.enable()
for _ in range(300): step()
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
