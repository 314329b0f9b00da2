"""Per-step API regime (65 / 85 us per step) on the default stream vs on a private non-default stream."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import routing_board_generation_b200 as rbg
from routing_board_generation_b200 import engine
G, N, B = 10, 5, 65536
def per_step(tag):
    env = rbg.VmapAutoResetWrapper(rbg.Connector(generator=rbg.ParallelRandomWalkGenerator(G, N), time_limit=50))
    st, _ = env.reset(rbg.split(rbg.PRNGKey(0), B))
    ts1 = engine.alloc_timestep(B, G, N)
    for _ in range(160):
        st, _, _ = engine.connector_step(st, None, 50, -0.03, 0.1, autoreset_kind="parallel_random_walk", inplace=True, random_policy=True, out=ts1, owner=env)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200):
        st, _, _ = engine.connector_step(st, None, 50, -0.03, 0.1, autoreset_kind="parallel_random_walk", inplace=True, random_policy=True, out=ts1, owner=env)
    e1.record(); torch.cuda.synchronize()
    print(tag, f"{e0.elapsed_time(e1) / 200 * 1e3:.1f} us/step")
for rep in range(4):
    per_step("default stream")
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        per_step("private stream")
    torch.cuda.synchronize()
