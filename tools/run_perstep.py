"""A few per-step auto-reset calls in the stationary regime (for ncu launch lists): python tools/run_perstep.py [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import routing_board_generation_b200 as rbg
from routing_board_generation_b200 import engine
G, N, B = 10, 5, 65536
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 12
env = rbg.VmapAutoResetWrapper(rbg.Connector(generator=rbg.ParallelRandomWalkGenerator(G, N), time_limit=50))
st, _ = env.reset(rbg.split(rbg.PRNGKey(0), B))
st, _, _ = env.rollout_random(st, 160)  # stationary regime (launches before the ones of interest)
ts1 = engine.alloc_timestep(B, G, N)
for _ in range(steps):
    st, _, _ = engine.connector_step(st, None, 50, -0.03, 0.1, autoreset_kind="parallel_random_walk", inplace=True, random_policy=True, out=ts1, owner=env)
torch.cuda.synchronize()
print("ok")
