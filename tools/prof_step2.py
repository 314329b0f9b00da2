import os, sys, time, argparse
sys.path.insert(0, os.getcwd())
import torch
import bench
import routing_board_generation_b200 as rbg
from routing_board_generation_b200 import engine
G, N, B = 10, 5, 65536
def per_step(tag, warm=160):
    env = rbg.VmapAutoResetWrapper(rbg.Connector(generator=rbg.ParallelRandomWalkGenerator(G, N), time_limit=50))
    st, _ = env.reset(rbg.split(rbg.PRNGKey(0), B))
    ts1 = engine.alloc_timestep(B, G, N)
    for _ in range(warm):
        st, _, _ = engine.connector_step(st, None, 50, -0.03, 0.1, autoreset_kind="parallel_random_walk", inplace=True, random_policy=True, out=ts1, owner=env)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t = time.perf_counter(); e0.record()
    for _ in range(200):
        st, _, _ = engine.connector_step(st, None, 50, -0.03, 0.1, autoreset_kind="parallel_random_walk", inplace=True, random_policy=True, out=ts1, owner=env)
    e1.record(); ti = time.perf_counter() - t; torch.cuda.synchronize()
    rbg._lib.kernel_timing(True)
    for k in ("env", "prw"): rbg._lib.kernel_time(k)
    for _ in range(100):
        st, _, _ = engine.connector_step(st, None, 50, -0.03, 0.1, autoreset_kind="parallel_random_walk", inplace=True, random_policy=True, out=ts1, owner=env)
    torch.cuda.synchronize()
    ne, me = rbg._lib.kernel_time("env"); npw, mp = rbg._lib.kernel_time("prw")
    rbg._lib.kernel_timing(False)
    print(tag, f"events {e0.elapsed_time(e1) / 200 * 1e3:.1f} us/step, host issue {ti / 200 * 1e6:.1f} us/call | env kernel {me / ne * 1e3:.1f} us x{ne}, prw {mp / npw * 1e3:.1f} us x{npw}, state ptr {st.grid.data_ptr() % (1 << 30) >> 20} MB")
per_step("fresh")
# what bench does before its secondary lines
env = rbg.VmapAutoResetWrapper(rbg.Connector(generator=rbg.ParallelRandomWalkGenerator(G, N), time_limit=50))
state, _ = env.reset(rbg.split(rbg.PRNGKey(0), B))
ts = engine.alloc_timestep(B, G, N, 20); act = torch.empty((20, B, N), dtype=torch.int32, device="cuda")
for _ in range(30):
    engine.connector_rollout_random(state, 20, 50, -0.03, 0.1, autoreset_kind="parallel_random_walk", out=ts, actions=act)
torch.cuda.synchronize()
per_step("after rollouts")
bench._e2e_host(argparse.Namespace(steps=10), rbg, rbg._lib.load(), state, B, 1, 0, torch.device("cuda", 0))
per_step("after e2e")
bench._e2e_host(argparse.Namespace(steps=10), rbg, rbg._lib.load(), state, B, 1, 0, torch.device("cuda", 0))
per_step("after e2e, warm 1000", 1000)
bench._e2e_host(argparse.Namespace(steps=10), rbg, rbg._lib.load(), state, B, 1, 0, torch.device("cuda", 0))
per_step("after e2e, warm 3000", 3000)
