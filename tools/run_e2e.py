"""Time rbg_connector_step_host_io alone (bench.py's e2e leg): pinned host actions in, whole TimeStep out."""
import os, sys, argparse
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import routing_board_generation_b200 as rbg

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
env = rbg.VmapAutoResetWrapper(rbg.Connector(generator=rbg.ParallelRandomWalkGenerator(bench.G, bench.N), time_limit=bench.TIME_LIMIT))
state, _ = env.reset(rbg.split(rbg.PRNGKey(0), B))
args = argparse.Namespace(steps=30)
for _ in range(3):
    r = bench._e2e_host(args, rbg, rbg._lib.load(), state, B, 1, 0, torch.device("cuda", 0))
    print(r["value"] / 1e6, "M env-steps/s", r["d2h_bytes_per_step"] / 1e6, "MB D2H per step ->", r["d2h_bytes_per_step"] * r["value"] / B / 1e9, "GB/s")
