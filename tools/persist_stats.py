"""Counters of the persistent rollout's generator warps (needs a -DRBG_PERSIST_STATS build:
RBG_NVCC_EXTRA=-DRBG_PERSIST_STATS python routing-board-generation_b200/build.py --force)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import routing_board_generation_b200 as rbg  # noqa: E402

G, N, B, T = (int(x) for x in (sys.argv[1:5] + ["10", "5", "65536", "20"][len(sys.argv) - 1:]))
lib = rbg._lib.load()
env = rbg.VmapAutoResetWrapper(rbg.Connector(generator=rbg.ParallelRandomWalkGenerator(G, N), time_limit=50))
st, _ = env.reset(rbg.split(rbg.PRNGKey(0), B))
ts = rbg.engine.alloc_timestep(B, G, N, T)
for _ in range(8):
    st, _, _ = env.rollout_random(st, T, out=ts)
buf = (C.c_ulonglong * 16)()
lib.rbg_debug_persist_stats(buf, 1)
reps = 10
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    st, _, _ = env.rollout_random(st, T, out=ts)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
lib.rbg_debug_persist_stats(buf, 0)
v = [x / reps for x in buf]
print(f"{G}x{G}/{N} B={B} T={T}: {ms:.4f} ms per call, {B * T / ms / 1e6:.1f} M env-steps/s")
print(f"resets/call {v[7]:.0f}  batches {v[0]:.0f}  boards {v[1]:.0f}  boards/batch {v[1] / max(v[0], 1):.2f}  urgent batches {v[3]:.0f}")
print(f"hold-back polls {v[2]:.0f} ({v[2] * 0.5:.0f} us total)  gen idle polls {v[4]:.0f} ({v[4] * 0.2 / 1e3:.1f} ms total over all gen warps)")
print(f"cycles in batches {v[5]:.3g} = {v[5] / max(v[0], 1):.0f} per batch;  env wait polls {v[6]:.0f} ({v[6] * 0.1 / 1e3:.2f} ms total over all env warps)")
print(f"gen warps {v[8]:.0f}, cycles each ran on after its env warps finished: {v[9] / max(v[8], 1):.0f};  env warps {v[12]:.0f}, mean life {(v[14] - v[13]) / max(v[12], 1):.0f} cycles")
