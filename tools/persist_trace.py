"""Per-warp start / end stamps of one rollout_persist_kernel launch (needs a -DRBG_PERSIST_TRACE build)."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import routing_board_generation_b200 as rbg  # noqa: E402

G, N, B, T = 10, 5, 65536, 20
lib = rbg._lib.load()
env = rbg.VmapAutoResetWrapper(rbg.Connector(generator=(rbg.UniformRandomGenerator if os.environ.get("KIND") == "uniform" else rbg.ParallelRandomWalkGenerator)(G, N), time_limit=50))
st, _ = env.reset(rbg.split(rbg.PRNGKey(0), B))
ts = rbg.engine.alloc_timestep(B, G, N, T)
for _ in range(12):
    st, _, _ = env.rollout_random(st, T, out=ts)
torch.cuda.synchronize()
n = 592 * 6
buf = (C.c_ulonglong * (3 * n))()
lib.rbg_debug_persist_trace(buf, 3 * n)
a = np.array(buf[:], dtype=np.int64).reshape(n, 3)
a = a[a[:, 0] > 0]
t0 = a[:, 0].min()
env_w, gen_w = a[a[:, 2] < 1000000], a[a[:, 2] >= 1000000]
pct = lambda x: np.percentile(x, [0, 5, 25, 50, 75, 95, 100]).round(1).tolist()  # noqa: E731
print("kernel span us:", (a[:, 1].max() - t0) / 1e3)
print("env warps:", len(env_w), "start us pct", pct((env_w[:, 0] - t0) / 1e3), "end us pct", pct((env_w[:, 1] - t0) / 1e3))
print("groups per env warp:", np.bincount(env_w[:, 2]).tolist())
print("gen warps:", len(gen_w), "end us pct", pct((gen_w[:, 1] - t0) / 1e3))
for g in sorted(set(env_w[:, 2].tolist())):
    sel = env_w[env_w[:, 2] == g]
    print(f"  warps with {g} groups: {len(sel)}, mean us per group {(sel[:, 1] - sel[:, 0]).mean() / 1e3 / max(g, 1):.1f}, end pct {pct((sel[:, 1] - t0) / 1e3)}")
