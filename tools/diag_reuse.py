"""A workspace shared by env batches of different generator kinds (what an id()-keyed cache did when
Python reused an object id): speculative (uniform / parallel_random_walk) -> non-speculative
(seed_extension) -> speculative.  Every step is checked against the oracle."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import routing_board_generation_b200 as rbg
from oracle import oracle as orc
import test_gpu_parity as T

GENS = {"parallel_random_walk": rbg.ParallelRandomWalkGenerator, "uniform": rbg.UniformRandomGenerator, "seed_extension": rbg.SeedExtensionGenerator}

def run(env, seed, tl, kind, steps):
    keys, kref = T._keys(rbg, orc, seed, 700)
    st, ts = env.reset(keys)
    rst, rts = orc.connector_reset_batch(kind, kref, 10, 5)
    for t in range(steps):
        act = orc.random_actions_batch(rst)
        st, ts = rbg.VmapAutoResetWrapper(env).step(st, torch.from_numpy(act).cuda())
        rst, rts = orc.connector_step_batch(rst, act, time_limit=tl, autoreset_kind=kind)
        bad = np.nonzero((st.grid.cpu().numpy() != rst["grid"]).reshape(700, -1).any(axis=1))[0]
        if len(bad):
            return f"{kind} step {t}: {len(bad)} wrong envs"
    return None

fails = 0
for rep in range(10):
    first = None
    for i, (kind, tl) in enumerate((("uniform", 1), ("seed_extension", 1), ("uniform", 1), ("parallel_random_walk", 2), ("seed_extension", 2), ("parallel_random_walk", 1))):
        env = rbg.Connector(generator=GENS[kind](10, 5), time_limit=tl)
        if first is None:
            first = env
            rbg.VmapAutoResetWrapper(env)  # noqa
        else:
            env._rbg_ws_token = first._rbg_ws_token if hasattr(first, "_rbg_ws_token") else None
        err = run(env, 3000 + 10 * rep + i, tl, kind, 4)
        if i == 0 and not hasattr(first, "_rbg_ws_token"):
            raise SystemExit("no token")
        if err:
            print("rep", rep, "batch", i, err)
            fails += 1
print("failures", fails)
