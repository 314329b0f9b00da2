"""Stress the step-wise auto-reset path (speculative next-episode cache + synchronous reset) in one
process: many short env batches of the same shape, so that workspaces and their contexts are reused
by later batches, every env terminating every 1-3 steps.  Checks every step against the oracle.

    python tools/stress_autoreset.py [rounds]"""
import gc, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import routing_board_generation_b200 as rbg
from oracle import oracle as orc
import test_gpu_parity as T

rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 10
n = 0
for r in range(rounds):
    for time_limit in (1, 2, 3):
        for kind in ("parallel_random_walk", "uniform", "seed_extension"):
            try:
                T._rollout(rbg, orc, kind, 10, 5, B=700, steps=12, autoreset=True, time_limit=time_limit, seed=17 + time_limit + 100 * r)
            except AssertionError as e:
                print("FAIL round", r, "time_limit", time_limit, kind, str(e)[:300])
                # diagnose: first step of the same batch again, env by env
                import torch
                seed = 17 + time_limit + 100 * r
                keys, kref = T._keys(rbg, orc, seed, 700)
                gen = {"parallel_random_walk": rbg.ParallelRandomWalkGenerator, "uniform": rbg.UniformRandomGenerator, "seed_extension": rbg.SeedExtensionGenerator}[kind](10, 5)
                for attempt in range(3):
                    env = rbg.Connector(generator=gen, time_limit=time_limit)
                    st, ts = env.reset(keys)
                    rst, rts = orc.connector_reset_batch(kind, kref, 10, 5)
                    act = orc.random_actions_batch(rst)
                    st, ts = rbg.VmapAutoResetWrapper(env).step(st, torch.from_numpy(act).cuda())
                    rst, rts = orc.connector_step_batch(rst, act, time_limit=time_limit, autoreset_kind=kind)
                    g = st.grid.cpu().numpy()
                    bad = np.nonzero((g != rst["grid"]).reshape(700, -1).any(axis=1))[0]
                    kb = np.nonzero((st.key.cpu().numpy() != rst["key"]).any(axis=1))[0]
                    print(" attempt", attempt, "envs with wrong grid", len(bad), bad[:20], "wrong key", len(kb), "workspaces", len(rbg.engine._workspaces))
                sys.exit(1)
            n += 1
    gc.collect()
print("ok", n, "batches")
