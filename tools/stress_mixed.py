"""Randomised API stress: several env batches of a few shapes and all generator kinds alive at the same
time, stepped in random order through step / step_random / rollout_random, created and dropped at
random (so workspaces change hands), every result checked against the oracle.

    python tools/stress_mixed.py [operations] [seed]"""
import gc, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import routing_board_generation_b200 as rbg
from oracle import oracle as orc
import test_gpu_parity as T

ops = int(sys.argv[1]) if len(sys.argv) > 1 else 300
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
GENS = {"parallel_random_walk": rbg.ParallelRandomWalkGenerator, "uniform": rbg.UniformRandomGenerator, "seed_extension": rbg.SeedExtensionGenerator,
        "sequential_random_walk": rbg.SequentialRandomWalkGenerator}
SHAPES = [(10, 5, 700), (10, 5, 4500), (8, 4, 300), (6, 3, 1000)]
live = []
made = 0


def make():
    global made
    kind = list(GENS)[int(rng.integers(0, len(GENS)))]
    G, N, B = SHAPES[int(rng.integers(0, len(SHAPES)))]
    tl = int(rng.integers(1, 6))
    env = rbg.VmapAutoResetWrapper(rbg.Connector(generator=GENS[kind](G, N), time_limit=tl))
    keys, kref = T._keys(rbg, orc, 5000 + made, B)
    made += 1
    st, ts = env.reset(keys)
    rst, rts = orc.connector_reset_batch(kind, kref, G, N)
    T._assert_state(st, rst, "reset")
    T._assert_timestep(ts, rts, "reset")
    return dict(env=env, kind=kind, G=G, N=N, B=B, tl=tl, st=st, rst=rst, n=0)


def advance(b):
    op = int(rng.integers(0, 3))
    where = f"{b['kind']} {b['G']}x{b['G']}/{b['N']} B={b['B']} tl={b['tl']} after {b['n']} steps, op {op}"
    if op == 0:  # rollout
        Tn = int(rng.integers(1, 9))
        b["st"], ts, act = b["env"].rollout_random(b["st"], Tn)
        for t in range(Tn):
            a = orc.random_actions_batch(b["rst"])
            assert np.array_equal(T._np(act[t]), a), where
            b["rst"], rts = orc.connector_step_batch(b["rst"], a, time_limit=b["tl"], autoreset_kind=b["kind"])
            T._assert_timestep(ts[t], rts, where)
        b["n"] += Tn
    elif op == 1:  # fused random step
        a = orc.random_actions_batch(b["rst"])
        b["st"], ts, act = b["env"].step_random(b["st"])
        assert np.array_equal(T._np(act), a), where
        b["rst"], rts = orc.connector_step_batch(b["rst"], a, time_limit=b["tl"], autoreset_kind=b["kind"])
        T._assert_timestep(ts, rts, where)
        b["n"] += 1
    else:  # caller's actions, any codes
        a = rng.integers(-1, 7, size=(b["B"], b["N"])).astype(np.int32)
        b["st"], ts = b["env"].step(b["st"], torch.from_numpy(a).cuda())
        b["rst"], rts = orc.connector_step_batch(b["rst"], a, time_limit=b["tl"], autoreset_kind=b["kind"])
        T._assert_timestep(ts, rts, where)
        b["n"] += 1
    T._assert_state(b["st"], b["rst"], where)


for i in range(ops):
    r = rng.random()
    if not live or (r < 0.15 and len(live) < 5):
        live.append(make())
    elif r < 0.25 and len(live) > 1:
        live.pop(int(rng.integers(0, len(live))))
        gc.collect()
    else:
        advance(live[int(rng.integers(0, len(live)))])
print("ok", ops, "operations,", made, "env batches, workspaces alive", len(rbg.engine._workspaces))
