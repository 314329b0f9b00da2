"""Randomised soak of SequentialRandomWalk against the oracle: random (G, N) incl. crowded boards (many retries, resumed attempts,
failed generations), both lane layouts of the kernel, generator States.  python tools/stress_seqrw.py [seconds] [seed]"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = r'''
import sys, time, numpy as np
sys.path.insert(0, %r)
import routing_board_generation_b200 as rbg
from oracle import oracle as orc
budget, seed = float(sys.argv[1]), int(sys.argv[2])
rng = np.random.default_rng(seed)
t0 = time.time(); n = 0; boards = 0; retried = 0; failed = 0
while time.time() - t0 < budget:
    G = int(rng.integers(3, 41))
    dense = rng.random() < 0.5
    N = int(rng.integers(1, min(32, max(2, G * G // (3 if dense else 12))) + 1))
    B = int(max(16, min(4000, 400000 // (G * G * max(1, N // 2)))))
    s = int(rng.integers(0, 1 << 30))
    k = rbg.split(rbg.PRNGKey(s), B); kr = orc.split(orc.PRNGKey(s), B)
    b, st = rbg.SequentialRandomWalkBoard(G, G, N).generate_with_stats(k)
    rb, rs = orc.seqrw_generate_batch(kr, G, N)
    assert np.array_equal(st.cpu().numpy(), rs), ("stats", G, N, s)
    assert np.array_equal(b.cpu().numpy(), rb), ("board", G, N, s)
    if n %% 4 == 0:
        g = rbg.SequentialRandomWalkGenerator(G, N)(k[:64])
        ref = orc.state_batch("sequential_random_walk", kr[:64], G, N)
        assert np.array_equal(g.grid.cpu().numpy(), ref["grid"]) and np.array_equal(g.agents.target.cpu().numpy(), ref["target"]), ("state", G, N, s)
    n += 1; boards += B; retried += int((rs[:, 0] > 1).sum()); failed += int((rs[:, 0] == 0).sum())
print(f"ok: {n} shapes, {boards} boards, {retried} retried, {failed} failed generations, lanes {sys.argv[3]}")
''' % ROOT
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
for lanes in ("", "6", "8"):
    env = dict(os.environ)
    if lanes:
        env["RBG_SEQRW_W"] = lanes
    r = subprocess.run([sys.executable, "-c", CODE, str(budget / 3), str(seed), lanes or "default"], env=env, capture_output=True, text=True, timeout=budget + 300)
    print(r.stdout.strip() or r.stderr[-1500:], flush=True)
    if r.returncode != 0:
        sys.exit(1)
