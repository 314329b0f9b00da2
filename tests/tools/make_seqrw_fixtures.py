#!/usr/bin/env python
"""Golden fixtures for SequentialRandomWalkBoard / SequentialRandomWalkGenerator, made by running the
REFERENCE'S OWN source (imported unmodified from /root/reference) on the NumPy stand-in for jax (tests/tools/jax_shim).

    python tests/tools/make_seqrw_fixtures.py          # writes tests/golden/seqrw_reference.json

Two things stand between the reference as shipped and a run:
  * the class derives from the NumPy AbstractBoard without implementing its two abstract methods
    (sequential_random_walk.py:23, abstract_board.py:47-53), so SequentialRandomWalkBoard(...) raises TypeError.
    The script clears `__abstractmethods__` on the imported class (no source is touched); the method bodies then run
    as they are.
  * every cell is drawn with jax.random.choice(p=..., replace=False) (:57-63, :211-217), a float32 Gumbel top-k in
    jax 0.4.8.  With p in {0, 1} the pick is decided by the order of the candidates' uniforms alone as long as float32
    log is strictly monotone over the values involved.  The script checks that exhaustively for NumPy's log (all 2^23
    uniforms) and checks, on EVERY draw of every fixture, that the float formula's pick equals the integer rule the
    oracle and the CUDA kernel use (largest 23-bit mantissa among the candidates, lowest index on ties).  XLA's log cannot be
    run here (no jax in this image); the fixtures are exact for any log with that monotonicity.

What is stored: generate() boards (as int; the reference returns float32 codes), generate_starts_ends(), the
generator's States, over shapes from 3x3 to 10x10 including crowded boards that need several attempts and boards on
which every attempt fails (zero board; every pin at (0, 0)).
"""
from __future__ import annotations

import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SHIM = os.path.join(HERE, "jax_shim")
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden", "seqrw_reference.json")

sys.path.insert(0, ROOT)
sys.path.insert(0, SHIM)
sys.path.insert(0, REF)

import numpy as np  # noqa: E402


def check_log_monotone():
    k = np.arange(0, 1 << 23, dtype=np.uint32)
    u = (k | np.uint32(0x3F800000)).view(np.float32) - np.float32(1.0)
    tiny = np.finfo(np.float32).tiny
    u = np.maximum(tiny, u * np.float32(1.0 - tiny) + tiny).astype(np.float32)
    with np.errstate(divide="ignore"):
        g = (-np.log(-np.log(u))).astype(np.float32)
    d = np.diff(g)
    assert (d > 0).all(), "float32 gumbel(u) is not strictly increasing in u under NumPy's log"
    return True


def main():
    import jax
    import jax.random as jr
    from routing_board_generation.board_generation_methods.jax_implementation.board_generation.sequential_random_walk import SequentialRandomWalkBoard
    from routing_board_generation.rl_training.online_generators.sequential_random_walk_generator import SequentialRandomWalkGenerator
    from oracle import oracle as orc  # only to CHOOSE interesting keys (attempt counts); nothing of it is stored

    try:
        SequentialRandomWalkBoard(5, 5, 2)
        raise SystemExit("the reference class became instantiable: drop the __abstractmethods__ patch")
    except TypeError:
        pass
    SequentialRandomWalkBoard.__abstractmethods__ = frozenset()

    def L(x):
        return np.asarray(x).astype(np.int64).tolist()

    t0 = time.time()
    monotone = check_log_monotone()
    boards, pins, states = [], [], []
    # (G, N, seed, how many keys are scanned, wanted: dict attempt-class -> count)
    plan = [
        (3, 1, 11, 8, dict(first=2)), (3, 2, 12, 64, dict(first=2, retry=2)), (3, 3, 23, 256, dict(first=1, retry=2, fail=2)), (3, 4, 24, 64, dict(retry=2, fail=3)),
        (4, 2, 13, 8, dict(first=2)), (4, 4, 14, 256, dict(first=1, retry=3)), (4, 6, 25, 64, dict(retry=2, fail=3)),
        (5, 3, 15, 64, dict(first=2, retry=2)), (5, 6, 16, 256, dict(first=1, retry=3)), (5, 8, 26, 128, dict(retry=2, fail=2)), (6, 3, 17, 8, dict(first=3)), (6, 8, 18, 256, dict(retry=3)),
        (7, 4, 19, 8, dict(first=3)), (8, 5, 20, 64, dict(first=2, retry=1)), (10, 5, 21, 64, dict(first=3, retry=1)), (10, 12, 22, 128, dict(retry=2)),
    ]
    for (G, N, seed, scan, want) in plan:
        keys = np.asarray(jr.split(jr.PRNGKey(seed), scan))
        _, stats = orc.seqrw_generate_batch(keys, G, N)
        cls = np.where(stats[:, 0] == 0, 2, np.where(stats[:, 0] == 1, 0, 1))
        chosen = []
        for name, c in (("first", 0), ("retry", 1), ("fail", 2)):
            idx = np.nonzero(cls == c)[0][: want.get(name, 0)]
            chosen += [int(i) for i in idx]
        gen = SequentialRandomWalkBoard(G, G, N)
        for i in sorted(chosen):
            k = jax.numpy.array(keys[i])
            b = np.asarray(gen.generate(k))
            assert b.dtype == np.float32 and (b == np.round(b)).all()
            boards.append(dict(G=G, N=N, key=L(keys[i]), board=L(b), attempt=int(stats[i, 0])))
            s, e = gen.generate_starts_ends(k)
            pins.append(dict(G=G, N=N, key=L(keys[i]), starts=L(np.stack([np.asarray(x) for x in s])), ends=L(np.stack([np.asarray(x) for x in e]))))
        print(f"generate {G}x{G}/{N}: {len(chosen)} boards, attempts {[int(stats[i, 0]) for i in sorted(chosen)]}  ({time.time() - t0:.0f} s)", flush=True)
    for (G, N, seed, n) in ((5, 2, 31, 3), (6, 3, 32, 3), (4, 4, 33, 6), (8, 4, 34, 2), (3, 4, 35, 6), (4, 6, 36, 4)):
        g = SequentialRandomWalkGenerator(G, N)
        keys = np.asarray(jr.split(jr.PRNGKey(seed), n))
        for k in keys:
            st = g(jax.numpy.array(k))
            states.append(dict(G=G, N=N, key_in=L(k), key=L(st.key), grid=L(st.grid), step_count=int(st.step_count), agent_id=L(st.agents.id),
                               start=L(st.agents.start), target=L(st.agents.target), position=L(st.agents.position)))
        print(f"generator {G}x{G}/{N}: {n} States  ({time.time() - t0:.0f} s)", flush=True)

    # every Gumbel draw of the runs above: float formula == integer rule
    draws = adjacent = 0
    for mant, cand, pick in jr.GUMBEL_TRACE:
        if not cand.any():
            continue
        m = np.where(cand, mant.astype(np.int64), -1)
        assert int(np.argmax(m)) == pick, "float32 Gumbel pick differs from the integer rule"
        top = np.sort(m[cand])[::-1]
        if len(top) > 1 and top[0] - top[1] == 1:
            adjacent += 1
        draws += 1
    out = dict(
        _about="reference's own SequentialRandomWalkBoard / SequentialRandomWalkGenerator run on tests/tools/jax_shim by tests/tools/make_seqrw_fixtures.py",
        numpy_log_strictly_monotone_on_all_uniforms=monotone, gumbel_draws_checked=draws, draws_whose_top_two_uniforms_are_adjacent=adjacent,
        generate=boards, starts_ends=pins, generator_states=states)
    with open(OUT, "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print(f"wrote {OUT}: {len(boards)} boards, {len(states)} States, {draws} Gumbel draws checked ({adjacent} with adjacent top-two uniforms), {time.time() - t0:.0f} s")


if __name__ == "__main__":
    main()
