#!/usr/bin/env python
"""Generate golden fixtures by running the REFERENCE'S OWN Python source (imported unmodified from
/root/reference) on the NumPy stand-in for jax in tests/tools/jax_shim (neither jax nor jumanji is
installed here; see that directory's README).

    python tests/tools/make_reference_fixtures.py            # writes tests/golden/reference_runs.json

What is pinned this way (none of it has a golden in the reference itself):
  * ParallelRandomWalkBoard.generate_board incl. the collision / revert branch (parallel_random_walk.py:101-145)
  * ParallelRandomWalkGenerator / UniformRandomGenerator / SeedExtensionGenerator States
  * SeedExtensionBoard.return_seeded_board / return_solved_board / generate_starts_ends with the
    default and the non-default options; extend_wires_jax and optimise_wire in isolation
  * BoardDatasetGeneratorJAX: the stored boards and the randint pick of __call__
    (`python tests/tools/make_reference_fixtures.py dataset` regenerates only this section)
  * the BASELINE shapes: generate_board at 20x20/10, 24x24/12, 32x32/16, 40x40/32 (`... large`) and
    SeedExtension at 14x14/7, 20x20/10 (`... seedext_large`), appended to the lists above
The shim's jax.random is checked first against the reference-owned goldens by running the reference's
own test file (test_parallel_random_walk_board.py, 35 tests) under it.
"""
from __future__ import annotations

import json
import os
import subprocess
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SHIM = os.path.join(HERE, "jax_shim")
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden", "reference_runs.json")
REF_TEST = os.path.join(REF, "routing_board_generation/board_generation_methods/jax_implementation/board_generation/test_parallel_random_walk_board.py")


def run_reference_tests() -> str:
    env = dict(os.environ, PYTHONPATH=SHIM + os.pathsep + REF)
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-p", "no:cacheprovider", REF_TEST], cwd="/tmp", env=env, capture_output=True, text=True)
    tail = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-300:]
    if r.returncode != 0:
        raise SystemExit(f"the reference's own tests fail under the shim:\n{r.stdout[-3000:]}")
    return tail


def dataset_section():
    """BoardDatasetGeneratorJAX (dataset_generator_jax.py:20-141) -> merged into the existing JSON."""
    sys.path.insert(0, SHIM)
    sys.path.insert(0, REF)
    import jax
    import numpy as np
    from routing_board_generation.rl_training.offline_generation.dataset_generator_jax import BoardDatasetGeneratorJAX

    def L(x):
        return np.asarray(x).astype(np.int64).tolist()

    out = []
    for (G, N, K, name, seed, n) in ((6, 4, 7, "offline_parallel_rw", 51, 24), (8, 4, 5, "offline_seed_extension", 52, 16)):
        gen = BoardDatasetGeneratorJAX(G, N, board_name=name, number_of_boards=K)
        keys = np.asarray(jax.random.split(jax.random.PRNGKey(seed), n))
        states = []
        for k in keys:
            st = gen(jax.numpy.array(k))
            states.append(dict(key=L(st.key), grid=L(st.grid), start=L(st.agents.start), target=L(st.agents.target)))
        out.append(dict(G=G, N=N, K=K, board_name=name, seed=seed, n=n, heads=L(gen.heads), targets=L(gen.targets), states=states))
        print("dataset generator", name, G, N, K)
    with open(OUT) as f:
        data = json.load(f)
    data["dataset_generator"] = out
    with open(OUT, "w") as f:
        json.dump(data, f, separators=(",", ":"))
    print("merged into", OUT)


def large_section():
    """ParallelRandomWalkBoard.generate_board at the BASELINE shapes 20x20/10 and 32x32/16 (slow under the
    shim: minutes per board) -> appended to the prw_generate_board list of the existing JSON."""
    sys.path.insert(0, SHIM)
    sys.path.insert(0, REF)
    import jax
    import numpy as np
    from routing_board_generation.board_generation_methods.jax_implementation.board_generation.parallel_random_walk import ParallelRandomWalkBoard

    def L(x):
        return np.asarray(x).astype(np.int64).tolist()

    with open(OUT) as f:
        data = json.load(f)
    have = {(e["G"], e["N"], e["seed"]) for e in data["prw_generate_board"]}
    t0 = time.time()
    for (G, N, seed, n) in ((20, 10, 61, 3), (32, 16, 62, 2), (20, 10, 63, 12), (32, 16, 64, 6), (40, 32, 65, 2), (24, 12, 66, 4)):
        if (G, N, seed) in have:
            continue
        board = ParallelRandomWalkBoard(G, G, N)
        ks = np.asarray(jax.random.split(jax.random.PRNGKey(seed), n))
        rows = []
        for k in ks:
            heads, targets, solved = board.generate_board(jax.numpy.array(k))
            rows.append(dict(heads=L(heads), targets=L(targets), solved=L(solved)))
            print(f"prw {G}x{G}/{N}: board {len(rows)}/{n}  [{time.time() - t0:.0f}s]", flush=True)
        data["prw_generate_board"].append(dict(G=G, N=N, seed=seed, n=n, boards=rows))
        with open(OUT, "w") as f:
            json.dump(data, f, separators=(",", ":"))
    print("merged into", OUT)


def seedext_large_section():
    """SeedExtensionBoard.return_solved_board / generate_starts_ends at 14x14/7 (BASELINE configs[3]) and
    20x20/10 -> appended to the seedext_solved list of the existing JSON."""
    sys.path.insert(0, SHIM)
    sys.path.insert(0, REF)
    import jax
    import numpy as np
    from routing_board_generation.board_generation_methods.jax_implementation.board_generation.seed_extension import SeedExtensionBoard

    def L(x):
        return np.asarray(x).astype(np.int64).tolist()

    with open(OUT) as f:
        data = json.load(f)
    have = {(e["G"], e["N"], e["seed"]) for e in data["seedext_solved"]}
    t0 = time.time()
    for (G, N, seed, n) in ((14, 7, 71, 10), (20, 10, 72, 3)):
        if (G, N, seed) in have:
            continue
        board = SeedExtensionBoard(G, G, N)
        ks = np.asarray(jax.random.split(jax.random.PRNGKey(seed), n))
        rows = []
        for k in ks:
            k = jax.numpy.array(k)
            solved = board.return_solved_board(k)
            (sr, sc), (er, ec) = board.generate_starts_ends(k)
            rows.append(dict(solved=L(solved), starts=[L(sr), L(sc)], ends=[L(er), L(ec)]))
            print(f"seed extension {G}x{G}/{N}: board {len(rows)}/{n}  [{time.time() - t0:.0f}s]", flush=True)
        data["seedext_solved"].append(dict(G=G, N=N, seed=seed, n=n, options={}, boards=rows))
        with open(OUT, "w") as f:
            json.dump(data, f, separators=(",", ":"))
    print("merged into", OUT)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "seedext_large":
        return seedext_large_section()
    if len(sys.argv) > 1 and sys.argv[1] == "dataset":
        return dataset_section()
    if len(sys.argv) > 1 and sys.argv[1] == "large":
        return large_section()
    print("reference tests under the shim:", run_reference_tests())
    sys.path.insert(0, SHIM)
    sys.path.insert(0, REF)
    import jax
    import numpy as np
    from routing_board_generation.board_generation_methods.jax_implementation.board_generation.parallel_random_walk import ParallelRandomWalkBoard
    from routing_board_generation.board_generation_methods.jax_implementation.board_generation.seed_extension import SeedExtensionBoard
    from routing_board_generation.board_generation_methods.jax_implementation.utils.grid_utils import optimise_wire
    from routing_board_generation.board_generation_methods.jax_implementation.utils.post_processor_utils_jax import extend_wires_jax
    from routing_board_generation.rl_training.online_generators.parallel_random_walk_generator import ParallelRandomWalkGenerator
    from routing_board_generation.rl_training.online_generators.random_seed_generator import SeedExtensionGenerator
    from routing_board_generation.rl_training.online_generators.uniform_generator import UniformRandomGenerator

    def L(x):
        return np.asarray(x).astype(np.int64).tolist()

    def keys_of(seed, n):
        return np.asarray(jax.random.split(jax.random.PRNGKey(seed), n))

    def state_dict(st):
        a = st.agents
        return dict(key=L(st.key), grid=L(st.grid), step_count=int(st.step_count), id=L(a.id), start=L(a.start), target=L(a.target), position=L(a.position))

    out = {"_about": "Outputs of the reference's own Python files run on tests/tools/jax_shim (NumPy stand-in for jax==0.4.8); made by tests/tools/make_reference_fixtures.py. keys = jax.random.split(PRNGKey(seed), n)."}
    t0 = time.time()

    # ---- ParallelRandomWalkBoard.generate_board (dense configs collide often)
    prw = []
    for (G, N, seed, n) in ((5, 3, 1, 48), (6, 6, 2, 48), (8, 8, 3, 32), (10, 5, 4, 48), (7, 12, 5, 24), (14, 7, 6, 8)):
        board = ParallelRandomWalkBoard(G, G, N)
        ks = keys_of(seed, n)
        rows = []
        for k in ks:
            heads, targets, solved = board.generate_board(jax.numpy.array(k))
            rows.append(dict(heads=L(heads), targets=L(targets), solved=L(solved)))
        prw.append(dict(G=G, N=N, seed=seed, n=n, boards=rows))
        print(f"prw {G}x{G}/{N}: {n} boards  [{time.time() - t0:.0f}s]")
    out["prw_generate_board"] = prw

    # ---- generators -> State
    gens = []
    for name, cls, cfgs in (("parallel_random_walk", ParallelRandomWalkGenerator, ((10, 5, 7, 12), (6, 4, 8, 12))),
                            ("uniform", UniformRandomGenerator, ((10, 5, 9, 16), (5, 3, 10, 16))),
                            ("seed_extension", SeedExtensionGenerator, ((10, 5, 11, 6), (6, 3, 12, 8)))):
        for (G, N, seed, n) in cfgs:
            gen = cls(G, N)
            ks = keys_of(seed, n)
            gens.append(dict(kind=name, G=G, N=N, seed=seed, n=n, states=[state_dict(gen(jax.numpy.array(k))) for k in ks]))
            print(f"generator {name} {G}x{G}/{N}: {n} states  [{time.time() - t0:.0f}s]")
    out["generator_states"] = gens

    # ---- SeedExtensionBoard
    se = []
    for (G, N, seed, n, kw) in ((10, 5, 21, 12, {}), (6, 3, 22, 16, {}), (8, 4, 23, 12, {}), (5, 4, 24, 12, {}), (14, 7, 25, 3, {}),
                                (8, 4, 26, 6, dict(randomness=0.5)), (8, 4, 27, 6, dict(randomness=1.0)), (8, 4, 28, 6, dict(two_sided=False)),
                                (8, 4, 29, 6, dict(extension_iterations=2)), (8, 4, 30, 6, dict(extension_steps=2)),
                                (7, 5, 31, 6, dict(randomness=0.3, two_sided=False, extension_iterations=2, extension_steps=3))):
        board = SeedExtensionBoard(G, G, N)
        ks = keys_of(seed, n)
        rows = []
        for k in ks:
            k = jax.numpy.array(k)
            solved = board.return_solved_board(k, **kw)
            (sr, sc), (er, ec) = board.generate_starts_ends(k, **kw)
            rows.append(dict(solved=L(solved), starts=[L(sr), L(sc)], ends=[L(er), L(ec)]))
        se.append(dict(G=G, N=N, seed=seed, n=n, options=kw, boards=rows))
        print(f"seed extension {G}x{G}/{N} {kw}: {n} boards  [{time.time() - t0:.0f}s]")
    out["seedext_solved"] = se

    # ---- the stages in isolation (seeding, extension, one BFS)
    stages = []
    for (G, N, seed, n) in ((10, 5, 41, 6), (7, 4, 42, 8)):
        board = SeedExtensionBoard(G, G, N)
        for k in keys_of(seed, n):
            k = jax.numpy.array(k)
            seeded = board.return_seeded_board(k)
            ext = extend_wires_jax(seeded, k)
            opt0 = optimise_wire(k, ext, 0)
            stages.append(dict(G=G, N=N, key=L(k), seeded=L(seeded), extended=L(ext), optimised_wire0=L(opt0)))
    out["seedext_stages"] = stages
    print(f"stages: {len(stages)}  [{time.time() - t0:.0f}s]")

    with open(OUT, "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print("wrote", OUT, os.path.getsize(OUT), "bytes")
    dataset_section()


if __name__ == "__main__":
    main()
