#!/usr/bin/env python
"""Connector TimeStep fixtures from an INDEPENDENT second restatement of jumanji==0.2.2.

    python tests/tools/make_connector_fixtures.py        # writes tests/golden/connector_reference.npz

jumanji is not installed here or on the GPU box and its source is not under /root/reference, so the
Connector env cannot be run "as is".  What CAN be run: tests/tools/jax_shim/jumanji/{env,types,wrappers}.py and
.../connector/{env,reward,utils}.py restate the upstream classes from their published source, in upstream's
own jax idiom (vmap / lax.cond / .at[]), on the NumPy stand-in for jax -- written without looking at
oracle/rbg_oracle.c or the CUDA kernels.  This script drives them the way the reference does:

  * the env composition of rl_training/setup_train.py:107-166: `Connector(generator=...)`, then
    `MultiToSingleWrapper`, then `VmapAutoResetWrapper`, with the reference's OWN generator classes
    (imported unmodified from /root/reference): ParallelRandomWalkGenerator, UniformRandomGenerator,
    SeedExtensionGenerator, BoardDatasetGeneratorJAX;
  * the reset recipe of demos/board_generator_demo.py:29-97 -- `state_from_board` and `board_to_env` are
    cut out of the reference file with `ast` and executed unmodified;
  * the episode loop of package_evaluation/load_and_test_agents.ipynb cell 10 (reset, step until
    `timestep.last()`, read `extras["num_connections"]`) on a single un-batched env.

Agreement of the C oracle and the CUDA path with these fixtures is N-VERSION AGREEMENT between independent
restatements of upstream, not an upstream pin (DESIGN.md §2 says so).
Actions are inputs: half of them uniform over {0..4} (walls, occupied cells, head-on collisions), half
uniform over the legal moves of the previous mask, from a NumPy generator.
"""
from __future__ import annotations

import ast
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SHIM = os.path.join(HERE, "jax_shim")
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden", "connector_reference.npz")

sys.path.insert(0, SHIM)
sys.path.insert(0, REF)

import jax  # noqa: E402
import jax.numpy as jnp  # noqa: E402
from jumanji.environments.routing.connector.env import Connector  # noqa: E402
from jumanji.environments.routing.connector.reward import DenseRewardFn  # noqa: E402
from jumanji.wrappers import MultiToSingleWrapper, VmapAutoResetWrapper  # noqa: E402


def A(x, dt):
    return np.asarray(x).astype(dt)


def record_state(st, rec, pre):
    rec[pre + "grid"].append(A(st.grid, np.int8))
    rec[pre + "step_count"].append(A(st.step_count, np.int32))
    rec[pre + "start"].append(A(st.agents.start, np.int8))
    rec[pre + "target"].append(A(st.agents.target, np.int8))
    rec[pre + "position"].append(A(st.agents.position, np.int8))
    rec[pre + "key"].append(A(st.key, np.uint32))


def record_timestep(ts, rec, pre):
    rec[pre + "obs"].append(A(ts.observation.grid, np.int8))
    rec[pre + "action_mask"].append(A(ts.observation.action_mask, np.uint8))
    rec[pre + "obs_step_count"].append(A(ts.observation.step_count, np.int32))
    rec[pre + "reward"].append(A(ts.reward, np.float32))
    rec[pre + "discount"].append(A(ts.discount, np.float32))
    rec[pre + "step_type"].append(A(ts.step_type, np.int8))
    rec[pre + "num_connections"].append(A(ts.extras["num_connections"], np.int32))
    rec[pre + "ratio_connections"].append(A(ts.extras["ratio_connections"], np.float32))
    rec[pre + "total_path_length"].append(A(ts.extras["total_path_length"], np.int32))


STATE_F = ("grid", "step_count", "start", "target", "position", "key")
TS_F = ("obs", "action_mask", "obs_step_count", "reward", "discount", "step_type", "num_connections", "ratio_connections", "total_path_length")


def pick_actions(rng, mask):
    """mask uint8[..., N, 5] -> int32[..., N]: half uniform over 0..4, half uniform over the legal moves."""
    shp = mask.shape[:-1]
    any_a = rng.integers(0, 5, size=shp)
    u = rng.random(size=shp + (5,)) * (mask > 0)
    legal = u.argmax(axis=-1)
    return np.where(rng.random(size=shp) < 0.5, any_a, legal).astype(np.int32)


def make_generator(name, G, N):
    from routing_board_generation.rl_training.offline_generation.dataset_generator_jax import BoardDatasetGeneratorJAX
    from routing_board_generation.rl_training.online_generators.parallel_random_walk_generator import ParallelRandomWalkGenerator
    from routing_board_generation.rl_training.online_generators.random_seed_generator import SeedExtensionGenerator
    from routing_board_generation.rl_training.online_generators.uniform_generator import UniformRandomGenerator

    if name == "parallel_random_walk":
        return ParallelRandomWalkGenerator(G, N)
    if name == "uniform":
        return UniformRandomGenerator(G, N)
    if name == "seed_extension":
        return SeedExtensionGenerator(G, N)
    if name == "sequential_random_walk":  # setup_train.py:137-141 online_seq_rw
        from routing_board_generation.board_generation_methods.jax_implementation.board_generation.sequential_random_walk import SequentialRandomWalkBoard
        from routing_board_generation.rl_training.online_generators.sequential_random_walk_generator import SequentialRandomWalkGenerator

        # not instantiable as shipped (abstract methods of the NumPy AbstractBoard): lift the check, see make_seqrw_fixtures.py
        SequentialRandomWalkBoard.__abstractmethods__ = frozenset()
        return SequentialRandomWalkGenerator(G, N)
    if name.startswith("offline_"):  # setup_train.py:119-131
        return BoardDatasetGeneratorJAX(grid_size=G, num_agents=N, board_name=name, number_of_boards=7)
    raise ValueError(name)


def vmapped_scenario(out, tag, gen_name, G, N, B, T, time_limit, seed, aggregate):
    """setup_train._make_raw_env + setup_env (setup_train.py:107-166): Vmap(MultiToSingle?(Connector(generator)))."""
    t0 = time.time()
    gen = make_generator(gen_name, G, N)
    env = Connector(generator=gen, reward_fn=DenseRewardFn(timestep_reward=-0.03, connected_reward=0.1), time_limit=time_limit)
    if aggregate:
        env = MultiToSingleWrapper(env)
    env = VmapAutoResetWrapper(env)
    keys = jax.random.split(jax.random.PRNGKey(seed), B)
    rng = np.random.default_rng(seed)
    rec = {"s_" + f: [] for f in STATE_F}
    rec.update({"t_" + f: [] for f in TS_F})
    rec["action"] = []
    state, ts = env.reset(keys)
    record_state(state, rec, "s_")
    record_timestep(ts, rec, "t_")
    for t in range(T):
        act = pick_actions(rng, A(ts.observation.action_mask, np.uint8))
        state, ts = env.step(state, jnp.array(act))
        rec["action"].append(act)
        record_state(state, rec, "s_")
        record_timestep(ts, rec, "t_")
    for k, v in rec.items():
        out[f"{tag}/{k}"] = np.stack(v)
    out[f"{tag}/keys"] = A(keys, np.uint32)
    if gen_name.startswith("offline_"):
        out[f"{tag}/dataset_heads"] = A(gen.heads, np.int32)
        out[f"{tag}/dataset_targets"] = A(gen.targets, np.int32)
    st = np.stack(rec["t_step_type"])
    print(f"{tag}: {gen_name} {G}x{G}/{N} B={B} T={T} time_limit={time_limit} aggregate={aggregate}: {int((st == 2).sum())} terminal env-steps  [{time.time() - t0:.0f}s]", flush=True)
    return dict(tag=tag, kind="vmapped", generator=gen_name, G=G, N=N, B=B, T=T, time_limit=time_limit, seed=seed, aggregate=bool(aggregate))


def episode_scenario(out, tag, gen_name, G, N, episodes, time_limit, seed):
    """load_and_test_agents.ipynb cell 10: single env, reset, step until timestep.last(), extras['num_connections']."""
    t0 = time.time()
    env = Connector(generator=make_generator(gen_name, G, N), time_limit=time_limit)
    rng = np.random.default_rng(seed)
    key = jax.random.PRNGKey(seed)
    rec = {"s_" + f: [] for f in STATE_F}
    rec.update({"t_" + f: [] for f in TS_F})
    rec["action"], rec["is_reset"], rec["reset_key"] = [], [], []
    for _ in range(episodes):
        key, reset_key = jax.random.split(key)
        state, ts = env.reset(reset_key)
        record_state(state, rec, "s_")
        record_timestep(ts, rec, "t_")
        rec["action"].append(np.zeros(N, np.int32))
        rec["is_reset"].append(np.int8(1))
        rec["reset_key"].append(A(reset_key, np.uint32))
        while not ts.last():
            act = pick_actions(rng, A(ts.observation.action_mask, np.uint8))
            state, ts = env.step(state, jnp.array(act))
            record_state(state, rec, "s_")
            record_timestep(ts, rec, "t_")
            rec["action"].append(act)
            rec["is_reset"].append(np.int8(0))
            rec["reset_key"].append(A(reset_key, np.uint32))
    for k, v in rec.items():
        out[f"{tag}/{k}"] = np.stack(v)
    print(f"{tag}: {episodes} episodes, {len(rec['action'])} records  [{time.time() - t0:.0f}s]", flush=True)
    return dict(tag=tag, kind="episodes", generator=gen_name, G=G, N=N, time_limit=time_limit, seed=seed, episodes=episodes)


def demo_recipe_scenario(out, tag, seed, n):
    """demos/board_generator_demo.py:29-97: state_from_board + board_to_env, cut out of the reference file and
    executed as they are (they hard-code 10x10 / 5 wires) on solved ParallelRandomWalk boards."""
    from jumanji.env import Environment
    from jumanji.environments.routing.connector.constants import POSITION, TARGET
    from jumanji.environments.routing.connector.types import Agent, Observation, State
    from jumanji.environments.routing.connector.utils import get_position, get_target
    from jumanji.types import restart
    from routing_board_generation.board_generation_methods.jax_implementation.board_generation.parallel_random_walk import ParallelRandomWalkBoard

    src = open(os.path.join(REF, "demos", "board_generator_demo.py")).read()
    tree = ast.parse(src)
    fns = [n_ for n_ in tree.body if isinstance(n_, ast.FunctionDef) and n_.name in ("state_from_board", "board_to_env")]
    assert len(fns) == 2
    ns = dict(jax=jax, jnp=jnp, Connector=Connector, Agent=Agent, Observation=Observation, State=State, Environment=Environment,
              get_position=get_position, get_target=get_target, POSITION=POSITION, TARGET=TARGET, restart=restart)
    exec(compile(ast.Module(body=fns, type_ignores=[]), "board_generator_demo.py", "exec"), ns)
    board_gen = ParallelRandomWalkBoard(10, 10, 5)
    rec = {"s_" + f: [] for f in STATE_F}
    rec.update({"t_" + f: [] for f in TS_F})
    rec["board"] = []
    for k in jax.random.split(jax.random.PRNGKey(seed), n):
        _, _, solved = board_gen.generate_board(k)
        # the recipe reads a board that carries heads and targets; the solved board does
        ns["key"] = k  # state_from_board closes over a module-level `key` in the demo
        state, ts = ns["board_to_env"](solved)
        rec["board"].append(A(solved, np.int8))
        record_state(state, rec, "s_")
        record_timestep(ts, rec, "t_")
    for k, v in rec.items():
        out[f"{tag}/{k}"] = np.stack(v)
    print(f"{tag}: {n} boards through the demo's reset recipe", flush=True)
    return dict(tag=tag, kind="demo_recipe", G=10, N=5, seed=seed, n=n)


def main():
    import json

    out, meta = {}, []
    meta.append(vmapped_scenario(out, "prw10", "parallel_random_walk", 10, 5, 12, 64, 50, 201, False))
    meta.append(vmapped_scenario(out, "prw10agg", "parallel_random_walk", 10, 5, 6, 56, 50, 202, True))
    meta.append(vmapped_scenario(out, "uni6", "uniform", 6, 4, 16, 30, 8, 203, False))
    meta.append(vmapped_scenario(out, "prw5", "parallel_random_walk", 5, 3, 16, 24, 5, 204, True))
    meta.append(vmapped_scenario(out, "prw7dense", "parallel_random_walk", 7, 12, 6, 10, 4, 205, False))
    meta.append(vmapped_scenario(out, "se6", "seed_extension", 6, 3, 6, 14, 6, 206, False))
    meta.append(vmapped_scenario(out, "ds6prw", "offline_parallel_rw", 6, 4, 8, 20, 6, 207, True))
    meta.append(vmapped_scenario(out, "ds6se", "offline_seed_extension", 6, 3, 6, 14, 5, 208, False))
    # small grids, long time limit: episodes end because every agent is connected or blocked
    meta.append(vmapped_scenario(out, "uni4", "uniform", 4, 2, 16, 40, 50, 212, False))
    meta.append(vmapped_scenario(out, "prw4agg", "parallel_random_walk", 4, 3, 12, 30, 50, 213, True))
    meta.append(episode_scenario(out, "episodes_uniform", "uniform", 10, 5, 4, 50, 209))
    meta.append(episode_scenario(out, "episodes_prw", "parallel_random_walk", 8, 4, 4, 20, 210))
    meta.append(demo_recipe_scenario(out, "demo_recipe", 211, 6))
    meta.append(vmapped_scenario(out, "seq6", "sequential_random_walk", 6, 3, 6, 14, 6, 214, False))
    meta.append(vmapped_scenario(out, "seq4agg", "sequential_random_walk", 4, 5, 6, 8, 3, 215, True))
    out["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


def append_seqrw():
    """`python tests/tools/make_connector_fixtures.py append_seqrw`: adds the SequentialRandomWalkGenerator scenarios to the
    existing file (the other scenarios are kept as they are)."""
    import json

    z = np.load(OUT)
    out = {k: z[k] for k in z.files}
    meta = [m for m in json.loads(bytes(z["meta"]).decode()) if m.get("generator") != "sequential_random_walk"]
    for k in [k for k in out if k.split("/")[0] in ("seq6", "seq4agg")]:
        del out[k]
    meta.append(vmapped_scenario(out, "seq6", "sequential_random_walk", 6, 3, 6, 14, 6, 214, False))
    meta.append(vmapped_scenario(out, "seq4agg", "sequential_random_walk", 4, 5, 6, 8, 3, 215, True))  # crowded: retries and failed generations
    out["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    append_seqrw() if sys.argv[1:] == ["append_seqrw"] else main()
