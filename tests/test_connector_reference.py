"""Connector reset / step / auto-reset / wrappers against tests/golden/connector_reference.npz.

The fixtures come from an independent second restatement of jumanji==0.2.2 (tests/tools/jax_shim/jumanji,
written from the published source in upstream's jax idiom, NOT from oracle/rbg_oracle.c) driven the way the
reference drives the env (tests/tools/make_connector_fixtures.py): the env composition of
rl_training/setup_train.py:107-166 with the reference's own generator classes, the reset recipe of
demos/board_generator_demo.py:29-97 executed unmodified, and the episode loop of
package_evaluation/load_and_test_agents.ipynb cell 10.  jumanji itself is installed nowhere, so this is
N-version agreement between independent restatements, not an upstream pin.

The C oracle is checked here; the CUDA path in test_gpu_parity.py::test_connector_reference_fixtures_on_gpu
through the same `run_*` drivers.
"""
import json
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

STATE_F = ("grid", "step_count", "start", "target", "position", "key")
TS_F = ("obs", "action_mask", "obs_step_count", "reward", "discount", "step_type", "num_connections", "ratio_connections", "total_path_length")
GEN_KIND = {"parallel_random_walk": "parallel_random_walk", "uniform": "uniform", "seed_extension": "seed_extension",
            "offline_parallel_rw": "dataset", "offline_seed_extension": "dataset", "sequential_random_walk": "sequential_random_walk"}


def load_connector_fixture():
    z = np.load(os.path.join(ROOT, "tests", "golden", "connector_reference.npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    return z, meta


def _f32_equal(a, b):
    return np.array_equal(np.asarray(a, np.float32).view(np.uint32), np.asarray(b, np.float32).view(np.uint32))


def compare_record(z, tag, i, state, ts, aggregate, where):
    """state / ts: dicts of numpy arrays in the oracle's layout (un-aggregated per-agent reward / discount)."""
    for f in STATE_F:
        exp = z[f"{tag}/s_{f}"][i]
        assert np.array_equal(np.asarray(state[f]).astype(np.int64), exp.astype(np.int64)), f"{where}: state.{f}"
    for f in ("obs", "action_mask", "obs_step_count", "step_type", "num_connections", "total_path_length"):
        exp = z[f"{tag}/t_{f}"][i]
        assert np.array_equal(np.asarray(ts[f]).astype(np.int64), exp.astype(np.int64)), f"{where}: timestep.{f}"
    assert _f32_equal(ts["ratio_connections"], z[f"{tag}/t_ratio_connections"][i]), f"{where}: ratio_connections"
    rew, disc = np.asarray(ts["reward"], np.float32), np.asarray(ts["discount"], np.float32)
    if aggregate and rew.ndim == z[f"{tag}/t_reward"][i].ndim + 1:
        # MultiToSingleWrapper: jnp.sum / jnp.max over the agent axis.  The sum of N float32 rewards is compared
        # to 1 ulp of its magnitude: XLA's / NumPy's reduction order is not part of the contract.
        np.testing.assert_allclose(rew.sum(axis=-1, dtype=np.float32), z[f"{tag}/t_reward"][i], rtol=0, atol=2e-7, err_msg=f"{where}: summed reward")
        assert _f32_equal(disc.max(axis=-1), z[f"{tag}/t_discount"][i]), f"{where}: max discount"
    elif aggregate:
        np.testing.assert_allclose(rew, z[f"{tag}/t_reward"][i], rtol=0, atol=2e-7, err_msg=f"{where}: summed reward")
        assert _f32_equal(disc, z[f"{tag}/t_discount"][i]), f"{where}: max discount"
    else:
        assert _f32_equal(rew, z[f"{tag}/t_reward"][i]), f"{where}: reward"
        assert _f32_equal(disc, z[f"{tag}/t_discount"][i]), f"{where}: discount"


# ------------------------------------------------------------------ oracle drivers
class OracleEnv:
    """reset / step of the C oracle in the fixture's terms."""

    def __init__(self, orc, m, z):
        self.orc, self.m = orc, m
        self.kind = GEN_KIND[m["generator"]]
        self.dataset = (z[f"{m['tag']}/dataset_heads"], z[f"{m['tag']}/dataset_targets"]) if self.kind == "dataset" else None

    def reset(self, keys):
        G, N = self.m["G"], self.m["N"]
        if self.kind == "dataset":
            st = self.orc.dataset_state_batch(keys, G, N, *self.dataset)
            return st, self.orc.connector_observe_batch(st)
        return self.orc.connector_reset_batch(self.kind, keys, G, N)

    def step(self, st, action, autoreset):
        return self.orc.connector_step_batch(st, action, time_limit=self.m["time_limit"], autoreset_kind=self.kind if autoreset else -1, dataset=self.dataset)


def run_vmapped(z, m, env):
    tag = m["tag"]
    st, ts = env.reset(z[f"{tag}/keys"])
    compare_record(z, tag, 0, st, ts, m["aggregate"], f"{tag} reset")
    n_term = n_early = n_conn = 0
    for t in range(m["T"]):
        prev_sc = np.asarray(st["step_count"]).copy()
        st, ts = env.step(st, z[f"{tag}/action"][t], True)
        compare_record(z, tag, t + 1, st, ts, m["aggregate"], f"{tag} step {t}")
        last = np.asarray(ts["step_type"]) == 2
        n_term += int(last.sum())
        n_early += int((last & (prev_sc + 1 < m["time_limit"])).sum())
        n_conn += int(np.asarray(ts["num_connections"]).sum())
    return n_term, n_early, n_conn


def run_episodes(z, m, env):
    tag = m["tag"]
    n = len(z[f"{tag}/is_reset"])
    st = None
    for i in range(n):
        if z[f"{tag}/is_reset"][i]:
            st, ts = env.reset(z[f"{tag}/reset_key"][i][None])
        else:
            st, ts = env.step(st, z[f"{tag}/action"][i][None], False)
        one = lambda d: {k: np.asarray(v)[0] for k, v in d.items()}  # noqa: E731
        compare_record(z, tag, i, one(st), one(ts), False, f"{tag} record {i}")
    return n


def test_fixture_covers_the_transition_surface():
    z, meta = load_connector_fixture()
    kinds = {m["generator"] for m in meta if "generator" in m}
    assert {"parallel_random_walk", "uniform", "seed_extension", "offline_parallel_rw", "offline_seed_extension", "sequential_random_walk"} <= kinds
    st = np.concatenate([z[f"{m['tag']}/t_step_type"].reshape(-1) for m in meta])
    assert (st == 0).sum() > 50 and (st == 1).sum() > 1000 and (st == 2).sum() > 150
    rew = np.concatenate([z[f"{m['tag']}/t_reward"].reshape(-1) for m in meta if not m.get("aggregate")])
    assert (np.abs(rew - 0.07) < 1e-6).sum() > 20, "connections (reward 0.1 - 0.03) must occur"
    assert (rew == 0).sum() > 100, "connected agents earn nothing"


def test_oracle_vmapped_autoreset_matches_reference_fixtures(orc):
    z, meta = load_connector_fixture()
    term = early = conn = 0
    for m in meta:
        if m["kind"] != "vmapped":
            continue
        a, b, c = run_vmapped(z, m, OracleEnv(orc, m, z))
        term, early, conn = term + a, early + b, conn + c
    assert term > 150 and early > 20 and conn > 100, (term, early, conn)


def test_oracle_episode_loop_matches_reference_fixtures(orc):
    z, meta = load_connector_fixture()
    for m in meta:
        if m["kind"] == "episodes":
            assert run_episodes(z, m, OracleEnv(orc, m, z)) > 50


def test_oracle_demo_reset_recipe_matches_reference_fixtures(orc):
    """demos/board_generator_demo.py:29-97 executed unmodified: State from a solved board + the private trio."""
    z, meta = load_connector_fixture()
    m = next(m for m in meta if m["kind"] == "demo_recipe")
    tag = m["tag"]
    for i in range(m["n"]):
        st = {f: np.ascontiguousarray(z[f"{tag}/s_{f}"][i][None].astype(np.uint32 if f == "key" else np.int32)) for f in STATE_F}
        st["agent_id"] = np.arange(m["N"], dtype=np.int32)[None]
        ts = orc.connector_observe_batch(st)
        one = {k: np.asarray(v)[0] for k, v in ts.items()}
        compare_record(z, tag, i, {k: v[0] for k, v in st.items()}, one, False, f"demo recipe {i}")


def test_total_path_length_counts_one_head_per_agent_even_for_zero_length_wires(orc):
    """Decision recorded in DESIGN.md §2: upstream's `_get_extras` counts PATH cells and then adds
    `num_agents` ("add agents' head"), whatever the grid holds; a zero-length wire (start == target, a lone
    TARGET cell, no POSITION cell) still contributes its 1."""
    z, meta = load_connector_fixture()
    m = next(m for m in meta if m["tag"] == "prw7dense")
    grid, tpl = z["prw7dense/s_grid"].astype(int), z["prw7dense/t_total_path_length"].astype(int)
    pos, tgt = z["prw7dense/s_position"], z["prw7dense/s_target"]
    zero_len = (z["prw7dense/s_start"] == tgt).all(axis=-1)
    assert zero_len.any(), "the dense config must contain zero-length wires"
    paths = ((grid > 0) & ((grid - 1) % 3 == 0)).sum(axis=(-1, -2))
    st = z["prw7dense/t_step_type"]
    keep = st != 2  # on auto-reset steps the extras belong to the terminal grid, the State to the new episode
    assert np.array_equal(tpl[keep], (paths + m["N"])[keep])
