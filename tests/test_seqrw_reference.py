"""The CPU oracle's SequentialRandomWalk restatement against the reference's own class run on the jax shim
(tests/golden/seqrw_reference.json, made by tests/tools/make_seqrw_fixtures.py).  CPU only."""
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def fx():
    with open(os.path.join(HERE, "golden", "seqrw_reference.json")) as f:
        return json.load(f)


def test_fixture_provenance(fx):
    # the float32 Gumbel pick equals the integer rule (largest mantissa, lowest index) on every draw of every fixture,
    # and NumPy's float32 log is strictly monotone over all 2^23 uniforms
    assert fx["numpy_log_strictly_monotone_on_all_uniforms"] is True
    assert fx["gumbel_draws_checked"] > 5000
    attempts = [c["attempt"] for c in fx["generate"]]
    assert 0 in attempts and 1 in attempts and max(attempts) > 3  # failed, first-attempt and retried boards


def test_oracle_generate_matches_reference(orc, fx):
    for cfg in fx["generate"]:
        G, N = cfg["G"], cfg["N"]
        boards, stats = orc.seqrw_generate_batch(np.array([cfg["key"]], np.uint32), G, N)
        assert boards[0].tolist() == cfg["board"], (G, N, cfg["key"])
        assert int(stats[0, 0]) == cfg["attempt"]
        if cfg["attempt"] == 0:
            assert not np.any(boards[0])


def test_oracle_starts_ends_match_reference(orc, fx):
    for cfg in fx["starts_ends"]:
        s, e = orc.seqrw_starts_ends(np.array(cfg["key"], np.uint32), cfg["G"], cfg["N"])
        assert s.tolist() == cfg["starts"] and e.tolist() == cfg["ends"], cfg["key"]


def test_oracle_generator_states_match_reference(orc, fx):
    saw_failed = False
    for cfg in fx["generator_states"]:
        st = orc.state_batch("sequential_random_walk", np.array([cfg["key_in"]], np.uint32), cfg["G"], cfg["N"])
        assert st["key"][0].tolist() == cfg["key"] and st["grid"][0].tolist() == cfg["grid"] and int(st["step_count"][0]) == cfg["step_count"]
        assert st["agent_id"][0].tolist() == cfg["agent_id"] and st["start"][0].tolist() == cfg["start"]
        assert st["target"][0].tolist() == cfg["target"] and st["position"][0].tolist() == cfg["position"]
        saw_failed |= all(p == [0, 0] for p in cfg["start"]) and all(p == [0, 0] for p in cfg["target"]) and cfg["N"] > 1
    assert saw_failed  # a failed generation: every pin at (0, 0), the last scatter stays


def test_oracle_boards_are_wires(orc):
    """Structure of what the walk produces: every wire of a successful board is one head, one target and a simple chain."""
    keys = orc.split(orc.PRNGKey(3), 512)
    for G, N in ((6, 3), (10, 5), (10, 12)):
        boards, stats = orc.seqrw_generate_batch(keys, G, N)
        ok = stats[:, 0] > 0
        assert ok.any()
        flags = orc.validate_batch(boards[ok], N)
        assert ((flags & (1 | 8 | 32 | 64 | 128)) == 0).all()
        for w in range(N):
            assert ((boards[ok] == 3 * w + 2).sum(axis=(1, 2)) == 1).all() and ((boards[ok] == 3 * w + 3).sum(axis=(1, 2)) == 1).all()
