"""Validity rules against the REFERENCE'S OWN verdicts.

tests/golden/validity_reference.npz holds what the reference's NumPy functions (imported unmodified:
post_processor_utils_numpy.verify_wire_validity, BoardProcessor.is_valid_board / verify_* /
get_path_from_head_and_target) say about 2 400 generated and corrupted boards
(tests/tools/make_validity_fixtures.py).  The C oracle is checked here, rbg_validate in
test_gpu_parity.py::test_validate_matches_reference_verdicts.
"""
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_validity_fixture():
    z = np.load(os.path.join(ROOT, "tests", "golden", "validity_reference.npz"))
    return {k: z[k] for k in z.files}


def check_flags_against_reference(fx, flags_of):
    """flags_of(boards int32[B,G,G], N) -> int32[B]; compares every reference verdict."""
    G_all, N_all = fx["G"].astype(int), fx["N"].astype(int)
    n_checked = 0
    for G, N in sorted(set(zip(G_all.tolist(), N_all.tolist()))):
        idx = np.nonzero((G_all == G) & (N_all == N))[0]
        boards = np.ascontiguousarray(fx["boards"][idx, :G, :G].astype(np.int32))
        flags = np.asarray(flags_of(boards, N)).astype(int)
        outcome, enc_ok, missing = fx["outcome"][idx].astype(int), fx["enc_ok"][idx].astype(bool), fx["missing"][idx].astype(bool)
        wire_valid, path_found = fx["wire_valid"][idx].astype(bool), fx["path_found"][idx, :N].astype(int)
        # EncodingOutOfRangeError <=> bit 0, and then nothing else is evaluated (is_valid_board raises there)
        np.testing.assert_array_equal((flags & 1) != 0, ~enc_ok)
        assert np.all(flags[~enc_ok] == 1)
        ok = enc_ok
        # MissingHeadTailError <=> bit 1; verify_wire_validity <=> not bit 2
        np.testing.assert_array_equal(((flags & 2) != 0)[ok], missing[ok])
        np.testing.assert_array_equal(((flags & 4) == 0)[ok], wire_valid[ok])
        # PathNotFoundError for some wire <=> bit 6
        np.testing.assert_array_equal(((flags & 64) != 0)[ok], (path_found == 0).any(axis=1)[ok])
        # first failing rule in the reference's order = lowest of bits 0, 1, 2
        expect = np.where(flags & 1, 1, np.where(flags & 2, 2, np.where(flags & 4, 3, 0)))
        np.testing.assert_array_equal(expect, outcome)
        # the derived bits are consistent with the pinned ones
        assert np.all(((flags & 64) == 0) | ((flags & 8) != 0))     # no path through own + EMPTY cells => none through own cells
        assert np.all(((flags & 128) == 0) | ((flags & 6) != 0))    # "not excused" only qualifies rules 2 / 4
        assert np.all(((flags & 16) == 0) | ((flags & 6) == 6))     # the reference rejects a lone TARGET on both counts
        n_checked += len(idx)
    return n_checked


def test_fixture_is_substantial():
    fx = load_validity_fixture()
    assert len(fx["outcome"]) >= 2000
    counts = np.bincount(fx["outcome"].astype(int), minlength=4)
    assert counts.min() >= 200, counts  # valid boards and each of the three exceptions
    assert (fx["path_found"] == 0).sum() >= 100 and (fx["path_found"] == 1).sum() >= 1000


def test_oracle_validate_matches_reference_verdicts(orc):
    fx = load_validity_fixture()
    assert check_flags_against_reference(fx, lambda boards, N: orc.validate_batch(boards, N)) == len(fx["outcome"])


def test_oracle_generated_boards_are_valid_up_to_zero_length_wires(orc):
    """What the generators promise (north_star invariants): every wire connects its own start and target, no
    crossings; the only excused defect is ParallelRandomWalk's zero-length wire (a lone TARGET)."""
    for G, N in ((10, 5), (7, 12), (14, 7)):
        keys = orc.split(orc.PRNGKey(5), 300)
        _, _, solved, _ = orc.prw_generate_batch(keys, G, N)
        fl = orc.validate_batch(solved, N)
        assert np.all((fl & (1 | 8 | 32 | 64 | 128)) == 0)
        assert np.all(((fl & 16) != 0) == ((fl & 6) != 0))
    keys = orc.split(orc.PRNGKey(6), 200)
    solved, _ = orc.seedext_solved_batch(keys, 10, 5)
    assert np.all(orc.validate_batch(solved, 5) == 0)
