"""Board statistics against the reference's own EvaluateEmptyBoard
(benchmarking/benchmarks/empty_board_evaluation.py:31-155, executed unmodified by
tests/tools/make_stats_fixtures.py -> tests/golden/board_stats_reference.npz).  The oracle here, the CUDA kernel in
test_gpu_parity.py::test_board_statistics_match_reference."""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_stats_fixture():
    z = np.load(os.path.join(ROOT, "tests", "golden", "board_stats_reference.npz"))
    return {k: z[k] for k in z.files}


def check_stats_against_reference(fx, stats_of):
    """stats_of(boards int32[B,G,G], N, count_current_wire) -> (scored[B,G,G], detours[B], diversity[B])."""
    n = 0
    for pre in ("", "p_"):
        Gs, Ns = fx[pre + "G"].astype(int), fx[pre + "N"].astype(int)
        for G, N in sorted(set(zip(Gs.tolist(), Ns.tolist()))):
            idx = np.nonzero((Gs == G) & (Ns == N))[0]
            boards = np.ascontiguousarray(fx[pre + "boards"][idx, :G, :G].astype(np.int32))
            scored, det, div = (np.asarray(x) for x in stats_of(boards, N, False))
            assert np.array_equal(scored, fx[pre + "scored"][idx, :G, :G]), (pre, G, N, "scored board")
            exp_det = fx[pre + "detours"][idx] if pre else fx["detours"][idx, 0]
            assert np.array_equal(det, exp_det), (pre, G, N, "count_detours")
            assert np.array_equal(div, fx[pre + "diversity"][idx]), (pre, G, N, "heatmap_score_diversity")
            if not pre:
                _, det1, _ = (np.asarray(x) for x in stats_of(boards, N, True))
                assert np.array_equal(det1, fx["detours"][idx, 1]), (G, N, "count_detours(count_current_wire=True)")
            n += len(idx)
    return n


def test_fixture_is_substantial():
    fx = load_stats_fixture()
    assert len(fx["G"]) >= 400 and len(fx["p_G"]) >= 100
    assert fx["detours"][:, 0].max() > 20 and (fx["detours"][:, 1] > fx["detours"][:, 0]).any()


def test_oracle_board_statistics_match_reference(orc):
    fx = load_stats_fixture()
    n = check_stats_against_reference(fx, lambda boards, N, ccw: orc.board_statistics_batch(boards, ccw))
    assert n == len(fx["G"]) + len(fx["p_G"])
