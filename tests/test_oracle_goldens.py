"""The CPU oracle against every golden vector the reference owns for this path.

Sources (paths relative to /root/reference, copied as literals into
tests/golden/reference_goldens.json):
  routing_board_generation/board_generation_methods/jax_implementation/board_generation/
      test_parallel_random_walk_board.py  (the reference's only test file)
  package_evaluation/profiling_generators.ipynb cells 4 and 13 (stored States)
These pin jax.random (threefry2x32, split, random_bits, _shuffle, choice with p,
uniform) and the ParallelRandomWalk / Uniform generators.  No GPU needed.
"""
import numpy as np
import pytest


def test_threefry_known_answers(orc, goldens):
    g = goldens["rng"]
    assert list(orc.threefry2x32(0, 0, 0, 0)) == g["block_k00_c00"]
    k0, k1 = orc.PRNGKey(0), orc.PRNGKey(1)
    assert orc.split(k0).tolist() == g["split_key0"]
    assert orc.split(k1).tolist() == g["split_key1"]
    assert orc.split(k0, 3).tolist() == g["split_key0_3"]
    assert orc.split(k0, 5).tolist() == g["split_key0_5"]
    assert orc.random_bits(k0, 5).tolist() == g["random_bits_key0_5"]
    assert float(orc.uniform(k0)) == g["uniform_key0"]
    assert [orc.randint(k0, 0, hi) for hi in (2, 4, 10)] == g["randint_key0_hi_2_4_10"]
    assert orc.shuffle_iota(k0, 4).tolist() == g["shuffle_key0_4"]


def test_prngkey_layout(orc):
    assert orc.PRNGKey(0).tolist() == [0, 0]
    assert orc.PRNGKey(2**32 + 7).tolist() == [1, 7]


def test_split_slice_matches_full_split(orc):
    k = orc.PRNGKey(42)
    full = orc.split(k, 1000)
    for off, cnt in ((0, 1000), (0, 1), (499, 3), (500, 500), (999, 1), (123, 456)):
        assert np.array_equal(orc.split_slice(k, 1000, off, cnt), full[off:off + cnt])


# ---- test_parallel_random_walk_board.py -----------------------------------
def test_initialise_agents(orc, goldens):  # :228-233
    g = goldens["prw_5x5_3"]
    grid, pos = orc.prw_initialise_agents(orc.PRNGKey(0), 5, 3)
    assert grid.tolist() == g["valid_starting_grid"]
    assert pos.tolist() == g["starts"]


def test_step(orc, goldens):  # :202-226
    g = goldens["prw_5x5_3"]
    key = orc.PRNGKey(0)
    nk, grid, pos, actions, coll = orc.prw_step(key, g["valid_starting_grid"], g["starts"])
    assert actions.tolist() == g["step_actions"]
    assert grid.tolist() == g["valid_starting_grid_after_1_step"]
    assert pos.tolist() == g["positions_after_1_step"]
    assert nk.tolist() == orc.split(key)[1].tolist()
    assert coll == 0


def test_generate_board(orc, goldens):  # :161-186
    g = goldens["prw_5x5_3"]
    heads, targets, solved, stats = orc.prw_generate(orc.PRNGKey(0), 5, 3)
    assert heads.tolist() == g["heads"]
    assert targets.tolist() == g["targets"]
    assert solved.tolist() == g["valid_end_grid2"]
    assert stats[0] == 6  # 6 trips


def test_generate_board_for_various_keys(orc):  # :188-200
    boards = [orc.prw_generate(orc.PRNGKey(s), 5, 3)[2].tobytes() for s in range(10)]
    assert len(set(boards)) == 10


def test_continue_stepping(orc, goldens):  # :245-268
    g = goldens["prw_5x5_3"]
    assert orc.prw_continue_stepping(g["valid_starting_grid"], g["starts"])
    assert not orc.prw_continue_stepping(g["valid_end_grid"], g["agents_finished_position"])


def test_adjacent_cells(orc, goldens):  # :380-397
    for cell, exp in goldens["prw_5x5_3"]["adjacent_cells"].items():
        assert orc.prw_adjacent_cells(5, int(cell)).tolist() == exp


def test_available_cells(orc, goldens):  # :399-415
    g = goldens["prw_5x5_3"]
    assert orc.prw_available_cells(g["valid_end_grid"], 8).tolist() == g["available_cells_end_grid_8"]
    assert orc.prw_available_cells(g["grid_to_test_available_cells"], 8).tolist() == g["available_cells_test_grid_8"]


def test_is_cell_free(orc, goldens):  # :417-433 (32 is out of range: the gather clamps)
    g = goldens["prw_5x5_3"]
    for cell, exp in g["is_cell_free"].items():
        assert orc.prw_is_cell_free(g["valid_starting_grid"], int(cell)) is exp


def test_action_from_positions(orc, goldens):  # :326-378
    for p1, p2, exp in goldens["prw_5x5_3"]["action_from_positions"]:
        assert orc.prw_action_from_positions(5, p1, p2) == exp


# ---- stored notebook States ------------------------------------------------
def test_prw_generator_state_notebook(orc, goldens):  # profiling_generators.ipynb cell 13
    g = goldens["prw_generator_10x10_5_key0"]
    st = orc.state_batch("parallel_random_walk", orc.PRNGKey(0), 10, 5)
    assert st["start"][0].tolist() == g["start"]
    assert st["target"][0].tolist() == g["target"]
    assert st["position"][0].tolist() == g["start"]
    assert st["key"][0].tolist() == g["key"]
    assert st["step_count"][0] == 0 and st["agent_id"][0].tolist() == [0, 1, 2, 3, 4]
    grid = np.zeros((10, 10), np.int32)
    for i, (s, t) in enumerate(zip(g["start"], g["target"])):
        grid[s[0], s[1]] = 2 + 3 * i
    for i, (s, t) in enumerate(zip(g["start"], g["target"])):
        grid[t[0], t[1]] = 3 + 3 * i
    assert np.array_equal(st["grid"][0], grid)
    # the solved board behind it (derived, SURVEY B6')
    _, _, solved, stats = orc.prw_generate(orc.split(orc.PRNGKey(0))[0], 10, 5)
    assert solved.tolist() == g["solved_grid_derived"]
    assert stats[0] == 13


def test_uniform_generator_state_notebook(orc, goldens):  # profiling_generators.ipynb cell 4
    g = goldens["uniform_generator_10x10_5_key0"]
    st = orc.state_batch("uniform", orc.PRNGKey(0), 10, 5)
    assert st["start"][0].tolist() == g["start"]
    assert st["target"][0].tolist() == g["target"]
    assert st["key"][0].tolist() == g["key"]


def test_seedext_seed_cells(orc, goldens):  # SURVEY Appendix C (derived; shares cells with notebook cell 10)
    g = goldens["seedext_seed_cells_10x10_5_key0"]
    key = orc.PRNGKey(0)
    gen_key = orc.split(key)[0]      # RSG:34
    seedkey = orc.split(gen_key)[1]  # SE:171
    board = orc.seedext_seeded_board(seedkey, 10, 5)
    for i, (tgt, pos) in enumerate(g["wires"]):
        assert board[tgt[0], tgt[1]] == 3 * i + 3
        assert board[pos[0], pos[1]] == 3 * i + 2
    assert (board > 0).sum() == 10


@pytest.mark.parametrize("G,N", [(5, 3), (10, 5), (14, 7), (20, 10), (32, 16)])
def test_prw_boards_are_valid(orc, G, N):
    """Every generated board is sound by the reference's validity rules
    (post_processor_utils_numpy.py:34-155, board_processor.py:111-162); the only defect is the
    zero-length-wire quirk the reference can emit (a lone TARGET, flagged 16, which the
    reference's own checker rejects as 2 | 4)."""
    keys = orc.split(orc.PRNGKey(1), 256)
    _, _, solved, stats = orc.prw_generate_batch(keys, G, N)
    flags = orc.validate_batch(solved, N)
    assert ((flags & (1 | 8 | 32 | 64 | 128)) == 0).all()
    assert (((flags & 16) != 0) == ((flags & 6) != 0)).all()
    assert stats[:, 0].min() >= 0


@pytest.mark.parametrize("G,N", [(6, 3), (10, 5), (14, 7)])
def test_seedext_boards_are_valid(orc, G, N):
    keys = orc.split(orc.PRNGKey(2), 64)
    boards, stats = orc.seedext_solved_batch(keys, G, N)
    assert (orc.validate_batch(boards, N) == 0).all()
    assert (stats[:, 2] == 0).all()
