"""ctypes front-end of the CPU ORACLE (oracle/rbg_oracle.c).

TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import this module.
The product package never does.

NumPy in / NumPy out; every function mirrors one reference entry point (see
rbg_oracle.h for the file:line citations).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Dict, Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "librbg_oracle.so")

GEN_PRW, GEN_UNIFORM, GEN_SEEDEXT, GEN_DATASET, GEN_SEQRW = 0, 1, 2, 3, 4
GEN_KINDS = {"parallel_random_walk": GEN_PRW, "uniform": GEN_UNIFORM, "seed_extension": GEN_SEEDEXT, "dataset": GEN_DATASET, "sequential_random_walk": GEN_SEQRW}

u32p = C.POINTER(C.c_uint32)
i32p = C.POINTER(C.c_int32)
u8p = C.POINTER(C.c_uint8)
i8p = C.POINTER(C.c_int8)
f32p = C.POINTER(C.c_float)


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (oracle/Makefile).  Building the checker is not using it."""
    src = [os.path.join(_HERE, f) for f in ("rbg_oracle.c", "rbg_oracle.h")]
    stale = (not os.path.exists(_SO)) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in src)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-s", "all"], check=True, capture_output=True)
    return _SO


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.orc_uniform.restype = C.c_float
        _lib.orc_randint.restype = C.c_int32
        _lib.orc_choice_p4.restype = C.c_int32
    return _lib


def _u32(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.uint32))


def _i32(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.int32))


def _p(a: np.ndarray, t):
    return a.ctypes.data_as(t)


# ---------------------------------------------------------------- jax.random
def PRNGKey(seed: int) -> np.ndarray:
    return np.array([(int(seed) >> 32) & 0xFFFFFFFF, int(seed) & 0xFFFFFFFF], dtype=np.uint32)


def threefry2x32(k0, k1, x0, x1) -> Tuple[int, int]:
    a, b = C.c_uint32(), C.c_uint32()
    lib().orc_threefry2x32(C.c_uint32(k0), C.c_uint32(k1), C.c_uint32(x0), C.c_uint32(x1), C.byref(a), C.byref(b))
    return a.value, b.value


def split(key, num: int = 2) -> np.ndarray:
    key = _u32(key)
    out = np.empty((num, 2), np.uint32)
    lib().orc_split(_p(key, u32p), C.c_int(num), _p(out, u32p))
    return out


def split_slice(key, B: int, offset: int, count: int) -> np.ndarray:
    key = _u32(key)
    out = np.empty((count, 2), np.uint32)
    lib().orc_split_batch_slice(_p(key, u32p), C.c_int64(B), C.c_int64(offset), C.c_int64(count), _p(out, u32p))
    return out


def random_bits(key, n: int) -> np.ndarray:
    key = _u32(key)
    out = np.empty((n,), np.uint32)
    lib().orc_random_bits(_p(key, u32p), C.c_int(n), _p(out, u32p))
    return out


def uniform(key) -> np.float32:
    key = _u32(key)
    return np.float32(lib().orc_uniform(_p(key, u32p)))


def shuffle_iota(key, n: int) -> np.ndarray:
    key = _u32(key)
    out = np.empty((n,), np.int32)
    lib().orc_shuffle_iota(_p(key, u32p), C.c_int(n), _p(out, i32p))
    return out


def randint(key, lo: int, hi: int, n: Optional[int] = None):
    key = _u32(key)
    if n is None:
        return int(lib().orc_randint(_p(key, u32p), C.c_int32(lo), C.c_int32(hi)))
    out = np.empty((n,), np.int32)
    lib().orc_randint_vec(_p(key, u32p), C.c_int(n), C.c_int32(lo), C.c_int32(hi), _p(out, i32p))
    return out


def choice_p4(key, a, p) -> int:
    key = _u32(key)
    a = _i32(a)
    p = np.ascontiguousarray(np.asarray(p, dtype=np.uint8))
    return int(lib().orc_choice_p4(_p(key, u32p), _p(a, i32p), _p(p, u8p)))


# ------------------------------------------------------- ParallelRandomWalk
def prw_adjacent_cells(G: int, cell: int) -> np.ndarray:
    out = np.empty(4, np.int32)
    lib().orc_prw_adjacent_cells(C.c_int(G), C.c_int(cell), _p(out, i32p))
    return out


def prw_available_cells(grid, cell: int) -> np.ndarray:
    grid = _i32(grid)
    out = np.empty(4, np.int32)
    lib().orc_prw_available_cells(C.c_int(grid.shape[0]), _p(grid, i32p), C.c_int(cell), _p(out, i32p))
    return out


def prw_is_cell_free(grid, cell: int) -> bool:
    grid = _i32(grid)
    return bool(lib().orc_prw_is_cell_free(C.c_int(grid.shape[0]), _p(grid, i32p), C.c_int(cell)))


def prw_action_from_positions(G: int, p1: int, p2: int) -> int:
    return int(lib().orc_prw_action_from_positions(C.c_int(G), C.c_int(p1), C.c_int(p2)))


def prw_initialise_agents(key, G: int, N: int):
    key = _u32(key)
    grid = np.empty((G, G), np.int32)
    pos = np.empty((N, 2), np.int32)
    lib().orc_prw_initialise_agents(_p(key, u32p), C.c_int(G), C.c_int(N), _p(grid, i32p), _p(pos, i32p))
    return grid, pos


def prw_step(key, grid, pos):
    """One `_step`: returns (next_key, grid, positions, actions, n_collided)."""
    key = _u32(key)
    grid = _i32(grid).copy()
    pos = _i32(pos).copy()
    G, N = grid.shape[0], pos.shape[0]
    actions = np.empty(N, np.int32)
    nk = np.empty(2, np.uint32)
    coll = lib().orc_prw_step(_p(key, u32p), C.c_int(G), C.c_int(N), _p(grid, i32p), _p(pos, i32p), _p(actions, i32p), _p(nk, u32p))
    return nk, grid, pos, actions, int(coll)


def prw_continue_stepping(grid, pos) -> bool:
    grid = _i32(grid)
    pos = _i32(pos)
    return bool(lib().orc_prw_continue_stepping(C.c_int(grid.shape[0]), C.c_int(pos.shape[0]), _p(grid, i32p), _p(pos, i32p)))


def prw_generate(key, G: int, N: int):
    """generate_board(key) -> heads[2,N], targets[2,N], solved[G,G], stats[2]."""
    keys = _u32(key).reshape(1, 2)
    h, t, s, st = prw_generate_batch(keys, G, N, nthreads=1)
    return h[0], t[0], s[0], st[0]


def prw_generate_batch(keys, G: int, N: int, nthreads: int = 0):
    keys = _u32(keys).reshape(-1, 2)
    B = keys.shape[0]
    heads = np.empty((B, 2, N), np.int32)
    targets = np.empty((B, 2, N), np.int32)
    solved = np.empty((B, G, G), np.int32)
    stats = np.empty((B, 2), np.int32)
    rc = lib().orc_prw_generate_batch(_p(keys, u32p), C.c_int64(B), C.c_int(G), C.c_int(N), _p(heads, i32p), _p(targets, i32p), _p(solved, i32p), _p(stats, i32p), C.c_int(nthreads))
    if rc:
        raise ValueError(f"orc_prw_generate_batch rc={rc}")
    return heads, targets, solved, stats


# ------------------------------------------------------------ SeedExtension
def seedext_seeded_board(key, G: int, N: int) -> np.ndarray:
    key = _u32(key)
    board = np.empty((G, G), np.int32)
    rc = lib().orc_seedext_seeded_board(_p(key, u32p), C.c_int(G), C.c_int(N), _p(board, i32p))
    if rc:
        raise ValueError(f"orc_seedext_seeded_board rc={rc}")
    return board


def seedext_solved_batch(keys, G: int, N: int, randomness: float = 0.0, two_sided: bool = True, iterations: int = 1, ext_steps: int = -1, nthreads: int = 0):
    keys = _u32(keys).reshape(-1, 2)
    B = keys.shape[0]
    boards = np.empty((B, G, G), np.int32)
    stats = np.empty((B, 3), np.int32)
    rc = lib().orc_seedext_solved_batch(_p(keys, u32p), C.c_int64(B), C.c_int(G), C.c_int(N), C.c_float(randomness), C.c_int(int(two_sided)), C.c_int(iterations), C.c_int64(ext_steps), _p(boards, i32p), _p(stats, i32p), C.c_int(nthreads))
    if rc:
        raise ValueError(f"orc_seedext_solved_batch rc={rc}")
    return boards, stats


def seedext_solved(key, G: int, N: int, **kw) -> np.ndarray:
    return seedext_solved_batch(_u32(key).reshape(1, 2), G, N, nthreads=1, **kw)[0][0]


def seedext_starts_ends(key, G: int, N: int, randomness: float = 0.0, two_sided: bool = True, iterations: int = 1, ext_steps: int = -1):
    key = _u32(key)
    s = np.empty((2, N), np.int32)
    e = np.empty((2, N), np.int32)
    rc = lib().orc_seedext_starts_ends(_p(key, u32p), C.c_int(G), C.c_int(N), C.c_float(randomness), C.c_int(int(two_sided)), C.c_int(iterations), C.c_int64(ext_steps), _p(s, i32p), _p(e, i32p))
    if rc:
        raise ValueError(f"orc_seedext_starts_ends rc={rc}")
    return s, e


def extend_wires(board, key, randomness: float = 0.0, two_sided: bool = True, ext_steps: int = -1):
    board = _i32(board).copy()
    key = _u32(key)
    sweeps = C.c_int32()
    lib().orc_extend_wires(C.c_int(board.shape[0]), _p(board, i32p), _p(key, u32p), C.c_float(randomness), C.c_int(int(two_sided)), C.c_int64(ext_steps), C.byref(sweeps))
    return board, sweeps.value


def optimise_wire(key, board, wire: int):
    board = _i32(board).copy()
    key = _u32(key)
    pops = C.c_int32()
    rc = lib().orc_optimise_wire(_p(key, u32p), C.c_int(board.shape[0]), _p(board, i32p), C.c_int(wire), C.byref(pops))
    return board, pops.value, rc


# ----------------------------------------------------- SequentialRandomWalk
def seqrw_generate_batch(keys, G: int, N: int, nthreads: int = 0):
    """SequentialRandomWalkBoard.generate over keys[B,2] -> boards[B,G,G] int32, stats[B,2] (attempt, steps)."""
    keys = _u32(keys).reshape(-1, 2)
    B = keys.shape[0]
    boards = np.empty((B, G, G), np.int32)
    stats = np.empty((B, 2), np.int32)
    rc = lib().orc_seqrw_generate_batch(_p(keys, u32p), C.c_int64(B), C.c_int(G), C.c_int(N), _p(boards, i32p), _p(stats, i32p), C.c_int(nthreads))
    if rc:
        raise ValueError(f"orc_seqrw_generate_batch rc={rc}")
    return boards, stats


def seqrw_generate(key, G: int, N: int) -> np.ndarray:
    return seqrw_generate_batch(_u32(key).reshape(1, 2), G, N, nthreads=1)[0][0]


def seqrw_starts_ends(key, G: int, N: int):
    key = _u32(key)
    s = np.empty((2, N), np.int32)
    e = np.empty((2, N), np.int32)
    rc = lib().orc_seqrw_starts_ends(_p(key, u32p), C.c_int(G), C.c_int(N), _p(s, i32p), _p(e, i32p))
    if rc:
        raise ValueError(f"orc_seqrw_starts_ends rc={rc}")
    return s, e


# ------------------------------------------------------- generator -> State
def state_batch(kind, keys, G: int, N: int, nthreads: int = 0) -> Dict[str, np.ndarray]:
    if isinstance(kind, str):
        kind = GEN_KINDS[kind]
    keys = _u32(keys).reshape(-1, 2)
    B = keys.shape[0]
    st = dict(
        grid=np.empty((B, G, G), np.int32),
        step_count=np.empty((B,), np.int32),
        agent_id=np.empty((B, N), np.int32),
        start=np.empty((B, N, 2), np.int32),
        target=np.empty((B, N, 2), np.int32),
        position=np.empty((B, N, 2), np.int32),
        key=np.empty((B, 2), np.uint32),
    )
    rc = lib().orc_state_batch(C.c_int(kind), _p(keys, u32p), C.c_int64(B), C.c_int(G), C.c_int(N), _p(st["grid"], i32p), _p(st["step_count"], i32p), _p(st["agent_id"], i32p), _p(st["start"], i32p), _p(st["target"], i32p), _p(st["position"], i32p), _p(st["key"], u32p), C.c_int(nthreads))
    if rc:
        raise ValueError(f"orc_state_batch rc={rc}")
    return st


def dataset_state_batch(keys, G: int, N: int, heads, targets) -> Dict[str, np.ndarray]:
    """BoardDatasetGeneratorJAX.__call__ over keys[B,2]; heads / targets int32[K,2,N]."""
    keys = _u32(keys).reshape(-1, 2)
    heads, targets = _i32(heads), _i32(targets)
    B, K = keys.shape[0], heads.shape[0]
    st = dict(grid=np.empty((B, G, G), np.int32), step_count=np.empty((B,), np.int32), agent_id=np.empty((B, N), np.int32), start=np.empty((B, N, 2), np.int32),
              target=np.empty((B, N, 2), np.int32), position=np.empty((B, N, 2), np.int32), key=np.empty((B, 2), np.uint32))
    for b in range(B):
        sc = C.c_int32()
        lib().orc_dataset_state(_p(keys[b], u32p), C.c_int(G), C.c_int(N), _p(heads, i32p), _p(targets, i32p), C.c_int64(K), _p(st["grid"][b], i32p), C.byref(sc),
                                _p(st["agent_id"][b], i32p), _p(st["start"][b], i32p), _p(st["target"][b], i32p), _p(st["position"][b], i32p), _p(st["key"][b], u32p))
        st["step_count"][b] = sc.value
    return st


# ---------------------------------------------------------------- Connector
def _alloc_timestep(B: int, G: int, N: int) -> Dict[str, np.ndarray]:
    return dict(
        obs=np.empty((B, N, G, G), np.int32),
        action_mask=np.empty((B, N, 5), np.uint8),
        reward=np.zeros((B, N), np.float32),
        discount=np.ones((B, N), np.float32),
        step_type=np.zeros((B,), np.int8),
        num_connections=np.empty((B,), np.int32),
        ratio_connections=np.empty((B,), np.float32),
        total_path_length=np.empty((B,), np.int32),
        obs_step_count=np.zeros((B,), np.int32),
    )


def connector_observe_batch(state: Dict[str, np.ndarray], nthreads: int = 0) -> Dict[str, np.ndarray]:
    """The observation/mask/extras part of Connector.reset (restart timestep)."""
    B, G, _ = state["grid"].shape
    N = state["target"].shape[1]
    ts = _alloc_timestep(B, G, N)
    lib().orc_connector_observe_batch(C.c_int64(B), C.c_int(G), C.c_int(N), _p(state["grid"], i32p), _p(state["target"], i32p), _p(state["position"], i32p), _p(ts["obs"], i32p), _p(ts["action_mask"], u8p), _p(ts["num_connections"], i32p), _p(ts["ratio_connections"], f32p), _p(ts["total_path_length"], i32p), C.c_int(nthreads))
    ts["obs_step_count"][:] = state["step_count"]
    return ts


def connector_reset_batch(kind, keys, G: int, N: int, nthreads: int = 0):
    st = state_batch(kind, keys, G, N, nthreads)
    return st, connector_observe_batch(st, nthreads)


def connector_step_batch(state: Dict[str, np.ndarray], action, time_limit: int = 50, timestep_reward: float = -0.03, connected_reward: float = 0.1, autoreset_kind=-1, nthreads: int = 0, inplace: bool = False, out: Optional[Dict[str, np.ndarray]] = None, dataset=None):
    """Connector.step (autoreset_kind < 0) or VmapAutoResetWrapper(Connector).step.
    `out`: a timestep dict from a previous call to write into (no fresh allocation / page faults).
    `dataset` = (heads[K,2,N], targets[K,2,N]) with autoreset_kind "dataset" (BoardDatasetGeneratorJAX)."""
    if isinstance(autoreset_kind, str):
        autoreset_kind = GEN_KINDS[autoreset_kind]
    if not inplace:
        state = {k: v.copy() for k, v in state.items()}
    B, G, _ = state["grid"].shape
    N = state["target"].shape[1]
    action = _i32(action).reshape(B, N)
    ts = _alloc_timestep(B, G, N) if out is None else out
    if autoreset_kind == GEN_DATASET:
        dh, dt = _i32(dataset[0]), _i32(dataset[1])
        ds_args = (_p(dh, i32p), _p(dt, i32p), C.c_int64(dh.shape[0]))
    else:
        ds_args = (None, None, C.c_int64(0))
    lib().orc_connector_step_batch_ds(
        C.c_int64(B), C.c_int(G), C.c_int(N), _p(state["grid"], i32p), _p(state["step_count"], i32p), _p(state["start"], i32p), _p(state["target"], i32p), _p(state["position"], i32p), _p(state["key"], u32p), _p(action, i32p),
        C.c_int(time_limit), C.c_float(timestep_reward), C.c_float(connected_reward), C.c_int(autoreset_kind), *ds_args,
        _p(ts["obs"], i32p), _p(ts["action_mask"], u8p), _p(ts["reward"], f32p), _p(ts["discount"], f32p), _p(ts["step_type"], i8p), _p(ts["num_connections"], i32p), _p(ts["ratio_connections"], f32p), _p(ts["total_path_length"], i32p), _p(ts["obs_step_count"], i32p), C.c_int(nthreads))
    return state, ts


def random_actions_batch(state: Dict[str, np.ndarray], nthreads: int = 0) -> np.ndarray:
    B, G, _ = state["grid"].shape
    N = state["target"].shape[1]
    action = np.empty((B, N), np.int32)
    lib().orc_random_actions_batch(C.c_int64(B), C.c_int(G), C.c_int(N), _p(state["grid"], i32p), _p(state["target"], i32p), _p(state["position"], i32p), _p(state["step_count"], i32p), _p(state["key"], u32p), _p(action, i32p), C.c_int(nthreads))
    return action


def validate_batch(boards, N: int, nthreads: int = 0) -> np.ndarray:
    boards = _i32(boards)
    if boards.ndim == 2:
        boards = boards[None]
    B, G, _ = boards.shape
    flags = np.empty((B,), np.int32)
    lib().orc_validate_batch(C.c_int64(B), C.c_int(G), C.c_int(N), _p(boards, i32p), _p(flags, i32p), C.c_int(nthreads))
    return flags


def board_statistics_batch(boards, count_current_wire: bool = False, nthreads: int = 0):
    """EvaluateEmptyBoard over boards[B,G,G] -> scored[B,G,G], count_detours[B], heatmap_score_diversity[B]."""
    boards = _i32(boards)
    if boards.ndim == 2:
        boards = boards[None]
    B, G, _ = boards.shape
    scored = np.empty((B, G, G), np.int32)
    det = np.empty((B,), np.int32)
    div = np.empty((B,), np.int32)
    lib().orc_board_statistics_batch(C.c_int64(B), C.c_int(G), _p(boards, i32p), C.c_int(int(count_current_wire)), _p(scored, i32p), _p(det, i32p), _p(div, i32p), C.c_int(nthreads))
    return scored, det, div


def max_threads() -> int:
    return int(lib().orc_max_threads())
