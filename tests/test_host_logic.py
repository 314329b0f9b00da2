"""Host-side logic that needs no GPU: shard bounds, the plugin surface (names,
constructor arguments, properties) and the multi-rank statistics gather over
gloo with world_size 2."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds_cover_everything():
    from routing_board_generation_b200 import sharding

    for total in (0, 1, 7, 8, 1024, 1048576, 1000003):
        for world in (1, 2, 3, 4, 8):
            spans = [sharding.shard_bounds(total, r, world) for r in range(world)]
            assert spans[0][0] == 0
            for (o0, c0), (o1, _) in zip(spans, spans[1:]):
                assert o0 + c0 == o1
            assert spans[-1][0] + spans[-1][1] == total
            counts = [c for _, c in spans]
            assert max(counts) - min(counts) <= 1


def test_sharded_split_equals_global_split(orc):
    """Each rank derives rows [offset, offset+count) of split(key, B) itself (setup_train.py:397-399)."""
    from routing_board_generation_b200 import sharding

    key = orc.PRNGKey(0)
    B = 4099
    full = orc.split(key, B)
    for world in (2, 8):
        parts = []
        for r in range(world):
            off, cnt = sharding.shard_bounds(B, r, world)
            parts.append(orc.split_slice(key, B, off, cnt))
        assert np.array_equal(np.concatenate(parts), full)


def test_plugin_surface_names():
    import routing_board_generation_b200 as pkg

    gen = pkg.ParallelRandomWalkGenerator(grid_size=10, num_agents=5)
    assert gen.grid_size == 10 and gen.num_agents == 5
    assert isinstance(gen, pkg.Generator)
    assert isinstance(gen.board_generator, pkg.ParallelRandomWalkBoard)
    assert (gen.board_generator.rows, gen.board_generator.cols, gen.board_generator.num_agents) == (10, 10, 5)
    assert isinstance(pkg.SeedExtensionGenerator(10, 5).board_generator, pkg.SeedExtensionBoard)
    assert isinstance(pkg.UniformRandomGenerator(10, 5), pkg.Generator)
    # sequential_random_walk_generator.py:19-31: same constructor, board_generator = SequentialRandomWalkBoard(size, size, n)
    sq = pkg.SequentialRandomWalkGenerator(grid_size=8, num_agents=3)
    assert isinstance(sq, pkg.Generator) and isinstance(sq.board_generator, pkg.SequentialRandomWalkBoard) and sq.kind == "sequential_random_walk"
    assert pkg.SequentialRandomWalkBoard(6, 6)._num_agents == 3  # sequential_random_walk.py:26 default num_agents=3
    for name in ("return_blank_board", "generate", "generate_starts_ends"):
        assert callable(getattr(sq.board_generator, name))
    with pytest.raises(ValueError):
        pkg.SequentialRandomWalkBoard(2, 2, 1)  # available_cells needs rows >= 3 (sequential_random_walk.py:138-140)
    assert pkg.Connector(generator=sq)._kind == "sequential_random_walk"  # reset / auto-reset inside the kernels, not the generic path
    # interface/board_generator_interface.py:45-46,55,64
    assert pkg.BoardGenerator.get_board_generator(pkg.BoardName("offline_parallel_rw")) is pkg.ParallelRandomWalkBoard
    assert pkg.BoardGenerator.get_board_generator(pkg.BoardName.JAX_SEED_EXTENSION) is pkg.SeedExtensionBoard
    env = pkg.Connector(generator=gen, time_limit=50)
    assert env.num_agents == 5 and env.grid_size == 10 and env.time_limit == 50
    for name in ("reset", "step", "_get_action_mask", "_obs_from_grid", "_get_extras"):
        assert callable(getattr(env, name))
    with pytest.raises(ValueError):
        pkg.ParallelRandomWalkBoard(4, 5, 2)


_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import torch, torch.distributed as dist
from routing_board_generation_b200 import sharding
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
r, w = sharding.rank_world()
off, cnt = sharding.shard_bounds(1001, r, w)
out = sharding.gather_stats({{"boards": cnt, "ms": 10.0 + r, "first": off}})
assert out["sum"]["boards"] == 1001, out
assert out["max"]["ms"] == 11.0 and out["min"]["ms"] == 10.0
assert out["max"]["first"] == 501
dist.barrier()
dist.destroy_process_group()
print("ok", r)
"""


def test_gather_stats_gloo_world2(tmp_path):
    port = 29500 + (os.getpid() % 2000)
    script = tmp_path / "w.py"
    script.write_text(_WORKER.format(root=ROOT, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
        assert "ok" in o


def test_bench_reference_arm_single_and_torchrun():
    """`bench.py --impl reference` prints one JSON line from rank 0; the other ranks exit 0 without work."""
    import json

    env = dict(os.environ, OMP_NUM_THREADS="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1", "--envs", "256"], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "connector_env_steps_per_sec" and line["value"] > 0
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["cpu_baseline"]["kind"] == "port" and line["gpu_launches"] == 0
    port = 29700 + (os.getpid() % 200)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", str(port),
                        os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1", "--envs", "128"], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1 and json.loads(lines[0])["n_gpus"] == 2
    # torchrun hands its workers OMP_NUM_THREADS=1 when the variable is unset: the CPU arm must still use
    # every core it may run on (it used to drop to one thread, and a slow one, at --gpus > 1)
    env1 = {k: v for k, v in os.environ.items() if k != "OMP_NUM_THREADS"}
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", str(port + 1),
                        os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1", "--envs", "128"], capture_output=True, text=True, env=env1, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][0])
    assert line["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))


def test_bench_product_arm_refuses_to_run_without_a_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], capture_output=True, text=True, timeout=300)
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)


def test_host_widen_matches_numpy():
    """rbg_host_widen (the host half of the byte transport of rbg_connector_step_host_io): bytes -> int32 on the
    library's thread pool; needs no GPU.  Ragged sizes and unaligned ends."""
    import ctypes as C

    import numpy as np

    import routing_board_generation_b200 as pkg

    lib = pkg._lib.load()
    rng = np.random.default_rng(0)
    for n in (0, 1, 31, 32, 33, 1000, 65536 * 5 + 7, 1 << 22):
        src = rng.integers(0, 97, size=n + 3, dtype=np.uint8)
        dst = np.full(n + 8, -1, np.int32)
        for off in (0, 1, 3):  # misaligned source
            assert lib.rbg_host_widen(C.c_void_p(src.ctypes.data + off), C.c_void_p(dst.ctypes.data), n - off if n >= off else 0) == 0
            m = n - off if n >= off else 0
            assert np.array_equal(dst[:m], src[off:off + m].astype(np.int32)) and (dst[m:m + 8 - 0][-1] == -1)
    assert lib.rbg_host_widen(None, None, 5) == -1


def test_host_widen4_matches_numpy():
    """rbg_host_widen4: two codes < 16 per byte (low nibble first) -> int32, on the library's thread pool; needs no GPU.
    Destinations at even and odd int32 offsets (the latter cannot use aligned streaming stores), ragged sizes."""
    import ctypes as C

    import numpy as np

    import routing_board_generation_b200 as pkg

    lib = pkg._lib.load()
    rng = np.random.default_rng(1)
    for n in (0, 2, 30, 32, 34, 1000, 65536 * 5 + 6, 1 << 22):
        codes = rng.integers(0, 16, size=n, dtype=np.uint8)
        packed = (codes[0::2] | (codes[1::2] << 4)).astype(np.uint8) if n else np.zeros(0, np.uint8)
        for doff in (0, 1, 2, 5):
            dst = np.full(n + 16, -1, np.int32)
            assert lib.rbg_host_widen4(C.c_void_p(packed.ctypes.data), C.c_void_p(dst.ctypes.data + 4 * doff), n) == 0
            assert np.array_equal(dst[doff:doff + n], codes.astype(np.int32)) and (dst[:doff] == -1).all() and (dst[doff + n:] == -1).all(), (n, doff)
    assert lib.rbg_host_widen4(None, None, 4) == -1 and lib.rbg_host_widen4(C.c_void_p(1), C.c_void_p(4), 3) == -1  # odd n
