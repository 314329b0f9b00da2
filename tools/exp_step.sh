cd /root/repo
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "autoreset or connector_step or interleave or stepped_twice or workspace or host or mixed or dataset or composition or two_batches" 2>&1 | tail -3
python - <<'PY'
import sys, os, torch, time
sys.path.insert(0, "/root/repo")
import routing_board_generation_b200 as rbg
for trial in range(3):
    g, n, b = 10, 5, 65536
    env = rbg.VmapAutoResetWrapper(rbg.Connector(generator=rbg.ParallelRandomWalkGenerator(g, n), time_limit=50))
    st, _ = env.reset(rbg.split(rbg.PRNGKey(trial), b))
    ts1 = rbg.engine.alloc_timestep(b, g, n)
    def one():
        global st
        st, _, _ = rbg.engine.connector_step(st, None, 50, -0.03, 0.1, autoreset_kind="parallel_random_walk", inplace=True, random_policy=True, out=ts1, owner=env)
    for _ in range(160): one()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(400): one()
    e1.record(); t1 = time.perf_counter(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 400
    print(f"trial {trial}: {ms*1e3:.1f} us per step, {b/ms/1e3:.1f} M env-steps/s, host issue {1e6*(t1-t0)/400:.1f} us per step")
PY
