"""SeedExtension pipeline time per batch for batch sizes x resident extend CTAs per SM (one subprocess per setting)."""
import os, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import sys, torch
sys.path.insert(0, %r)
import routing_board_generation_b200 as rbg
import os
G, N = int(os.environ.get('SE_G', '14')), int(os.environ.get('SE_N', '7'))
for B in [int(x) for x in os.environ.get('SE_B', '65536,262144').split(',')]:
    keys = rbg.split(rbg.PRNGKey(0), B)
    board = rbg.SeedExtensionBoard(G, G, N)
    for _ in range(2):
        solved = board.return_solved_board(keys)
    torch.cuda.synchronize()
    rbg._lib.kernel_timing(True); rbg._lib.kernel_time("seedext")
    reps = 3
    for _ in range(reps):
        solved = board.return_solved_board(keys)
    torch.cuda.synchronize()
    n, kms = rbg._lib.kernel_time("seedext")
    rbg._lib.kernel_timing(False)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        solved = board.return_solved_board(keys)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print("B=%%d: %%.3f ms (kernels timed one by one: %%.3f)  %%.2f M boards/s  checksum %%d" %% (B, ms, kms / reps, B / ms / 1e3, int(solved.sum())), flush=True)
''' % root
settings = [{}] + [{"RBG_SE_CTAS_PER_SM": str(n)} for n in (6, 5, 4, 3, 2)] + [{"RBG_SE_EXT_WARPS": "1"}, {"RBG_SE_EXT_WARPS": "4"}]
if len(sys.argv) > 1:
    settings = [dict(kv.split("=") for kv in a.split(",") if kv) for a in sys.argv[1:]]
for env in settings:
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, **env), capture_output=True, text=True, timeout=900)
    print(env, "\n" + (r.stdout.strip() or r.stderr[-800:]), flush=True)
