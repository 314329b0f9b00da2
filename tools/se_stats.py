"""Counters of se_extend_kernel (needs a -DRBG_SE_STATS build: RBG_NVCC_EXTRA=-DRBG_SE_STATS python
routing-board-generation_b200/build.py --force; the library is written to lib/ as usual, so rebuild afterwards)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import routing_board_generation_b200 as rbg
a = [int(x) for x in sys.argv[1:]]
G, N, B = (a + [14, 7, 65536][len(a):])[:3]
lib = rbg._lib.load()
keys = rbg.split(rbg.PRNGKey(0), B)
board = rbg.SeedExtensionBoard(G, G, N)
board.return_solved_board(keys)
buf = (C.c_ulonglong * 16)()
lib.rbg_debug_se_stats(buf, 1)
board.return_solved_board(keys)
lib.rbg_debug_se_stats(buf, 0)
v = list(buf)
ws = max(v[0], 1)
print(f"{G}x{G}/{N} B={B}: warp-sweeps {v[0]} ({v[8] / ws:.1f} lanes sweeping), lane-sweeps per board {v[8] / B:.2f}")
print(f"per warp-sweep: passes {v[1] / ws:.1f} ({v[1] / ws / G:.2f} per row), lanes with a cell per pass {v[2] / max(v[1], 1):.1f}, moves {v[7] / ws:.1f}")
print(f"per warp-sweep: warp_pick calls {v[3] / ws:.1f}, loop iterations {v[4] / ws:.1f}, lanes needing a pick {v[5] / ws:.1f}, picks that outran the parked keys {v[6] / ws:.2f}")
print(f"refill events {v[9]} retire events {v[10]} (per warp-sweep {v[9] / ws:.2f} / {v[10] / ws:.2f})")
