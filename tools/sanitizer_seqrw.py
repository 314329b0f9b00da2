"""Small SequentialRandomWalk workloads for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool racecheck python tools/sanitizer_seqrw.py
Crowded shapes (retries, failed boards), ragged batch sizes, the generator State path and an auto-reset step."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import routing_board_generation_b200 as rbg  # noqa: E402
from oracle import oracle as orc  # noqa: E402

for G, N, B in ((3, 4, 40), (4, 6, 70), (6, 3, 33), (10, 5, 130), (14, 7, 50), (20, 10, 20)):
    keys = rbg.split(rbg.PRNGKey(7), B)
    kref = orc.split(orc.PRNGKey(7), B)
    board, stats = rbg.SequentialRandomWalkBoard(G, G, N).generate_with_stats(keys)
    ref, rstats = orc.seqrw_generate_batch(kref, G, N)
    assert np.array_equal(board.cpu().numpy(), ref) and np.array_equal(stats.cpu().numpy(), rstats), (G, N)
    st = rbg.SequentialRandomWalkGenerator(G, N)(keys)
    assert np.array_equal(st.grid.cpu().numpy(), orc.state_batch("sequential_random_walk", kref, G, N)["grid"])
env = rbg.VmapAutoResetWrapper(rbg.Connector(generator=rbg.SequentialRandomWalkGenerator(6, 3), time_limit=2))
st, ts = env.reset(rbg.split(rbg.PRNGKey(1), 64))
for _ in range(4):
    st, ts, _ = env.step_random(st)
torch.cuda.synchronize()
print("sanitizer workload ok")
