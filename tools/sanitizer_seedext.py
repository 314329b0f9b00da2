"""SeedExtension under compute-sanitizer: enough boards that lanes are refilled from the queue (RBG_SE_CTAS_PER_SM=1 outside),
two board sizes (one-word and two-word row masks), the overlapped optimise launch, validate, and one host-transport step."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import routing_board_generation_b200 as rbg
for (G, N, B) in ((8, 4, 700), (14, 7, 300), (34, 6, 40)):
    keys = rbg.split(rbg.PRNGKey(1), B)
    solved = rbg.SeedExtensionBoard(G, G, N).return_solved_board(keys)
    rbg.engine.validate(solved, N)
L, lib = rbg._lib, rbg._lib.load()
G, N, B = 10, 5, 4200
env = rbg.VmapAutoResetWrapper(rbg.Connector(generator=rbg.ParallelRandomWalkGenerator(G, N), time_limit=3))
st, _ = env.reset(rbg.split(rbg.PRNGKey(2), B))
h = dict(obs=np.empty((B, N, G, G), np.int32), mask=np.empty((B, N, 5), np.uint8), sc=np.empty(B, np.int32), reward=np.empty((B, N), np.float32), discount=np.empty((B, N), np.float32),
         step_type=np.empty(B, np.int8), nc=np.empty(B, np.int32), rc=np.empty(B, np.float32), tpl=np.empty(B, np.int32))
a = st.agents
s = L.rbg_state(st.grid.data_ptr(), st.step_count.data_ptr(), a.id.data_ptr(), a.start.data_ptr(), a.target.data_ptr(), a.position.data_ptr(), st.key.data_ptr())
t = L.rbg_timestep(*(h[k].ctypes.data for k in ("obs", "mask", "sc", "reward", "discount", "step_type", "nc", "rc", "tpl")))
params = L.rbg_env_params(3, -0.03, 0.1, 0)
act = np.random.default_rng(0).integers(0, 5, size=(B, N)).astype(np.int32)
for _ in range(2):
    L.check(lib.rbg_connector_step_host_io(C.byref(s), act.ctypes.data, B, G, N, C.byref(params), C.byref(t), -1))
torch.cuda.synchronize()
print("sanitizer seedext workload done")
