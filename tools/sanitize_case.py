"""Small workload that touches every kernel and code path (all three variants, runtime-R kernels,
replay + native RNG, custom order, auto-reset, flat obs, solver, host-buffer layer) — the target of
`compute-sanitizer --tool memcheck|racecheck` runs recorded under profiles/."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rllib_warehouse_b200 import VARIANTS, BatchedWarehouse, WarehouseConfig  # noqa: E402
from rllib_warehouse_b200 import _native as nv  # noqa: E402

rng = np.random.Generator(np.random.PCG64(0))
cfgs = [VARIANTS["small"], VARIANTS["medium"], VARIANTS["large"],
        VARIANTS["large"].replace(random_num_agents=True), WarehouseConfig(6, 14, (3, 7, 11), 30, 10, 6),
        WarehouseConfig(20, 20, (4, 8, 12, 16), 30, 10, 17)]
for cfg in cfgs:
    n = 37
    env = BatchedWarehouse(cfg, n, seed=1, auto_reset=True)
    env.reset()
    R = env.R
    for t in range(35):
        a = rng.integers(-1, 9, size=(n, R)).astype(np.int32)
        order = np.stack([rng.permutation(R) for _ in range(n)]).astype(np.int32) if t % 3 == 0 else None
        env.step(a, order=order)
        env.step(env.greedy_actions(random_action_prob=0.2, solver_seed=3))
        env.greedy_step(random_action_prob=0.1)
        env.build_obs_flat()
        env.build_obs(0)
    # replay mode: re-create the freshly reset state from "recorded draws", then one replayed step
    env2 = BatchedWarehouse(cfg, n, seed=2)
    env2.reset()
    st = env2.get_state()
    pick = np.stack([np.nonzero(st["pickup_tgt"][e] >= 0)[0] for e in range(n)])
    tg = np.stack([st["pickup_tgt"][e][pick[e]] for e in range(n)])
    env2.reset(agent_pos=st["agent_pos"], init_pickups=pick, init_targets=tg, num_agents=st["num_agents"])
    none = np.full((n, R), -1, np.int8)
    env2.step(np.full((n, R), 4, np.int32), spawn_pickups=none, spawn_targets=none)
    torch.cuda.synchronize()
L = nv.lib()
h = C.c_void_p()
ccfg = nv.make_config(VARIANTS["medium"])
nv.check(L.wh_env_create(C.byref(ccfg), 101, 0, 0, 5, 3, C.byref(h)), "create")
nv.check(L.wh_env_reset(h), "reset")
acts = torch.randint(0, 9, (101, 9), dtype=torch.int32).pin_memory()
rew = torch.zeros((101, 9)).pin_memory()
dn = torch.zeros(101, dtype=torch.uint8).pin_memory()
for _ in range(5):
    nv.check(L.wh_env_step_host(h, acts.data_ptr(), rew.data_ptr(), dn.data_ptr(), None), "step")
    nv.check(L.wh_env_greedy_step_host(h, rew.data_ptr(), dn.data_ptr()), "gstep")
L.wh_env_destroy(h)
torch.cuda.synchronize()
print("sanitize case ok")
