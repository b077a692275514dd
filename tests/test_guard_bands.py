"""Out-of-bounds write detection without compute-sanitizer (closed on this GPU pool): every state,
observation and I/O tensor is re-homed into an arena with 0xA5 guard bands on both sides; after
running every kernel (tail warps, ghost lanes, runtime-R and compile-time-R variants) the guards
must be untouched."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GUARD = 1 << 16     # wide enough to catch a write one whole warp-tile past the end (dead tail warps)


def rehome(env):
    from rllib_warehouse_b200 import _native as nv
    arenas = []

    def move(t):
        nbytes = t.numel() * t.element_size()
        arena = torch.full((nbytes + 2 * GUARD,), 0xA5, dtype=torch.uint8, device=t.device)
        view = arena[GUARD:GUARD + nbytes].view(t.dtype).view(t.shape)
        view.copy_(t)
        arenas.append((arena, nbytes))
        return view

    for d in (env.state, env.obs):
        for k in list(d):
            d[k] = move(d[k])
    env.rewards, env.dones, env.actions, env.stats = move(env.rewards), move(env.dones), move(env.actions), move(env.stats)
    env._flat = move(torch.zeros((env.N, env.R, 9 * env.R + 1), dtype=torch.float32, device=env.device))
    env._st = nv.State(**{k: env.state[k].data_ptr() for k in nv.STATE_KEYS})
    env._ob = nv.Obs(**{k: env.obs[k].data_ptr() for k in nv.OBS_KEYS})
    return arenas


def guards_intact(arenas):
    for arena, nbytes in arenas:
        if not (bool((arena[:GUARD] == 0xA5).all()) and bool((arena[GUARD + nbytes:] == 0xA5).all())):
            return False
    return True


@pytest.mark.parametrize("n", [1, 37, 259])
def test_no_out_of_bounds_writes(n):
    from rllib_warehouse_b200 import VARIANTS, BatchedWarehouse, WarehouseConfig
    rng = np.random.Generator(np.random.PCG64(n))
    cfgs = [VARIANTS["small"], VARIANTS["medium"], VARIANTS["large"].replace(random_num_agents=True),
            WarehouseConfig(6, 14, (3, 7, 11), 30, 10, 6), WarehouseConfig(20, 20, (4, 8, 12, 16), 30, 10, 17)]
    for ci, cfg in enumerate(cfgs):
        # the warp-specialised wh_multi_step kernels exist for the reference variants' geometries only
        kerns = ("throughput", "low_occupancy") + (("ws1", "ws2") if ci < 3 else ())
        env = BatchedWarehouse(cfg, n, seed=7, auto_reset=True)
        arenas = rehome(env)
        env.reset()
        for t in range(40):
            a = rng.integers(-1, 9, size=(n, env.R)).astype(np.int32)
            order = np.stack([rng.permutation(env.R) for _ in range(n)]).astype(np.int32)
            env.step(a, order=order if t % 2 else None)
            env.step(env.greedy_actions(random_action_prob=0.3, solver_seed=1))
            env.greedy_step(random_action_prob=0.2)
            env.build_obs_flat()
            env.step_flat(a)
            env.build_obs(t % 2)
            if t % 8 == 0:
                for kern in kerns:
                    env.multi_step(3, kernel=kern)                             # greedy, observations every step
                    env.multi_step(2, actions=rng.integers(-1, 9, size=(2, n, env.R)).astype(np.int32), kernel=kern)
        env.reset(env_mask=(rng.random(n) < 0.5).astype(np.uint8))
        torch.cuda.synchronize()
        assert guards_intact(arenas), cfg
