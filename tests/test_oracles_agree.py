"""The two independent CPU restatements (C: occupancy grid + invalid-move list; numpy: the
reference's own array formulation) must agree with each other far outside the golden fixtures:
random geometries (irregular racks, short waits so that requests expire mid-episode), random
agent counts, absent agents and random action-dict orders. The C oracle's native draws are
replayed into the numpy port."""
import numpy as np
import pytest

import golden_util as gu
from oracle import ref_port as rp
from oracle import wh_oracle as wo

GEOMETRIES = [
    (4, 12, (4, 8), 200, 200), (5, 11, (3, 7), 25, 6), (3, 9, (4,), 15, 3), (9, 16, (4, 8, 12), 30, 9),
    (7, 14, (3, 6, 10), 20, 5), (16, 20, (4, 8, 12, 16), 35, 12), (12, 19, (4, 9, 14), 28, 4),
]


@pytest.mark.parametrize("geo", GEOMETRIES)
def test_c_oracle_vs_numpy_port(geo):
    R, dim, racks, episode, wait = geo
    rng = np.random.Generator(np.random.PCG64(R * 1000 + dim))
    n = 24
    A = rng.integers(1, R + 1, size=n).astype(np.int32)
    c = wo.OracleEnv(wo.make_config(R, dim, list(racks), episode, wait), n, seed=R)
    c.state["num_agents"][:] = A
    c.reset()
    kw = dict(num_requests=R, area_dimension=dim, racks=list(racks), episode=episode, wait=wait)
    p = rp.PortEnv(kw, n, None)

    def spawned():
        sp = np.full((n, R), -1, np.int32); st = np.full((n, R), -1, np.int32)
        for e in range(n):
            idx = np.nonzero(c.state["pickup_timer"][e] == wait)[0]
            sp[e, :len(idx)] = idx; st[e, :len(idx)] = c.state["pickup_tgt"][e, idx]
        return sp, st

    sp, st = spawned()
    obs = p.reset(agent_pos=c.state["agent_pos"], init_pickups=sp, init_targets=st, num_agents=A)
    gu.assert_obs(c.obs, {"obs_" + k: obs[k] for k in gu.OBS_KEYS}, A, "reset")
    for t in range(episode + 3):
        actions = rng.integers(-1, 9, size=(n, R)).astype(np.int32)
        order = np.full((n, R), -1, np.int32)
        for e in range(n):
            order[e, :A[e]] = rng.permutation(A[e])
        c.step(actions, order=order)
        sp, st = spawned()
        obs, rew, dones = p.step(actions, order=order, spawn_pickups=sp, spawn_targets=st)
        for k in gu.STATE_KEYS:
            assert np.array_equal(p.state[k], c.state[k].reshape(p.state[k].shape)), (t, k)
        assert np.array_equal(rew, c.rewards) and np.array_equal(dones, c.dones), t
        gu.assert_obs(c.obs, {"obs_" + k: obs[k] for k in gu.OBS_KEYS}, A, f"step {t}")
    assert c.stats[4] > 0 or wait >= episode          # expiries happened when they could
