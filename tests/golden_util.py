"""Shared golden-fixture checkers. They drive any "env-like" object — the C oracle front-end
(oracle.wh_oracle.OracleEnv), the numpy port (oracle.ref_port.PortEnv) or the CUDA product
(rllib_warehouse_b200.BatchedWarehouse) — through the same replay protocol and compare every
state array, observation key, reward and done flag bit-for-bit with what the unmodified
reference produced (tests/golden/*.npz, written by oracle/make_golden.py)."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SIZES = ("small", "medium", "large")
STATE_KEYS = ("agent_pos", "agent_tgt", "pickup_tgt", "pickup_timer", "time", "num_agents")
OBS_KEYS = (
    "num_agents", "self_position", "self_availability", "self_delivery_target",
    "other_positions", "other_availabilities", "other_delivery_targets", "requests",
)


def load(name):
    return dict(np.load(os.path.join(GOLDEN, name), allow_pickle=False))


def cfg_kwargs(d, prefix=""):
    return dict(num_requests=int(d[prefix + "R"]), area_dimension=int(d[prefix + "dim"]),
                racks=[int(r) for r in d[prefix + "racks"]], episode=int(d[prefix + "episode"]),
                wait=int(d[prefix + "wait"]))


def _np(x):
    if hasattr(x, "detach"):
        x = x.detach().cpu().numpy()
    return np.asarray(x)


def get_state(env):
    st = env.get_state() if hasattr(env, "get_state") else env.state
    return {k: _np(st[k]) for k in STATE_KEYS}


def assert_state(env, exp, where, envs=None):
    st = get_state(env)
    for k in STATE_KEYS:
        got = st[k] if envs is None else st[k][envs]
        want = np.asarray(exp[k]).astype(np.int64).reshape(got.shape)
        assert np.array_equal(got.astype(np.int64), want), f"{where}: state '{k}' differs\n{got}\n{want}"


def assert_obs(obs, exp, num_agents, where):
    """exp[k]: [N,R,...] with rows >= A equal to -1 (not produced by the reference)."""
    num_agents = np.asarray(num_agents).reshape(-1)
    for k in OBS_KEYS:
        got = _np(obs[k]).astype(np.int64)
        want = np.asarray(exp["obs_" + k]).astype(np.int64).reshape(got.shape)
        R = got.shape[1]
        live = (np.arange(R)[None, :] < num_agents[:, None])
        live = live.reshape(live.shape + (1,) * (got.ndim - 2))
        bad = (got != want) & live
        assert not bad.any(), (
            f"{where}: obs '{k}' differs at {np.argwhere(bad)[:5].tolist()}\n"
            f"got {got[bad][:8]} want {want[bad][:8]}")


def check_episode(make_env, d, pre):
    """Replay one recorded reference episode (keys prefixed `pre`) on a 1-env instance."""
    A = int(d[pre + "reset_num_agents"])
    env = make_env(cfg_kwargs(d, pre), 1, A)
    R = int(d[pre + "R"])
    obs = env.reset(agent_pos=d[pre + "reset_agent_pos"][None], init_pickups=d[pre + "reset_init_pickups"][None],
                    init_targets=d[pre + "reset_init_targets"][None], num_agents=np.array([A]))
    reset_exp = {k: d[pre + "reset_" + k][None] for k in STATE_KEYS}
    assert_state(env, reset_exp, pre + "reset")
    assert_obs(obs, {"obs_" + k: d[pre + "reset_obs_" + k][None] for k in OBS_KEYS}, [A], pre + "reset")
    T = d[pre + "actions"].shape[0]
    for t in range(T):
        obs, rew, dones = env.step(d[pre + "actions"][t][None], spawn_pickups=d[pre + "spawn_pickups"][t][None],
                                   spawn_targets=d[pre + "spawn_targets"][t][None])
        where = f"{pre}step{t}"
        assert_state(env, {k: d[pre + k][t][None] for k in STATE_KEYS}, where)
        assert np.array_equal(_np(rew).reshape(1, R)[:, :A], d[pre + "rewards"][t][None, :A]), where
        assert _np(rew).dtype == np.float32
        assert int(_np(dones).reshape(-1)[0]) == int(d[pre + "done"][t]), where
        assert_obs(obs, {"obs_" + k: d[pre + "obs_" + k][t][None] for k in OBS_KEYS}, [A], where)
    return T


def episode_prefixes(d):
    return sorted({k.split("_")[0] + "_" for k in d if k.endswith("_reset_time")})


def check_single_steps(make_env, d, names=None):
    """All recorded injected-state cases of a fixture as ONE batched step (per-env num_agents,
    per-env action order, absent agents)."""
    n = d["actions"].shape[0]
    env = make_env(cfg_kwargs(d), n, None)
    env.load_state(**{k: d["pre_" + k] for k in STATE_KEYS})
    obs, rew, dones = env.step(d["actions"], order=d["order"], spawn_pickups=d["spawn_pickups"],
                               spawn_targets=d["spawn_targets"])
    A = d["pre_num_agents"].astype(np.int64)
    st = get_state(env)
    for i in range(n):
        tag = f"case {i}" + (f" ({names[i]})" if names is not None else "")
        for k in STATE_KEYS:
            got, want = st[k][i].astype(np.int64), d[k][i].astype(np.int64).reshape(st[k][i].shape)
            assert np.array_equal(got, want), f"{tag}: state '{k}'\n got {got.tolist()}\nwant {want.tolist()}"
        a = int(A[i])
        assert np.array_equal(_np(rew)[i, :a], d["rewards"][i, :a]), f"{tag}: rewards"
        assert int(_np(dones)[i]) == int(d["done"][i]), f"{tag}: done"
    assert_obs(obs, {"obs_" + k: d["obs_" + k] for k in OBS_KEYS}, A, "single_steps")
    return n


def check_solver(greedy_fn, d):
    """greedy_fn(cfg_kwargs, obs_dict[N,R,...], num_agents[N], rand_prob, is_random, random_actions)
    -> actions [N,R]; compared with the reference solver's output (p=0 part and replayed eps part)."""
    n = d["actions"].shape[0]
    R = int(d["R"])
    obs = {k: d["obs_" + k] for k in OBS_KEYS}
    A = d["actions"].shape[1]
    num_agents = np.full(n, A, np.int32)
    pad = lambda a, fill: np.concatenate([a, np.full((n, R - A) + a.shape[2:], fill, a.dtype)], 1)
    got = greedy_fn(cfg_kwargs(d), obs, num_agents, float(d["rand_prob"]), pad(d["is_random"], 0),
                    pad(d["random_actions"], -1))
    assert np.array_equal(_np(got)[:, :A].astype(np.int64), d["actions"].astype(np.int64))
    return n


# ---- full-size reference digests (oracle/make_golden_batch.py) --------------------------------
def _crc(a, dtype):
    import zlib
    return zlib.crc32(np.ascontiguousarray(_np(a), dtype=dtype).tobytes())


def batch_actions(d):
    """The action tensor [T, N, A] of a `random` batch, regenerated from its recorded PCG64 seed
    (guarded by a CRC so that a changed numpy stream is reported as such, not as a parity bug)."""
    T, n, A = int(d["T"]), int(d["n"]), int(d["A"])
    acts = np.random.Generator(np.random.PCG64(int(d["seed_actions"]))).integers(0, 9, size=(T, n, A)).astype(np.int32)
    assert _crc(acts, np.int32) == int(d["actions_crc_all"]), "numpy PCG64 integers() stream changed"
    return acts


def check_batch_digests(make_env, d, greedy_fn=None, steps=None):
    """Replays a whole recorded reference batch (all N envs at once, T steps) on an env-like object
    and compares, at every step, the CRC-32 of every full [N, ...] output array — state, all
    observation keys, actions, rewards, dones — with what the UNMODIFIED reference produced.
    greedy_fn(env) -> actions [N, R] for `greedy` batches (the solver under test)."""
    n, A, R = int(d["n"]), int(d["A"]), int(d["R"])
    T = int(d["T"]) if steps is None else steps
    policy = str(d["policy"])
    # per-env agent counts (the *Train batches; all equal to A otherwise); rows >= num_agents are
    # padding the reference never produces: -1 in the digest
    num_agents = d["reset_num_agents"].astype(np.int32) if "reset_num_agents" in d else np.full(n, A, np.int32)
    pad_rows = np.arange(R)[None, :] >= num_agents[:, None]
    env = make_env(cfg_kwargs(d), n, None if pad_rows.any() else A)
    obs = env.reset(agent_pos=d["reset_agent_pos"], init_pickups=d["reset_init_pickups"],
                    init_targets=d["reset_init_targets"], num_agents=num_agents)

    def outputs(obs):
        out = dict(get_state(env))
        for k in OBS_KEYS:
            v = _np(obs[k]).astype(np.int32).copy()
            v[pad_rows] = -1
            out["obs_" + k] = v
        return out

    got = outputs(obs)
    for k, want in zip(d["reset_keys"], d["reset_crc"]):
        assert _crc(got[str(k)], np.int32) == int(want), f"reset: '{k}' differs from the reference"
    acts_all = batch_actions(d) if policy == "random" else None
    keys = [str(k) for k in d["out_keys"]]
    ret = np.zeros(n, np.float64)
    for t in range(T):
        acts = (_np(greedy_fn(env))[:, :R] if policy == "greedy" else acts_all[t]).astype(np.int32).copy()
        acts[pad_rows] = -1
        obs, rew, dones = env.step(acts, spawn_pickups=d["spawn_pickups"][t], spawn_targets=d["spawn_targets"][t])
        got = outputs(obs)
        got.update(actions=acts, rewards=_np(rew)[:, :R], dones=_np(dones))
        for j, k in enumerate(keys):
            dt = np.float32 if k == "rewards" else (np.uint8 if k == "dones" else np.int32)
            assert _crc(got[k], dt) == int(d["step_crc"][t, j]), f"step {t}: '{k}' differs from the reference"
        ret += _np(rew)[:, :R].sum(axis=1)
    if T == int(d["T"]):
        assert_state(env, {k: d["final_" + k] for k in STATE_KEYS}, "final state")
        assert np.array_equal(ret.astype(np.float32), d["return_per_env"])
    return n * T
