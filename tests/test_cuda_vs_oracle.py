"""GPU parity, part 2: CUDA vs the C oracle on identical seeded inputs at benchmark-like sizes
(BASELINE.json configs[1]: Small, 4096 envs, random actions, full 200-step episodes), in both RNG
modes, plus the fused greedy/auto-reset/stats paths and shard invariance."""
import numpy as np
import pytest
import torch

import golden_util as gu
from oracle import wh_oracle as wo

pytestmark = pytest.mark.gpu

SIZES = {"small": 4096, "medium": 1024, "large": 512}


def pair(size, n, seed, num_agents=None, train=False, auto_reset=False, env_id0=0):
    from rllib_warehouse_b200 import VARIANTS, BatchedWarehouse
    cfg = VARIANTS[size].replace(random_num_agents=train)
    gpu = BatchedWarehouse(cfg, n, num_agents=num_agents, seed=seed, env_id0=env_id0, auto_reset=auto_reset)
    cpu = wo.OracleEnv(wo.variant_config(size, random_num_agents=train), n, num_agents=num_agents,
                       seed=seed, env_id0=env_id0)
    return gpu, cpu


def same_state(gpu, cpu, where):
    st = gpu.get_state()
    for k in gu.STATE_KEYS + ("episode", "acc"):
        assert np.array_equal(st[k].reshape(cpu.state[k].shape), cpu.state[k]), f"{where}: state {k}"


def same_obs(gpu, cpu, where):
    for k in gu.OBS_KEYS:
        got = gpu.obs[k].cpu().numpy()
        assert np.array_equal(got.astype(np.int32), cpu.obs[k].astype(np.int32)), f"{where}: obs {k}"


@pytest.mark.parametrize("size", list(SIZES))
def test_native_rng_random_actions_full_episode(size):
    """configs[1] shape: every env, every step, state + obs + rewards + dones identical."""
    n = SIZES[size]
    gpu, cpu = pair(size, n, seed=0xC0FFEE)
    gpu.reset(); cpu.reset()
    same_state(gpu, cpu, "reset"); same_obs(gpu, cpu, "reset")
    rng = np.random.Generator(np.random.PCG64(5))
    R = cpu.R
    for t in range(205):
        actions = rng.integers(-1, 9, size=(n, R)).astype(np.int32)   # includes absent agents
        obs, rew, dones = gpu.step(actions)
        cpu.step(actions)
        same_state(gpu, cpu, f"step {t}")
        assert np.array_equal(rew.cpu().numpy(), cpu.rewards), f"step {t}: rewards"
        assert np.array_equal(dones.cpu().numpy(), cpu.dones), f"step {t}: dones"
        if t % 10 == 0 or t >= 198:
            same_obs(gpu, cpu, f"step {t}")
    assert np.array_equal(gpu.stats.cpu().numpy(), cpu.stats)
    assert int(cpu.stats[0]) == n


@pytest.mark.parametrize("size", list(SIZES))
def test_replayed_draws(size):
    """Replay mode at scale: the oracle's realised spawns are fed to the kernels as draw tensors."""
    n = SIZES[size] // 4
    gpu, cpu = pair(size, n, seed=31337, num_agents=None)
    cpu.reset()
    wait = cpu.cfg.pickup_wait_duration
    R = cpu.R

    def spawned():
        sp = np.full((n, R), -1, np.int32); st = np.full((n, R), -1, np.int32)
        for e in range(n):
            idx = np.nonzero(cpu.state["pickup_timer"][e] == wait)[0]
            sp[e, :len(idx)] = idx; st[e, :len(idx)] = cpu.state["pickup_tgt"][e, idx]
        return sp, st

    sp, st = spawned()
    gpu.reset(agent_pos=cpu.state["agent_pos"], init_pickups=sp, init_targets=st,
              num_agents=cpu.state["num_agents"])
    same_state(gpu, cpu, "reset")
    rng = np.random.Generator(np.random.PCG64(9))
    for t in range(60):
        cpu.greedy()
        actions = cpu.actions.copy()
        flip = rng.random((n, R)) < 0.2
        actions[flip] = rng.integers(0, 9, size=int(flip.sum()))
        order = np.stack([rng.permutation(R) for _ in range(n)]).astype(np.int32)
        cpu.step(actions, order=order)
        sp, st = spawned()
        gpu.step(actions, order=order, spawn_pickups=sp, spawn_targets=st)
        same_state(gpu, cpu, f"step {t}")
        same_obs(gpu, cpu, f"step {t}")


@pytest.mark.parametrize("size", list(SIZES))
def test_train_variant_greedy_autoreset_stats(size):
    """*Train (per-env random agent counts), solver kernel on obs, fused greedy+step kernel,
    in-kernel auto-reset across episode boundaries and the episode statistics."""
    n = SIZES[size] // 2
    gpu, cpu = pair(size, n, seed=77, train=True, auto_reset=True)
    gpu2, _ = pair(size, n, seed=77, train=True, auto_reset=True)
    gpu.reset(); gpu2.reset(); cpu.reset()
    same_state(gpu, cpu, "reset")
    assert len(np.unique(cpu.state["num_agents"])) > 1
    for t in range(420):
        acts = gpu.greedy_actions()                       # solver kernel on the resident obs
        cpu.greedy()
        assert np.array_equal(acts.cpu().numpy(), cpu.actions), f"step {t}: solver actions"
        gpu.step(acts)
        gpu2.greedy_step()                                # same thing in one fused kernel
        assert torch.equal(gpu2.actions, acts), f"step {t}: fused solver actions"
        cpu.step(cpu.actions, with_obs=False)
        done = cpu.dones.astype(bool)
        rew_cpu = cpu.rewards.copy()
        if done.any():                                    # oracle-side auto reset
            cpu.build_obs(0)
            cpu.reset(env_mask=done.astype(np.uint8))
            cpu.state["acc"][done] = 0
            ob_reset = {k: v.copy() for k, v in cpu.obs.items()}
            cpu.build_obs(0)
            for k in cpu.obs:
                cpu.obs[k][done] = ob_reset[k][done]
        else:
            cpu.build_obs(0)
        for g in (gpu, gpu2):
            same_state(g, cpu, f"step {t}")
            assert np.array_equal(g.rewards.cpu().numpy(), rew_cpu)
            assert np.array_equal(g.dones.cpu().numpy(), cpu.dones)
            if t % 20 == 0 or done.any():
                same_obs(g, cpu, f"step {t}")
    for g in (gpu, gpu2):
        assert np.array_equal(g.stats.cpu().numpy(), cpu.stats)
    assert int(cpu.stats[0]) == 2 * n
    sd = gpu.stats_dict()
    assert sd["episodes"] == 2 * n and sd["return_sum"] == sd["pickups"] + sd["deliveries"]


def test_shard_invariance():
    """N envs on one shard == the same global env ids split into shards (RNG keyed by global id)."""
    from rllib_warehouse_b200 import LARGE, BatchedWarehouse
    n = 256
    whole = BatchedWarehouse(LARGE, n, seed=5, auto_reset=True)
    parts = [BatchedWarehouse(LARGE, n // 4, seed=5, env_id0=i * (n // 4), auto_reset=True) for i in range(4)]
    whole.reset()
    for p in parts:
        p.reset()
    for _ in range(230):
        whole.greedy_step()
        for p in parts:
            p.greedy_step()
    for k in whole.state:
        assert torch.equal(whole.state[k], torch.cat([p.state[k] for p in parts])), k
    for k in whole.obs:
        assert torch.equal(whole.obs[k], torch.cat([p.obs[k] for p in parts])), k
    assert torch.equal(whole.stats, sum(p.stats for p in parts))


def test_generic_config_paths():
    """Non-variant geometry (irregular racks, R != {4,9,16}) goes through the runtime-R kernels."""
    from rllib_warehouse_b200 import BatchedWarehouse, WarehouseConfig
    for (R, dim, racks, A) in [(6, 14, (3, 7, 11), 5), (3, 11, (5,), 2), (12, 18, (4, 9, 14), 12), (20, 20, (4, 8, 12, 16), 17)]:
        cfg = WarehouseConfig(R, dim, racks, 40, 25, R)
        n = 300
        gpu = BatchedWarehouse(cfg, n, num_agents=A, seed=3, auto_reset=True)
        cpu = wo.OracleEnv(wo.make_config(R, dim, list(racks), 40, 25), n, num_agents=A, seed=3)
        gpu.reset(); cpu.reset()
        same_state(gpu, cpu, f"R={R} reset"); same_obs(gpu, cpu, f"R={R} reset")
        rng = np.random.Generator(np.random.PCG64(R))
        for t in range(39):
            if t % 2:
                actions = rng.integers(0, 9, size=(n, R)).astype(np.int32)
            else:
                actions = gpu.greedy_actions().cpu().numpy()
                assert np.array_equal(actions, cpu.greedy()), f"R={R} step {t} solver"
            gpu.step(actions); cpu.step(actions)
            same_state(gpu, cpu, f"R={R} step {t}"); same_obs(gpu, cpu, f"R={R} step {t}")
            assert np.array_equal(gpu.rewards.cpu().numpy(), cpu.rewards)


def test_philox_matches_oracle():
    """The device RNG against Random123 known answers is covered through the oracle (CPU test);
    here: first native reset of 1 env must equal the oracle's, for several seeds/env ids."""
    from rllib_warehouse_b200 import SMALL, BatchedWarehouse
    for seed, eid in [(0, 0), (2**63 + 12345, 7), (0xFFFFFFFFFFFFFFFF, 2**31 + 5)]:
        g = BatchedWarehouse(SMALL, 3, seed=seed, env_id0=eid)
        c = wo.OracleEnv(wo.variant_config("small"), 3, seed=seed, env_id0=eid)
        g.reset(); c.reset()
        same_state(g, c, f"seed {seed}")


def oracle_flat(cpu):
    """The ORACLE's dict observations in RLlib's Dict-flattening order (alphabetical keys, core.py:119-148),
    float32 [N, R, 9R+1] — the comparand of every flat-observation test (numpy only, no CUDA involved)."""
    n, R = cpu.N, cpu.R
    return np.concatenate([cpu.obs[k].reshape(n, R, -1).astype(np.float32) for k in sorted(cpu.obs)], axis=2)


def oracle_auto_reset(cpu, done):
    """What WH_FLAG_AUTO_RESET does in-kernel, on the oracle: finished envs are reset and show their
    reset-flavour observation, everything else keeps the step-flavour one."""
    cpu.build_obs(0)
    if done.any():
        step_obs = {k: v.copy() for k, v in cpu.obs.items()}
        cpu.reset(env_mask=done.astype(np.uint8))          # rebuilds every env's obs in reset flavour
        cpu.state["acc"][done] = 0
        for k in cpu.obs:
            cpu.obs[k][~done] = step_obs[k][~done]


@pytest.mark.parametrize("size", list(SIZES))
def test_flat_observations(size):
    """RLlib-flattened float32 observations (wh_build_obs_flat) == the ORACLE's dict observations
    flattened in alphabetical key order, for both flavours, per-env agent counts, batch sizes that leave
    partial warps, and the runtime-R path."""
    from rllib_warehouse_b200 import _native as nv
    for n in (1000, 37, 1):
        gpu, cpu = pair(size, n, seed=4, train=True)
        gpu.reset(); cpu.reset()
        flat = gpu.build_obs_flat(nv.OBS_RESET)
        assert flat.shape == (n, gpu.R, 9 * gpu.R + 1) and flat.dtype == torch.float32
        assert np.array_equal(flat.cpu().numpy(), oracle_flat(cpu)), "reset flavour"
        for t in range(30 if n == 1000 else 6):
            gpu.greedy_step(with_obs=False)
            cpu.greedy(); cpu.step(cpu.actions)
            if t % 5 == 0:
                assert np.array_equal(gpu.build_obs_flat().cpu().numpy(), oracle_flat(cpu)), (n, t)
                # the oracle also agrees with a plain flatten of the CUDA dict observations
                gpu.build_obs()
                assert np.array_equal(gpu.flatten_obs(gpu.obs).cpu().numpy(), oracle_flat(cpu)), (n, t)
    from rllib_warehouse_b200 import BatchedWarehouse, WarehouseConfig
    for R, dim, racks in [(6, 14, (3, 7, 11)), (3, 11, (5,))]:           # runtime-R path
        gpu = BatchedWarehouse(WarehouseConfig(R, dim, racks, 40, 25, R), 77, seed=1)
        cpu = wo.OracleEnv(wo.make_config(R, dim, list(racks), 40, 25, R), 77, seed=1)
        gpu.reset(); cpu.reset()
        assert np.array_equal(gpu.build_obs_flat(nv.OBS_RESET).cpu().numpy(), oracle_flat(cpu))
        gpu.greedy_step(with_obs=False)
        cpu.greedy(); cpu.step(cpu.actions)
        assert np.array_equal(gpu.build_obs_flat().cpu().numpy(), oracle_flat(cpu))


@pytest.mark.parametrize("compact,chunks", [(False, 3), (True, 3), (False, 0), (True, 0)])
def test_host_buffer_layer(compact, chunks):
    """Layer 2 of the C ABI (wh_env_*: host buffers; chunked copy/compute pipeline, the direct mode in which
    the kernel reads / writes the page-locked host buffers itself; both wire formats) gives exactly what the
    device-pointer layer gives on the same seed and actions."""
    import ctypes as C
    from rllib_warehouse_b200 import MEDIUM, BatchedWarehouse
    from rllib_warehouse_b200 import _native as nv
    L = nv.lib()
    n, R, seed, id0 = 5003, 9, 321, 1000
    h = C.c_void_p()
    cfg = nv.make_config(MEDIUM)
    nv.check(L.wh_env_create(C.byref(cfg), n, 0, id0, seed, chunks, C.byref(h)), "create")
    nv.check(L.wh_env_reset(h), "reset")
    twin = BatchedWarehouse(MEDIUM, n, seed=seed, env_id0=id0, auto_reset=True)
    twin.reset()
    rng = np.random.Generator(np.random.PCG64(3))
    adt, rdt = (np.int8, np.uint8) if compact else (np.int32, np.float32)
    rewards = torch.zeros((n, R), dtype=torch.uint8 if compact else torch.float32).pin_memory()
    dones = torch.zeros(n, dtype=torch.uint8).pin_memory()
    obs_host = {k: torch.zeros_like(v, device="cpu").pin_memory() for k, v in twin.obs.items()}
    for t in range(205):
        acts = torch.from_numpy(rng.integers(-1, 9, size=(n, R)).astype(adt)).pin_memory()
        if compact:
            nv.check(L.wh_env_step_host_compact(h, acts.data_ptr(), rewards.data_ptr(), dones.data_ptr()), "step")
        else:
            oh = nv.Obs(**{k: v.data_ptr() for k, v in obs_host.items()}) if t % 50 == 0 else None
            nv.check(L.wh_env_step_host(h, acts.data_ptr(), rewards.data_ptr(), dones.data_ptr(),
                                        C.byref(oh) if oh is not None else None), "step")
        twin.step(acts.to(torch.int32))
        assert torch.equal(rewards.to(torch.float32), twin.rewards.cpu()), t
        assert torch.equal(dones, twin.dones.cpu()), t
        if not compact and t % 50 == 0:
            for k in obs_host:
                assert torch.equal(obs_host[k], twin.obs[k].cpu()), (t, k)
    stats = (C.c_ulonglong * nv.NUM_STATS)()
    nv.check(L.wh_env_stats_host(h, stats), "stats")
    assert list(stats) == twin.stats.cpu().tolist() and stats[0] == n
    st = nv.State()
    nv.check(L.wh_env_state_ptrs(h, C.byref(st)), "state ptrs")
    assert st.agent_pos and L.wh_env_launch_count(h) == 1 + 205 * (chunks if chunks > 0 else 1)   # direct: one kernel per step
    L.wh_env_destroy(h)


@pytest.mark.parametrize("size", list(SIZES))
def test_step_with_fused_flat_observations(size):
    """wh_step_flat (the step kernel emitting RLlib-flattened float32 observations itself) against the
    ORACLE (step + dict observations flattened alphabetically), with per-env agent counts, random dict
    orders and absent agents, across auto-reset boundaries (finished envs show their reset-flavour
    observation), at batch sizes with partial warps."""
    from rllib_warehouse_b200 import VARIANTS, BatchedWarehouse
    for n in (777, 5):
        cfg = VARIANTS[size].replace(random_num_agents=True, episode_duration=12)
        b = BatchedWarehouse(cfg, n, seed=9, auto_reset=True)
        kw = dict(wo.VARIANTS[size]); kw["episode"] = 12
        cpu = wo.OracleEnv(wo.make_config(random_num_agents=True, **kw), n, seed=9)
        b.reset(); cpu.reset()
        rng = np.random.Generator(np.random.PCG64(2))
        for t in range(30):
            acts = rng.integers(-1, 9, size=(n, b.R)).astype(np.int32)
            order = np.stack([rng.permutation(b.R) for _ in range(n)]).astype(np.int32) if t % 4 == 0 else None
            flat, rew, dones = b.step_flat(acts, order=order)
            cpu.step(acts, order=order, with_obs=False)
            assert np.array_equal(rew.cpu().numpy(), cpu.rewards), t
            assert np.array_equal(dones.cpu().numpy(), cpu.dones), t
            oracle_auto_reset(cpu, cpu.dones.astype(bool))
            assert np.array_equal(flat.cpu().numpy(), oracle_flat(cpu)), (n, t)
            same_state(b, cpu, f"step {t}")
        assert np.array_equal(b.stats.cpu().numpy(), cpu.stats) and int(cpu.stats[0]) == 2 * n


def test_cuda_graph_replay_matches_eager():
    """StepGraph (CUDA-graph capture of several steps, for launch-bound batch sizes) is bit-identical
    to issuing the same launches eagerly — greedy policy and externally supplied actions."""
    from rllib_warehouse_b200 import SMALL, BatchedWarehouse, StepGraph
    n, T = 4096, 8
    eager = BatchedWarehouse(SMALL, n, seed=11, auto_reset=True)
    graphed = BatchedWarehouse(SMALL, n, seed=11, auto_reset=True)
    eager.reset(); graphed.reset()
    g = StepGraph(graphed, steps=T, policy="greedy")
    total = torch.zeros_like(eager.rewards)
    for rep in range(30):                       # 240 steps: crosses an episode boundary
        total.zero_()
        for _ in range(T):
            eager.greedy_step()
            total += eager.rewards
        g.replay()
        assert torch.equal(g.reward_sum, total), rep
    for k in eager.state:
        assert torch.equal(eager.state[k], graphed.state[k]), k
    for k in eager.obs:
        assert torch.equal(eager.obs[k], graphed.obs[k]), k
    assert torch.equal(eager.stats, graphed.stats) and graphed.launches == eager.launches
    ga = StepGraph(graphed, steps=T, policy="actions")
    rng = np.random.Generator(np.random.PCG64(1))
    for rep in range(5):
        acts = torch.from_numpy(rng.integers(-1, 9, size=(T, n, 4)).astype(np.int32)).cuda()
        ga.actions.copy_(acts)
        ga.replay()
        for t in range(T):
            eager.step(acts[t])
    for k in eager.state:
        assert torch.equal(eager.state[k], graphed.state[k]), k


def test_full_size_properties_and_oracle_prefix():
    """BASELINE configs[3] at its FULL size (Large, 262 144 envs, 205 greedy steps with auto-reset):
    size-independent invariants on the whole batch, and — because the RNG is keyed by the global
    env id — a bit-exact comparison of the first 2 048 envs with the CPU oracle."""
    from rllib_warehouse_b200 import LARGE, BatchedWarehouse
    n, m = 262144, 2048
    gpu = BatchedWarehouse(LARGE, n, seed=20261018, auto_reset=True)
    cpu = wo.OracleEnv(wo.variant_config("large"), m, seed=20261018)
    gpu.reset(); cpu.reset()
    R, dim = gpu.R, LARGE.area_dimension
    total_reward = torch.zeros((), dtype=torch.float64, device=gpu.device)
    for t in range(205):
        obs, rew, dones = gpu.greedy_step()
        total_reward += rew.sum()
        cpu.greedy(); cpu.step(cpu.actions)
        if cpu.dones.all():
            cpu.reset(env_mask=cpu.dones)
        if t % 40 == 0 or t in (198, 199, 200, 204):
            st = gpu.state
            assert bool(((st["pickup_tgt"] > -1).sum(dim=1) == R).all()), "exactly R active requests (core.py:338-351)"
            assert bool(((st["pickup_timer"] > 0) == (st["pickup_tgt"] > -1)).all())
            pos = st["agent_pos"]
            assert bool(((pos >= 0) & (pos < dim)).all()) and bool(((rew == 0) | (rew == 1)).all())
            assert bool((dones == dones[0]).all()) and bool((st["time"] == st["time"][0]).all())
            req = obs["requests"]
            assert bool((req[:, 0] == req[:, 5]).all()), "requests identical for every agent of an env (core.py:429)"
            cells = (req[:, 0, :, 0] * 32 + req[:, 0, :, 1]).sort(dim=1).values
            assert bool((cells.diff(dim=1) > 0).all()), "the R waiting requests sit on R distinct pickup cells"
            for k in ("agent_pos", "agent_tgt", "pickup_tgt", "pickup_timer", "time", "episode"):
                got = st[k][:m].to(torch.int32).cpu().numpy()
                assert np.array_equal(got.reshape(cpu.state[k].shape), cpu.state[k]), (t, k)
    s = gpu.stats.cpu().numpy()
    assert s[0] == n and s[1] == s[2] + s[3]
    acc = gpu.state["acc"].sum(dim=0).cpu().numpy()          # the 5 steps of the second episode
    assert float(total_reward.item()) == float(s[1] + acc[0] + acc[1])
    assert np.array_equal(cpu.stats[:5], gpu_prefix_stats(LARGE, m, 20261018))


def gpu_prefix_stats(cfg, m, seed):
    from rllib_warehouse_b200 import BatchedWarehouse
    g = BatchedWarehouse(cfg, m, seed=seed, auto_reset=True)
    g.reset()
    for _ in range(205):
        g.greedy_step(with_obs=False, want_actions=False)
    return g.stats[:5].cpu().numpy()


def test_limits_and_empty_batch():
    """Largest supported geometry (R = 32, P = D = 64: one env per warp) and an empty batch."""
    from rllib_warehouse_b200 import BatchedWarehouse, WarehouseConfig
    R, dim, racks = 32, 20, (4, 8, 12, 16)
    n = 65
    gpu = BatchedWarehouse(WarehouseConfig(R, dim, racks, 30, 7, R), n, num_agents=29, seed=8, auto_reset=True)
    cpu = wo.OracleEnv(wo.make_config(R, dim, list(racks), 30, 7), n, num_agents=29, seed=8)
    gpu.reset(); cpu.reset()
    same_state(gpu, cpu, "R=32 reset"); same_obs(gpu, cpu, "R=32 reset")
    rng = np.random.Generator(np.random.PCG64(5))
    for t in range(29):
        a = rng.integers(-1, 9, size=(n, R)).astype(np.int32) if t % 2 else gpu.greedy_actions().cpu().numpy()
        gpu.step(a); cpu.step(a)
        same_state(gpu, cpu, f"R=32 step {t}"); same_obs(gpu, cpu, f"R=32 step {t}")
        assert torch.equal(gpu.build_obs_flat(), gpu.flatten_obs(gpu.obs))
    empty = BatchedWarehouse(WarehouseConfig(4, 12, (4, 8)), 0)
    empty.reset(); empty.step(torch.zeros((0, 4), dtype=torch.int32)); empty.greedy_step(); empty.build_obs_flat()
    assert empty.obs["requests"].shape == (0, 4, 4, 4)


def test_random_geometries_fuzz():
    """Seeded fuzz over the whole supported configuration space (R 2..32, 1..4 racks at arbitrary —
    also adjacent, i.e. overlapping-cell — positions, dim up to 20, short waits/episodes, partial
    agent counts): native RNG, random absent agents and dict orders, auto-reset; state, obs,
    rewards, dones and stats must equal the C oracle's."""
    from rllib_warehouse_b200 import BatchedWarehouse, WarehouseConfig
    rng = np.random.Generator(np.random.PCG64(2026))
    tried = 0
    while tried < 14:
        L = int(rng.integers(1, 5))
        dim = int(rng.integers(9, 21))
        racks = sorted(rng.choice(np.arange(2, dim - 1), size=L, replace=False).tolist())
        P, D = 4 * L * L, 4 * (dim - 4)
        R = int(rng.integers(2, min(32, P, D) + 1))
        episode, wait = int(rng.integers(5, 30)), int(rng.integers(1, 12))
        A = int(rng.integers(1, R + 1))
        train = bool(rng.integers(0, 2))
        tried += 1
        n = int(rng.integers(1, 200))
        cfg = WarehouseConfig(R, dim, tuple(racks), episode, wait, R, train)
        gpu = BatchedWarehouse(cfg, n, num_agents=A, seed=tried, auto_reset=True)
        cpu = wo.OracleEnv(wo.make_config(R, dim, racks, episode, wait, random_num_agents=train), n, num_agents=A, seed=tried)
        gpu.reset(); cpu.reset()
        tag = f"cfg R={R} dim={dim} racks={racks} ep={episode} wait={wait} A={A} train={train} n={n}"
        same_state(gpu, cpu, tag + " reset"); same_obs(gpu, cpu, tag + " reset")
        for t in range(episode + 4):
            a = rng.integers(-1, 9, size=(n, R)).astype(np.int32)
            order = np.stack([rng.permutation(R) for _ in range(n)]).astype(np.int32) if t % 2 else None
            gpu.step(a, order=order)
            cpu.step(a, order=order, with_obs=False)
            rew, done = cpu.rewards.copy(), cpu.dones.astype(bool)
            cpu.build_obs(0)
            if done.any():
                step_obs = {k: v.copy() for k, v in cpu.obs.items()}
                cpu.reset(env_mask=done.astype(np.uint8))
                for k in cpu.obs:
                    step_obs[k][done] = cpu.obs[k][done]
                    cpu.obs[k][...] = step_obs[k]
            same_state(gpu, cpu, f"{tag} step {t}"); same_obs(gpu, cpu, f"{tag} step {t}")
            assert np.array_equal(gpu.rewards.cpu().numpy(), rew), f"{tag} step {t}"
            assert np.array_equal(gpu.dones.cpu().numpy().astype(bool), done), f"{tag} step {t}"
        assert np.array_equal(gpu.stats.cpu().numpy(), cpu.stats), tag


@pytest.mark.parametrize("size,n", [("small", 37), ("small", 4096), ("medium", 65536), ("large", 3000)])
def test_back_to_back_launches_match_synchronised_ones(size, n):
    """The step kernels are launched with programmatic stream serialization: step k+1's blocks may be
    scheduled while step k drains (tiny grids are fully co-resident). 450 launches issued back to back
    (fused greedy step and action-driven step alternating, across two auto-reset boundaries) must
    leave exactly what the same launches leave with a device synchronise after each one, and what
    the oracle computes."""
    from rllib_warehouse_b200 import VARIANTS, BatchedWarehouse
    fast = BatchedWarehouse(VARIANTS[size], n, seed=77, auto_reset=True)
    slow = BatchedWarehouse(VARIANTS[size], n, seed=77, auto_reset=True)
    fast.reset(); slow.reset()
    g = torch.Generator(device="cuda"); g.manual_seed(3)
    acts = [torch.randint(-1, 9, (n, fast.R), dtype=torch.int32, device="cuda", generator=g) for _ in range(8)]
    torch.cuda.synchronize()
    for t in range(450):                       # nothing but step kernels on the stream, no host sync
        if t % 3 == 0:
            fast.greedy_step(want_actions=False)
        else:
            fast.step(acts[t % 8])
    for t in range(450):
        if t % 3 == 0:
            slow.greedy_step(want_actions=False)
        else:
            slow.step(acts[t % 8])
        torch.cuda.synchronize()
    a, b = fast.get_state(), slow.get_state()
    for k in a:
        assert np.array_equal(a[k], b[k]), f"state {k}"
    for k in gu.OBS_KEYS:
        assert torch.equal(fast.obs[k], slow.obs[k]), f"obs {k}"
    assert torch.equal(fast.rewards, slow.rewards) and torch.equal(fast.dones, slow.dones)
    assert torch.equal(fast.stats, slow.stats) and int(fast.stats[0]) == 2 * n      # per-episode returns
    if n <= 4096:
        cpu = wo.OracleEnv(wo.variant_config(size), n, seed=77)
        cpu.reset()
        for t in range(450):
            if t % 3 == 0:
                cpu.greedy()
                cpu.step(cpu.actions)
            else:
                cpu.step(acts[t % 8].cpu().numpy())
            done = cpu.dones.astype(bool)
            if done.any():                      # what WH_FLAG_AUTO_RESET does in-kernel
                cpu.reset(env_mask=done.astype(np.uint8))
        same_state(fast, cpu, "after 450 back-to-back launches")


@pytest.mark.parametrize("size,n", [("small", 1001), ("medium", 1000), ("large", 333)])
def test_masked_reset_and_desynchronised_episodes(size, n):
    """Envs of one warp in different episode phases: (1) a masked reset rewrites the observations of
    exactly the masked envs and leaves their warp neighbours' untouched, (2) afterwards the envs end
    their episodes at different steps, so warps hold reset-flavour and step-flavour observations side
    by side in the auto-reset step. State, observations, rewards, dones vs the oracle throughout."""
    gpu, cpu = pair(size, n, seed=4242, auto_reset=True)
    gpu.reset(); cpu.reset()
    rng = np.random.Generator(np.random.PCG64(9))
    R = cpu.R

    def step_both(t):
        actions = rng.integers(0, 9, size=(n, R)).astype(np.int32)
        _, rew, dones = gpu.step(actions)
        cpu.step(actions)
        assert np.array_equal(rew.cpu().numpy(), cpu.rewards), f"step {t}: rewards"
        assert np.array_equal(dones.cpu().numpy(), cpu.dones), f"step {t}: dones"
        done = cpu.dones.astype(bool)
        if done.any():
            masked_reset_cpu(done.astype(np.uint8))          # what WH_FLAG_AUTO_RESET does in-kernel
        same_state(gpu, cpu, f"step {t}")
        same_obs(gpu, cpu, f"step {t}")

    def masked_reset_cpu(mask):
        # the oracle front-end rebuilds EVERY env's observation in reset flavour; only the masked
        # envs get one, the others keep the observation they had
        kept = {k: v.copy() for k, v in cpu.obs.items()}
        cpu.reset(env_mask=mask)
        for k in cpu.obs:
            cpu.obs[k][mask == 0] = kept[k][mask == 0]

    for t in range(60):
        step_both(t)
    before = {k: v.clone() for k, v in gpu.obs.items()}
    mask = (rng.random(n) < 0.4).astype(np.uint8)
    mask[:9] = [1, 0, 0, 0, 1, 0, 1, 1, 0]
    gpu.reset(env_mask=mask); masked_reset_cpu(mask)
    keep = torch.from_numpy(mask == 0).cuda()
    for k in gu.OBS_KEYS:
        assert torch.equal(gpu.obs[k][keep], before[k][keep]), f"masked reset touched a neighbour's '{k}'"
    same_state(gpu, cpu, "masked reset"); same_obs(gpu, cpu, "masked reset")
    for t in range(60, 270):                        # unmasked envs finish at 200, masked ones at 260
        step_both(t)
    assert np.array_equal(gpu.stats.cpu().numpy(), cpu.stats)


@pytest.mark.parametrize("size,n,p", [("small", 4097, 0.0), ("medium", 1000, 0.25), ("large", 515, 0.0),
                                      # batches past the launch-sized threshold: the 256-thread-block instantiation
                                      ("small", 19001, 0.1), ("medium", 7200, 0.0), ("large", 4801, 0.05)])
def test_multi_step_greedy_rollout_kernel(size, n, p):
    """wh_greedy_rollout: T solver+step iterations in one launch (state in registers, no per-step
    observations) == T wh_greedy_step launches: state, statistics, dones, reward sums; across two
    episode boundaries (in-kernel auto-reset), with and without the eps-random branch."""
    from rllib_warehouse_b200 import VARIANTS, BatchedWarehouse
    one = BatchedWarehouse(VARIANTS[size], n, seed=31, auto_reset=True)
    many = BatchedWarehouse(VARIANTS[size], n, seed=31, auto_reset=True)
    one.reset(); many.reset()
    total = torch.zeros_like(many.rewards)
    done_steps = 0
    for chunk in (1, 7, 192, 150, 100):                 # 450 steps: crosses t = 200 and t = 400
        rs = one.greedy_rollout(chunk, random_action_prob=p, solver_seed=5).clone()
        acc = torch.zeros_like(rs)
        for _ in range(chunk):
            many.greedy_step(random_action_prob=p, solver_seed=5, want_actions=False)
            acc += many.rewards
        done_steps += chunk
        assert torch.equal(rs, acc), f"reward sums after {done_steps} steps"
        assert torch.equal(one.dones, many.dones)
        a, b = one.get_state(), many.get_state()
        for k in a:
            assert np.array_equal(a[k], b[k]), f"state {k} after {done_steps} steps"
        for k in gu.OBS_KEYS:                           # rebuilt from the final state
            if int(many.dones.max()) == 0:              # (a just-reset env shows its reset flavour in `many`)
                assert torch.equal(one.obs[k], many.obs[k]), f"obs {k} after {done_steps} steps"
        total += rs
    assert torch.equal(one.stats, many.stats) and int(one.stats[0]) == 2 * n
    if p == 0.0 and n <= 4097:
        cpu = wo.OracleEnv(wo.variant_config(size), n, seed=31)
        cpu.reset()
        for t in range(450):
            cpu.greedy(); cpu.step(cpu.actions)
            done = cpu.dones.astype(bool)
            if done.any():
                cpu.reset(env_mask=done.astype(np.uint8))
        same_state(one, cpu, "rollout kernel vs oracle")
        assert np.array_equal(one.stats.cpu().numpy(), cpu.stats)


MULTI_KERNELS = ["throughput", "low_occupancy", "ws1", "ws2"]   # every wh_multi_step kernel (WH_FLAG_MULTI_KERNEL)


@pytest.mark.parametrize("kernel", MULTI_KERNELS)
@pytest.mark.parametrize("size,n,T", [("small", 4097, 205), ("medium", 1000, 60), ("large", 515, 40)])
def test_multi_step_open_loop_actions_every_step_vs_oracle(size, n, T, kernel):
    """wh_multi_step with an open-loop [T,N,R] action tensor and per-step outputs: ONE launch runs the whole
    BASELINE configs[1] episode (Small, 4 096(+1) envs, random actions incl. absent agents, T = 205 passes
    the step-200 mass expiry); every step's observations, rewards and dones — and the final state — are
    bit-identical to the oracle stepped T times."""
    gpu, cpu = pair(size, n, seed=0xBEEF)
    gpu.reset(); cpu.reset()
    rng = np.random.Generator(np.random.PCG64(17))
    actions = rng.integers(-1, 9, size=(T, n, cpu.R)).astype(np.int32)
    obs, rew, dones = gpu.multi_step(T, actions=actions, per_step=True, kernel=kernel)
    obs = {k: v.cpu().numpy() for k, v in obs.items()}
    rew, dones = rew.cpu().numpy(), dones.cpu().numpy()
    for t in range(T):
        cpu.step(actions[t])
        assert np.array_equal(rew[t], cpu.rewards), f"step {t}: rewards"
        assert np.array_equal(dones[t], cpu.dones), f"step {t}: dones"
        for k in gu.OBS_KEYS:
            assert np.array_equal(obs[k][t].astype(np.int32), cpu.obs[k].astype(np.int32)), f"step {t}: obs {k}"
    same_state(gpu, cpu, "after the launch")
    assert np.array_equal(gpu.stats.cpu().numpy(), cpu.stats)


@pytest.mark.parametrize("kernel", MULTI_KERNELS + ["auto"])
@pytest.mark.parametrize("size,n,p", [("small", 4099, 0.0), ("medium", 1000, 0.25), ("large", 515, 0.1)])
def test_multi_step_greedy_equals_single_launches(size, n, p, kernel):
    """wh_multi_step with the in-kernel greedy solver, auto-reset and observations written every step over
    the resident tensors == the same number of wh_greedy_step launches: state, last observations
    (reset-flavour for envs that just finished), last dones, per-agent reward sums, statistics — across two
    episode boundaries, in chunks whose ends fall before, on and after an episode end."""
    from rllib_warehouse_b200 import VARIANTS, BatchedWarehouse
    cfg = VARIANTS[size].replace(random_num_agents=True)
    one = BatchedWarehouse(cfg, n, seed=31, auto_reset=True)
    many = BatchedWarehouse(cfg, n, seed=31, auto_reset=True)
    one.reset(); many.reset()
    for chunk in (1, 150, 49, 200, 7):                       # ends at t = 1, 151, 200 (episode end), 400, 407
        total = torch.zeros((n, one.R), device=one.device)
        for _ in range(chunk):
            _, r, _ = one.greedy_step(random_action_prob=p, solver_seed=5)
            total += r
        _, sums, dones = many.multi_step(chunk, random_action_prob=p, solver_seed=5, kernel=kernel)
        assert torch.equal(sums, total), chunk
        assert torch.equal(dones, one.dones), chunk
        for k in one.state:
            assert torch.equal(one.state[k], many.state[k]), (chunk, k)
        for k in one.obs:
            assert torch.equal(one.obs[k], many.obs[k]), (chunk, k)
    assert torch.equal(one.stats, many.stats) and int(one.stats[0]) == 2 * n


@pytest.mark.parametrize("kernel", MULTI_KERNELS)
@pytest.mark.parametrize("size,n", [("small", 1029), ("medium", 334), ("large", 131)])
def test_multi_step_greedy_per_step_slices_equal_single_launches(size, n, kernel):
    """wh_multi_step with the in-kernel solver AND per-step outputs (every step's observations / rewards / dones in
    its own [t] slice), across an episode boundary with in-kernel auto-reset: slice t == what the t-th
    wh_greedy_step launch leaves in the resident tensors."""
    from rllib_warehouse_b200 import VARIANTS, BatchedWarehouse
    cfg = VARIANTS[size].replace(random_num_agents=True)
    one = BatchedWarehouse(cfg, n, seed=77, auto_reset=True)
    many = BatchedWarehouse(cfg, n, seed=77, auto_reset=True)
    one.reset(); many.reset()
    for _ in range(190):
        one.greedy_step(random_action_prob=0.1, solver_seed=9, want_actions=False)
    many.multi_step(190, random_action_prob=0.1, solver_seed=9, kernel=kernel)
    T = 25                                                # steps 191 .. 215: the episode ends at 200
    obs, rew, dones = many.multi_step(T, random_action_prob=0.1, solver_seed=9, per_step=True, kernel=kernel)
    for t in range(T):
        o, r, d = one.greedy_step(random_action_prob=0.1, solver_seed=9, want_actions=False)
        assert torch.equal(rew[t], r), f"step {t}: rewards"
        assert torch.equal(dones[t], d), f"step {t}: dones"
        for k in gu.OBS_KEYS:
            assert torch.equal(obs[k][t], o[k]), f"step {t}: obs {k}"
    for k in one.state:
        assert torch.equal(one.state[k], many.state[k]), k
    assert torch.equal(one.stats, many.stats) and int(one.stats[0]) == n


@pytest.mark.parametrize("keep_mb,what", [("0", "plain PLAIN kernels"), ("0.1", "partial evict_last policy: every state array but the timers (KEEP = 2)"),
                                          ("1000", "full evict_last policy (KEEP = 1)")])
def test_l2_keep_variants_are_bit_exact(keep_mb, what):
    """The throughput kernels exist in three L2-policy variants for the state accesses (plain / evict_last /
    evict_last for every array but the timers), selected from the state size (WH_KEEP_MAX_MB; WH_B200_KEEP_MB overrides, read once per
    process — hence a subprocess). All three must give the oracle's results: Medium and Large, random actions
    and the in-kernel solver, flat observations, across an episode boundary."""
    import os
    import subprocess
    import sys
    code = r"""
import sys, os
sys.path.insert(0, os.path.join(os.getcwd(), "tests")); sys.path.insert(0, os.getcwd())
import numpy as np, torch
import golden_util as gu
from oracle import wh_oracle as wo
from rllib_warehouse_b200 import VARIANTS, BatchedWarehouse
for size, n in (("medium", 1000), ("large", 400)):
    kw = dict(wo.VARIANTS[size]); kw["episode"] = 25
    cfg = VARIANTS[size].replace(episode_duration=25)
    gpu = BatchedWarehouse(cfg, n, seed=3, auto_reset=True)
    fused = BatchedWarehouse(cfg, n, seed=3, auto_reset=True)
    cpu = wo.OracleEnv(wo.make_config(**kw), n, seed=3)
    gpu.reset(); fused.reset(); cpu.reset()
    rng = np.random.Generator(np.random.PCG64(1))
    for t in range(60):
        a = rng.integers(0, 9, size=(n, cpu.R)).astype(np.int32)
        if t % 3 == 2:
            flat, rew, dones = gpu.step_flat(torch.from_numpy(a).cuda())
        else:
            _, rew, dones = gpu.step(torch.from_numpy(a).cuda())          # device int32 actions: the PLAIN path
        cpu.step(a, with_obs=False)
        assert np.array_equal(rew.cpu().numpy(), cpu.rewards) and np.array_equal(dones.cpu().numpy(), cpu.dones), (size, t)
        done = cpu.dones.astype(bool)
        cpu.build_obs(0)
        if done.any():
            keep = {k: v.copy() for k, v in cpu.obs.items()}
            cpu.reset(env_mask=done.astype(np.uint8)); cpu.state["acc"][done] = 0
            for k in cpu.obs: cpu.obs[k][~done] = keep[k][~done]
        st = gpu.get_state()
        for k in gu.STATE_KEYS + ("episode", "acc"):
            assert np.array_equal(st[k].reshape(cpu.state[k].shape), cpu.state[k]), (size, t, k)
        if t % 3 == 2:
            want = np.concatenate([cpu.obs[k].reshape(n, cpu.R, -1).astype(np.float32) for k in sorted(cpu.obs)], axis=2)
            assert np.array_equal(flat.cpu().numpy(), want), (size, t)
        else:
            for k in gu.OBS_KEYS:
                assert np.array_equal(gpu.obs[k].cpu().numpy().astype(np.int32), cpu.obs[k].astype(np.int32)), (size, t, k)
    twin = BatchedWarehouse(cfg, n, seed=3, auto_reset=True); twin.reset()
    for t in range(40):                                                     # in-kernel solver (K_GSTEP variants) vs solver kernel + step
        fused.greedy_step(want_actions=False)
        twin.step(twin.greedy_actions())
        for k in fused.state:
            assert torch.equal(fused.state[k], twin.state[k]), (size, t, k)
    assert torch.equal(fused.stats, twin.stats)
print("keep variants ok")
"""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, cwd=root,
                         env=dict(os.environ, WH_B200_KEEP_MB=keep_mb))
    assert res.returncode == 0 and "keep variants ok" in res.stdout, what + "\n" + res.stdout[-1000:] + res.stderr[-3000:]
