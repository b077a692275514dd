"""GPU tests of the reference-facing surface: the `warehouse` import path, the MultiAgentEnv
dict API (core.py:167,262), variant constructors (variants.py), the solver interface
(solvers.py:18-29) and the baseline driver loop (run.py:35-62)."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_multiagent_env_contract():
    import warehouse
    assert set(warehouse.__all__) == {"Warehouse", "WarehouseSmall", "WarehouseMedium", "WarehouseLarge",
                                      "WarehouseSmallTrain", "WarehouseMediumTrain", "WarehouseLargeTrain"}
    np.random.seed(3)
    for cls, A, R, dim in [(warehouse.WarehouseSmall, 3, 4, 12), (warehouse.WarehouseMedium, 9, 9, 16),
                           (warehouse.WarehouseLarge, 16, 16, 20)]:
        env = cls(A)
        assert (env.num_agents, env.num_requests) == (A, R)
        assert env.action_space.n == 9 and env.reward_range == (0.0, 1.0)
        assert env.animate_frames_per_step == 10 and env.metadata == {"render.modes": ["human"]}
        obs = env.reset()
        assert list(obs) == [str(i) for i in range(A)]
        for o in obs.values():
            assert env.observation_space.contains(o)
            assert o["self_availability"].dtype == np.int8 and o["requests"].dtype == np.int32
            assert o["requests"].shape == (R, 4) and o["other_positions"].shape == (R - 1, 2)
            assert int(o["self_availability"][0]) == 0                      # quirk 8
            assert o["self_delivery_target"].tolist() == [dim // 2] * 2
        obs, rew, dones, infos = env.step({str(i): 4 for i in range(A)})
        assert set(dones) == {str(i) for i in range(A)} | {"__all__"} and not dones["__all__"]
        assert all(type(r) is np.float32 for r in rew.values())
        assert infos == {str(i): {} for i in range(A)}
        for o in obs.values():
            assert env.observation_space.contains(o)
        for _ in range(199):
            obs, rew, dones, _ = env.step({str(i): int(np.random.randint(9)) for i in range(A)})
        assert dones["__all__"] and all(dones.values())
    with pytest.raises(AssertionError):
        warehouse.WarehouseSmall(5)
    with pytest.raises(IndexError):
        warehouse.WarehouseSmall(2).step({"0": 9})


def test_action_dict_order_is_semantic():
    """core.py:279: {"1":..,"0":..} != {"0":..,"1":..} when both want the same cell."""
    import warehouse
    res = []
    for keys in (("0", "1"), ("1", "0")):
        env = warehouse.WarehouseSmall(2, seed=1)
        env.reset()
        env._batched.load_state(agent_pos=np.array([[[5, 4], [5, 6], [-1, -1], [-1, -1]]]))
        acts = {"0": 5, "1": 3}
        obs, *_ = env.step({k: acts[k] for k in keys})
        res.append([obs["0"]["self_position"].tolist(), obs["1"]["self_position"].tolist()])
    assert res[0] == [[5, 5], [5, 6]] and res[1] == [[5, 4], [5, 5]]


def test_train_variants_redraw_agent_count():
    import warehouse
    np.random.seed(0)
    env = warehouse.WarehouseLargeTrain()
    seen = set()
    for _ in range(12):
        obs = env.reset()
        assert len(obs) == env.num_agents and 1 <= env.num_agents <= 16
        assert int(obs["0"]["num_agents"][0]) == env.num_agents
        seen.add(env.num_agents)
        obs, rew, dones, _ = env.step({str(i): 0 for i in range(env.num_agents)})
        assert len(rew) == env.num_agents
    assert len(seen) > 3


def test_baseline_driver_loop(capsys):
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    import importlib
    run = importlib.import_module("run")
    np.random.seed(5)
    total = run.run_single("small", 4, 0.0, False)
    out = capsys.readouterr().out
    assert "=== Done (200 steps) ===" in out and total > 0           # greedy policy collects rewards
    total_noisy = run.run_single("medium", 5, 0.5, False)
    assert total_noisy >= 0
    run.run_batched("large", 16, 0.0, 2048)
    out = capsys.readouterr().out
    assert "agent-steps/s" in out


def test_greedy_returns_match_reference_statistics():
    """Native-RNG mode is distribution-equivalent to the reference (SURVEY.md §6 [probe], 200
    reference episodes): greedy p=0 total return per episode 84.3+-48.7 / 102.0+-57.3 / 127.3+-47.9."""
    import torch
    from rllib_warehouse_b200 import VARIANTS, BatchedWarehouse
    ref = {"small": (84.3, 48.7), "medium": (102.0, 57.3), "large": (127.3, 47.9)}
    for size, (mean, std) in ref.items():
        env = BatchedWarehouse(VARIANTS[size], 8192, seed=123)
        env.reset()
        for _ in range(200):
            env.greedy_step(with_obs=False, want_actions=False)
        ret = (env.state["acc"][:, 0] + env.state["acc"][:, 1]).double()
        assert abs(ret.mean().item() - mean) < 4 * std / np.sqrt(200), (size, ret.mean().item())
        assert abs(ret.std().item() - std) < 0.2 * std, (size, ret.std().item())


@pytest.mark.parametrize("flat", [False, True])
def test_vector_env_base_env_protocol(flat):
    """poll / send_actions / try_reset (RLlib BaseEnv protocol) against a BatchedWarehouse twin."""
    import torch
    from rllib_warehouse_b200 import MEDIUM, BatchedWarehouse, WarehouseVectorEnv
    from rllib_warehouse_b200 import _native as nv
    n, A = 7, 5
    venv = WarehouseVectorEnv(MEDIUM, n, num_agents=A, seed=42, flat_obs=flat)
    twin = BatchedWarehouse(MEDIUM, n, num_agents=A, seed=42)
    twin.reset()

    def expect(e, i, flavour):
        if flat:
            return twin.build_obs_flat(flavour)[e, i].cpu().numpy()
        return {k: twin.obs[k][e, i].cpu().numpy() for k in twin.obs}

    def same(a, b):
        if flat:
            return np.array_equal(a, b)
        return set(a) == set(b) and all(np.array_equal(a[k], b[k]) for k in a)

    obs, rew, dones, infos, off = venv.poll()
    assert sorted(obs) == list(range(n)) and off == {}
    assert all(rew[e][str(i)] is None and not dones[e]["__all__"] for e in obs for i in range(A))
    assert all(same(obs[e][str(i)], expect(e, i, nv.OBS_RESET)) for e in range(n) for i in range(A))
    rng = np.random.Generator(np.random.PCG64(0))
    resets = 0
    for t in range(203):
        acts = {e: {str(i): int(rng.integers(0, 9)) for i in (range(A) if e % 2 else reversed(range(A)))} for e in range(n)}
        venv.send_actions(acts)
        a = np.full((n, twin.R), -1, np.int32); order = np.full((n, twin.R), -1, np.int32)
        for e in range(n):
            for k, (ag, v) in enumerate(acts[e].items()):
                a[e, int(ag)] = v; order[e, k] = int(ag)
        twin.step(a, order=order)
        obs, rew, dones, infos, _ = venv.poll()
        assert venv.poll()[0] == {}                                   # nothing pending until the next send
        for e in range(n):
            assert dones[e]["__all__"] == bool(twin.dones[e].item())
            for i in range(A):
                assert rew[e][str(i)] == twin.rewards[e, i].item()
                assert same(obs[e][str(i)], expect(e, i, nv.OBS_STEP)), (t, e, i)
            if dones[e]["__all__"] and e < 3:                          # reset some of the finished envs
                mask = np.zeros(n, np.uint8); mask[e] = 1
                twin.reset(env_mask=mask)
                fresh = venv.try_reset(e)
                assert all(same(fresh[str(i)], expect(e, i, nv.OBS_RESET)) for i in range(A))
                resets += 1
    assert resets >= 3
    # tensor API
    o = venv.reset_tensors()
    o2, r2, d2 = venv.step_tensors(torch.zeros((n, twin.R), dtype=torch.int32, device="cuda"))
    assert r2.shape == (n, twin.R) and d2.shape == (n,)


def test_batched_rollout_driver_with_torch_policy(tmp_path):
    """scripts/rollout_batched.py: greedy policy and a TorchScript policy fed by the flat-obs kernel."""
    import subprocess
    import torch
    R = 4

    class Tiny(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.l = torch.nn.Linear(9 * R + 1, 9)

        def forward(self, x):
            return self.l(x / 12.0)

    torch.manual_seed(0)
    path = str(tmp_path / "policy.pt")
    torch.jit.script(Tiny()).save(path)
    script = os.path.join(ROOT, "scripts", "rollout_batched.py")
    for extra in ([], ["--policy", path], ["--train-variant"]):
        res = subprocess.run([sys.executable, script, "small", "--envs", "2048", "--episodes", "1", *extra],
                             capture_output=True, text=True, timeout=300)
        assert res.returncode == 0, res.stderr[-3000:]
        assert "episodes: 2048" in res.stdout and "avg_agent_reward_all" in res.stdout, res.stdout


@pytest.mark.parametrize("compact", [False, True])
def test_host_warehouse_numpy_api(compact):
    """HostWarehouse (numpy in / numpy out over the wh_env_* layer) against BatchedWarehouse."""
    import torch
    from rllib_warehouse_b200 import SMALL, BatchedWarehouse, HostWarehouse
    n = 1001
    host = HostWarehouse(SMALL, n, seed=6, chunks=3, compact=compact)
    twin = BatchedWarehouse(SMALL, n, seed=6, auto_reset=True)
    host.reset(); twin.reset()
    rng = np.random.Generator(np.random.PCG64(1))
    for t in range(203):
        a = rng.integers(-1, 9, size=(n, 4))
        rew, dones, obs = host.step(a, with_obs=(not compact and t % 67 == 0))
        twin.step(a.astype(np.int32))
        assert np.array_equal(rew, twin.rewards.cpu().numpy()) and np.array_equal(dones, twin.dones.cpu().numpy().astype(bool))
        if obs is not None:
            for k, v in obs.items():
                assert np.array_equal(v, twin.obs[k].cpu().numpy()), k
    dev_obs = host.obs_tensors()
    for k, v in dev_obs.items():
        assert torch.equal(v, twin.obs[k]), k
    assert np.array_equal(host.stats(), twin.stats.cpu().numpy())
    host.close()


def test_vector_env_partial_send_actions_and_action_validation():
    """BaseEnv contract: only the envs named in send_actions advance (the others keep state, time and
    observations — env_mask of wh_step); actions are indices into MOVES (core.py:38,282): -9..-1 wrap like a
    Python list index, anything else raises IndexError; agent ids outside the env's agents raise too."""
    import torch
    from rllib_warehouse_b200 import MEDIUM, BatchedWarehouse, WarehouseVectorEnv
    n, A = 9, 4
    for flat in (False, True):
        venv = WarehouseVectorEnv(MEDIUM, n, num_agents=A, seed=5, flat_obs=flat)
        twin = BatchedWarehouse(MEDIUM, n, num_agents=A, seed=5)
        twin.reset()
        venv.poll()
        rng = np.random.Generator(np.random.PCG64(1))
        for t in range(25):
            chosen = sorted(rng.choice(n, size=int(rng.integers(1, n + 1)), replace=False).tolist())
            acts = {e: {str(i): int(rng.integers(-9, 9)) for i in range(A)} for e in chosen}
            venv.send_actions(acts)
            a = np.full((n, twin.R), -1, np.int32)
            mask = np.zeros(n, np.uint8)
            for e in chosen:
                mask[e] = 1
                for ag, v in acts[e].items():
                    a[e, int(ag)] = v % 9                                     # Python list index wrap (core.py:282)
            before = {k: v.clone() for k, v in twin.state.items()}
            twin.step(a, env_mask=mask)
            keep = torch.from_numpy(mask == 0).cuda()
            for k in before:                                                   # masked-out envs did not move in time
                assert torch.equal(twin.state[k][keep], before[k][keep]), k
            obs, rew, dones, _, _ = venv.poll()
            assert sorted(obs) == chosen
            for k in twin.state:
                assert torch.equal(venv.env.state[k], twin.state[k]), (t, k)
            for e in chosen:
                for i in range(A):
                    assert rew[e][str(i)] == twin.rewards[e, i].item()
                    if flat:
                        want = twin.build_obs_flat()[e, i].cpu().numpy()
                        assert np.array_equal(obs[e][str(i)], want), (t, e, i)
                    else:
                        assert all(np.array_equal(obs[e][str(i)][k], twin.obs[k][e, i].cpu().numpy()) for k in twin.obs)
        assert int(venv.env.state["time"].min()) < int(venv.env.state["time"].max())   # envs really desynchronised
        with pytest.raises(IndexError):
            venv.send_actions({0: {"0": 9}})
        with pytest.raises(IndexError):
            venv.send_actions({0: {"0": -10}})
        with pytest.raises(IndexError):
            venv.send_actions({0: {str(A): 0}})
        with pytest.raises(IndexError):
            venv.send_actions({n: {"0": 0}})
        assert venv.action_space.n == 9
        assert venv.observation_space.shape == (9 * 9 + 1,) if flat else venv.observation_space.contains(
            {k: twin.obs[k][0, 0].cpu().numpy() for k in twin.obs})


def test_rollout_sampler_through_the_base_env_adapter():
    """The ray-free sampler loop (RolloutSampler) drives WarehouseVectorEnv strictly through poll /
    send_actions / try_reset with a torch policy, across episode ends with per-env agent counts (*Train);
    the tensor mode (reset_tensors / step_tensors, in-kernel auto-reset) sees the same environment: both
    modes produce the same per-episode returns for the same policy and seed."""
    import torch
    from rllib_warehouse_b200 import SMALL, RolloutSampler, WarehouseVectorEnv, mlp_policy
    cfg = SMALL.replace(random_num_agents=True, episode_duration=15)
    n, iters = 48, 47                                            # 3 episodes and 2 steps
    policy = mlp_policy(9 * cfg.num_requests + 1, hidden=32, device="cuda:0", seed=3)
    a = RolloutSampler(WarehouseVectorEnv(cfg, n, seed=11, flat_obs=True), policy)
    ra = a.run_base_env(iters)
    b = RolloutSampler(WarehouseVectorEnv(cfg, n, seed=11, flat_obs=True, auto_reset=True), policy)
    rb = b.run_tensor(iters)
    assert ra["env_steps"] == rb["env_steps"] == n * iters
    assert len(a.episode_returns) == len(b.episode_returns) == 3 * n
    assert np.allclose(sorted(a.episode_returns), sorted(b.episode_returns))
    assert ra["agent_steps"] < n * cfg.num_requests * iters       # fewer than R agents in many envs (*Train)
    assert max(a.episode_returns) > 0
