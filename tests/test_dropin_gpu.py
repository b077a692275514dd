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
    from rllib_warehouse_b200 import spaces
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
