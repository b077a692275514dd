"""Greedy baseline solver with the reference's interface (baseline/solvers.py:10-58), evaluated by
the batched warp-argmin kernel (`wh_greedy`)."""
from abc import ABC, abstractmethod
from typing import Dict

import numpy as np
import torch

from . import _native as nv
from .batched import OBS_KEYS, BatchedWarehouse
from .config import WarehouseConfig

__all__ = ["WarehouseSolver", "WarehouseRandomGreedySolver", "BatchedGreedySolver"]


class WarehouseSolver(ABC):               # solvers.py:10-15
    @abstractmethod
    def compute_action(self, observations: Dict[str, Dict[str, np.ndarray]]) -> Dict[str, np.ndarray]:
        ...


class BatchedGreedySolver:
    """solvers.py:27-58 for [N,R] agent rows at once. Works on any observation tensors with the
    layout of `BatchedWarehouse.obs` (only self_position, self_availability, self_delivery_target
    and requests are read, as in the reference)."""

    def __init__(self, env: BatchedWarehouse, random_action_prob: float = 0.0, seed: int = 0):
        self.env, self.p, self.seed = env, float(random_action_prob), int(seed)

    def compute_actions(self, obs=None, is_random=None, random_actions=None, out=None) -> torch.Tensor:
        return self.env.greedy_actions(obs, self.p, self.seed, is_random, random_actions, out)


class WarehouseRandomGreedySolver(WarehouseSolver):
    """Drop-in for baseline/solvers.py:18-58 (same constructor and compute_action contract)."""

    def __init__(self, num_agents: int, num_requests: int, random_action_prob: float, action_space,
                 *, device: str = "cuda:0") -> None:
        self._num_agents = num_agents
        self._num_requests = num_requests
        self._random_action_prob = random_action_prob
        self._action_space = action_space
        # geometry is irrelevant to the solver; any config with R = num_requests serves
        L = 1
        while 4 * L * L < num_requests:
            L += 1
        dim = max(20, num_requests // 4 + 5)
        dim = min(dim, 20)
        cfg = WarehouseConfig(num_requests, dim, tuple(4 * (i + 1) for i in range(L)), 200, 200, num_requests)
        self._env = BatchedWarehouse(cfg, 1, num_agents=num_agents, device=device)
        self._obs = {k: np.zeros(tuple(self._env.obs[k].shape), dtype=np.int32) for k in OBS_KEYS}

    def compute_action(self, observations: Dict[str, Dict[str, np.ndarray]]) -> Dict[str, np.ndarray]:
        A, R = self._num_agents, self._num_requests
        for i in range(A):
            o = observations[f"{i}"]
            self._obs["self_position"][0, i] = o["self_position"]
            self._obs["self_availability"][0, i] = o["self_availability"]
            self._obs["self_delivery_target"][0, i] = o["self_delivery_target"]
            self._obs["requests"][0, i] = o["requests"]
        # solvers.py:44-45: one uniform draw per agent per step (even when prob == 0), and the
        # random action comes from the action space's own sampler
        is_random = np.zeros((1, R), np.uint8)
        random_actions = np.full((1, R), -1, np.int32)
        for i in range(A):
            if np.random.uniform() < self._random_action_prob:
                is_random[0, i] = 1
                random_actions[0, i] = int(self._action_space.sample())
        acts = self._env.greedy_actions(self._obs, 0.0, 0, is_random, random_actions).cpu().numpy()
        return {f"{i}": acts[0, i] for i in range(A)}
