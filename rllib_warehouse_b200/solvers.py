"""Greedy baseline solver with the reference's interface (baseline/solvers.py:10-58), evaluated by
the batched warp-argmin kernel (`wh_greedy`)."""
from abc import ABC, abstractmethod
from typing import Dict

import ctypes as C

import numpy as np
import torch

from . import _native as nv
from .batched import OBS_KEYS, Arena, BatchedWarehouse
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
        R = num_requests
        # the four keys the solver reads (solvers.py:33-39) + the replayed eps-random branch, packed in one
        # page-locked buffer that wh_greedy reads directly; the actions come back the same way: one launch
        # and one stream synchronisation per call, no copies
        self._in = Arena([("self_position", (1, R, 2), torch.int32), ("self_availability", (1, R, 1), torch.int8),
                          ("self_delivery_target", (1, R, 2), torch.int32), ("requests", (1, R, R, 4), torch.int32),
                          ("is_random", (1, R), torch.uint8), ("random_actions", (1, R), torch.int32)],
                         self._env.device, mapped=True)
        self._ob = nv.Obs(**{k: (self._in.views[k].data_ptr() if k in self._in.views else None) for k in OBS_KEYS})
        self._out = Arena([("actions", (1, R), torch.int32)], self._env.device, mapped=True)

    def compute_action(self, observations: Dict[str, Dict[str, np.ndarray]]) -> Dict[str, np.ndarray]:
        A = self._num_agents
        h = self._in.host_views
        for i in range(A):
            o = observations[f"{i}"]
            h["self_position"][0, i] = o["self_position"]
            h["self_availability"][0, i] = o["self_availability"]
            h["self_delivery_target"][0, i] = o["self_delivery_target"]
            h["requests"][0, i] = o["requests"]
        # solvers.py:44-45: one uniform draw per agent per step (even when prob == 0), and the
        # random action comes from the action space's own sampler
        is_random, random_actions = h["is_random"], h["random_actions"]
        is_random.fill(0)
        random_actions.fill(-1)
        for i in range(A):
            if np.random.uniform() < self._random_action_prob:
                is_random[0, i] = 1
                random_actions[0, i] = int(self._action_space.sample())
        env = self._env
        d = self._in.to_device()
        with torch.cuda.device(env.device):
            rc = env.lib.wh_greedy(C.byref(env._cfg), C.byref(self._ob), env.state["num_agents"].data_ptr(),
                                   None, None, 1, 0, 0, 0, d["is_random"].data_ptr(),
                                   d["random_actions"].data_ptr(), self._out.views["actions"].data_ptr(),
                                   env._stream())
        nv.check(rc, "wh_greedy")
        env.launches += 1
        acts = self._out.to_host()["actions"]
        return {f"{i}": acts[0, i].copy() for i in range(A)}
