"""WarehouseVectorEnv — the GPU environment as a vectorised multi-agent env for RLlib-style samplers.

RLlib (ray 0.8.x, the API generation the reference targets: `scripts/train.py:29-43`) wraps a
`MultiAgentEnv` into a `BaseEnv` whose `poll()` steps `num_envs_per_worker` Python env copies one
after the other. This adapter exposes the same `BaseEnv` protocol —

    poll() -> (obs, rewards, dones, infos, off_policy_actions)   dicts: env_id -> agent_id -> value
    send_actions({env_id: {agent_id: action}})
    try_reset(env_id) -> {agent_id: obs}
    get_unwrapped(), stop()

— but every `send_actions` is ONE kernel launch for all environments. Agent ids are `str(i)` as in
the reference (core.py:13). With `flat_obs=True` observations are float32 vectors in RLlib's
Dict-flattening order, produced directly by `wh_build_obs_flat`.

For callers that can consume tensors (a torch policy on the same GPU) `reset_tensors()` /
`step_tensors(actions)` skip the dict construction and all host copies.
"""
from typing import Dict, Optional, Tuple

import numpy as np
import torch

from . import _native as nv
from .batched import OBS_KEYS, BatchedWarehouse
from .config import WarehouseConfig

try:  # pragma: no cover - ray is absent in the build image
    from ray.rllib.env.base_env import BaseEnv  # type: ignore
except Exception:  # noqa: BLE001
    class BaseEnv:  # protocol stand-in
        pass

__all__ = ["WarehouseVectorEnv"]


class WarehouseVectorEnv(BaseEnv):
    def __init__(self, config: WarehouseConfig, num_envs: int, num_agents: Optional[int] = None,
                 device: str = "cuda:0", seed: int = 0, flat_obs: bool = False, env_id0: int = 0):
        self.env = BatchedWarehouse(config, num_envs, num_agents=num_agents, device=device, seed=seed,
                                    env_id0=env_id0, auto_reset=False)
        self.num_envs, self.R, self.flat_obs = int(num_envs), config.num_requests, bool(flat_obs)
        self._initialized = False
        self._pending: Optional[Tuple] = None
        self._actions = np.full((self.num_envs, self.R), -1, np.int32)
        self._order = np.full((self.num_envs, self.R), -1, np.int32)

    # ---- tensor API (zero-copy) ---------------------------------------------------------------
    def reset_tensors(self):
        obs = self.env.reset()
        return self.env.build_obs_flat(nv.OBS_RESET) if self.flat_obs else obs

    def step_tensors(self, actions: torch.Tensor):
        """actions [N,R] integer tensor (-1 = no action). Returns (obs, rewards[N,R], dones[N])."""
        if self.flat_obs:
            return self.env.step_flat(actions)       # one kernel: step + flattened observations
        return self.env.step(actions)

    # ---- BaseEnv protocol ----------------------------------------------------------------------
    def _host_obs(self, flavour, envs=None):
        """Per-env, per-agent observation dicts (or flat vectors) on the host."""
        A = self.env.state["num_agents"].cpu().numpy()
        host = self.env.outputs_to_host()           # every key + rewards + dones: ONE device->host copy
        if self.flat_obs:
            flat = self.env.build_obs_flat(flavour).cpu().numpy()
            get = lambda e, i: flat[e, i]
        else:                                       # fresh arrays: the pinned buffer is reused next step
            get = lambda e, i: {k: host[k][e, i].copy() for k in OBS_KEYS}
        envs = range(self.num_envs) if envs is None else envs
        return {e: {str(i): get(e, i) for i in range(int(A[e]))} for e in envs}, host

    def poll(self):
        if not self._initialized:
            self.env.reset()
            self._initialized = True
            obs, _ = self._host_obs(nv.OBS_RESET)
            rewards = {e: {a: None for a in obs[e]} for e in obs}
            dones = {e: {**{a: False for a in obs[e]}, "__all__": False} for e in obs}
            infos = {e: {a: {} for a in obs[e]} for e in obs}
            return obs, rewards, dones, infos, {}
        if self._pending is None:
            return {}, {}, {}, {}, {}
        obs, rewards, dones, infos = self._pending
        self._pending = None
        return obs, rewards, dones, infos, {}

    def send_actions(self, action_dict: Dict[int, Dict[str, int]]) -> None:
        self._actions.fill(-1)
        self._order.fill(-1)
        ascending = True
        for e, agent_actions in action_dict.items():
            for t, (agent_id, action) in enumerate(agent_actions.items()):   # dict order is semantic (core.py:279)
                i = int(agent_id)
                self._actions[e, i] = int(action)
                self._order[e, t] = i
                ascending &= t == 0 or self._order[e, t - 1] < i
        self.env.step(self._actions, order=None if ascending else self._order, with_obs=not self.flat_obs)
        obs, host = self._host_obs(nv.OBS_STEP, envs=list(action_dict.keys()))
        rew = host["rewards"].copy()
        done = host["dones"].astype(bool)
        rewards = {e: {a: rew[e, int(a)] for a in obs[e]} for e in obs}
        dones = {e: {**{a: bool(done[e]) for a in obs[e]}, "__all__": bool(done[e])} for e in obs}
        infos = {e: {a: {} for a in obs[e]} for e in obs}
        self._pending = (obs, rewards, dones, infos)

    def try_reset(self, env_id: int):
        mask = np.zeros(self.num_envs, np.uint8)
        mask[env_id] = 1
        self.env.reset(env_mask=mask)
        obs, _ = self._host_obs(nv.OBS_RESET, envs=[env_id])
        return obs[env_id]

    def get_unwrapped(self):
        return [self.env]

    def stop(self) -> None:
        pass
