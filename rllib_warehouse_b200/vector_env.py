"""WarehouseVectorEnv — the GPU environment as a vectorised multi-agent env for RLlib-style samplers.

RLlib (ray 0.8.x, the API generation the reference targets: `scripts/train.py:29-43`) wraps a
`MultiAgentEnv` into a `BaseEnv` whose `poll()` steps `num_envs_per_worker` Python env copies one
after the other. This adapter exposes the same `BaseEnv` protocol —

    poll() -> (obs, rewards, dones, infos, off_policy_actions)   dicts: env_id -> agent_id -> value
    send_actions({env_id: {agent_id: action}})
    try_reset(env_id) -> {agent_id: obs}
    get_unwrapped(), stop()

— but every `send_actions` is ONE kernel launch for all environments it names. Semantics kept from the
reference env that RLlib would otherwise wrap:
  * agent ids are `str(i)` (core.py:13); the per-env dict iteration order is the move-resolution order
    (core.py:279); agents missing from an env's dict do not move;
  * an action is an index into MOVES (core.py:38,282): -9..-1 wrap like a Python list index, anything
    else outside 0..8 raises IndexError;
  * only the environments present in `action_dict` advance (BaseEnv contract) — the others keep their
    state, time and observations (`env_mask` of wh_step).
With `flat_obs=True` observations are float32 vectors in RLlib's Dict-flattening order, produced by the
step kernel itself (`wh_step_flat`); `observation_space` is then the matching flat Box.

For callers that can consume tensors (a torch policy on the same GPU) `reset_tensors()` /
`step_tensors(actions)` skip the dict construction and all host copies.
"""
from typing import Dict, Optional, Tuple

import numpy as np
import torch

from . import _native as nv
from . import spaces
from .batched import OBS_KEYS, BatchedWarehouse
from .config import WarehouseConfig

try:  # pragma: no cover - ray is absent in the build image
    from ray.rllib.env.base_env import BaseEnv  # type: ignore
except Exception:  # noqa: BLE001
    class BaseEnv:  # protocol stand-in
        pass

__all__ = ["WarehouseVectorEnv"]

NUM_MOVES = 9   # len(MOVES), core.py:38


class WarehouseVectorEnv(BaseEnv):
    def __init__(self, config: WarehouseConfig, num_envs: int, num_agents: Optional[int] = None,
                 device: str = "cuda:0", seed: int = 0, flat_obs: bool = False, env_id0: int = 0,
                 auto_reset: bool = False):
        self.env = BatchedWarehouse(config, num_envs, num_agents=num_agents, device=device, seed=seed,
                                    env_id0=env_id0, auto_reset=auto_reset)
        self.config = config
        self.num_envs, self.R, self.flat_obs = int(num_envs), config.num_requests, bool(flat_obs)
        self.F = 9 * self.R + 1
        # what RLlib reads off the env when no explicit policy spaces are configured (core.py:117-148)
        self.action_space = spaces.Discrete(NUM_MOVES)
        self.observation_space = (spaces.Box(low=0, high=max(config.area_dimension, self.R), shape=(self.F,),
                                             dtype=np.float32)
                                  if self.flat_obs else spaces.observation_space(self.R, config.area_dimension))
        self._initialized = False
        self._pending: Optional[Tuple] = None
        self._ids = [str(i) for i in range(self.R)]
        self._num_agents = None                      # host copy of state["num_agents"], refreshed after resets
        N, R, dev = self.num_envs, self.R, self.env.device
        # host-side staging: actions + order + mask go up in one pinned buffer, results come down in pinned buffers
        self._in = torch.empty((2 * N * R + N,), dtype=torch.int32).pin_memory()
        self._in_dev = torch.empty((2 * N * R + N,), dtype=torch.int32, device=dev)
        self._actions = self._in[:N * R].view(N, R).numpy()
        self._order = self._in[N * R:2 * N * R].view(N, R).numpy()
        self._mask32 = self._in[2 * N * R:].numpy()
        self._mask_dev = torch.empty((N,), dtype=torch.uint8, device=dev)
        if self.flat_obs:
            self._flat_dev = torch.empty((N, R, self.F), dtype=torch.float32, device=dev)
            self._flat_host = torch.empty((N, R, self.F), dtype=torch.float32).pin_memory()
            self._rew_host = torch.empty((N, R), dtype=torch.float32).pin_memory()
            self._done_host = torch.empty((N,), dtype=torch.uint8).pin_memory()

    # ---- tensor API (zero-copy) ---------------------------------------------------------------
    def reset_tensors(self):
        obs = self.env.reset(with_obs=not self.flat_obs)
        self._num_agents = None
        return self.env.build_obs_flat(nv.OBS_RESET, out=self._flat_dev) if self.flat_obs else obs

    def step_tensors(self, actions: torch.Tensor):
        """actions [N,R] integer tensor (-1 = no action). Returns (obs, rewards[N,R], dones[N])."""
        if self.flat_obs:
            return self.env.step_flat(actions, out=self._flat_dev)       # one kernel: step + flattened observations
        return self.env.step(actions)

    # ---- BaseEnv protocol ----------------------------------------------------------------------
    def _agent_counts(self):
        if self._num_agents is None:
            self._num_agents = self.env.state["num_agents"].cpu().numpy().astype(np.int64)
        return self._num_agents

    def _fetch(self):
        """Device -> host for everything a poll() returns; arrays are fresh copies (the pinned buffers
        are reused by the next step, RLlib keeps observations until the batch is built)."""
        if self.flat_obs:
            self._flat_host.copy_(self._flat_dev, non_blocking=True)
            self._rew_host.copy_(self.env.rewards, non_blocking=True)
            self._done_host.copy_(self.env.dones, non_blocking=True)
            torch.cuda.current_stream(self.env.device).synchronize()
            return self._flat_host.numpy().copy(), self._rew_host.numpy().copy(), self._done_host.numpy().astype(bool)
        host = self.env.outputs_to_host()            # every key + rewards + dones: ONE device->host copy
        return {k: host[k].copy() for k in OBS_KEYS}, host["rewards"].copy(), host["dones"].astype(bool)

    def _obs_dicts(self, data, envs):
        A, ids = self._agent_counts(), self._ids
        if self.flat_obs:
            return {e: {ids[i]: data[e, i] for i in range(int(A[e]))} for e in envs}
        return {e: {ids[i]: {k: data[k][e, i] for k in OBS_KEYS} for i in range(int(A[e]))} for e in envs}

    def poll(self):
        if not self._initialized:
            self.reset_tensors()
            self._initialized = True
            data, _, _ = self._fetch()
            obs = self._obs_dicts(data, range(self.num_envs))
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
        N, A = self.num_envs, self._agent_counts()
        self._actions.fill(-1)
        self._order.fill(-1)
        self._mask32.fill(0)
        ascending = True
        for e, agent_actions in action_dict.items():
            if not 0 <= e < N:
                raise IndexError(f"env id {e!r} out of range")
            self._mask32[e] = 1
            prev = -1
            for t, (agent_id, action) in enumerate(agent_actions.items()):   # dict order is semantic (core.py:279)
                i, action = int(agent_id), int(action)
                if not 0 <= i < A[e]:
                    raise IndexError(f"agent id {agent_id!r} out of range for env {e}")            # core.py:281
                if not -NUM_MOVES <= action < NUM_MOVES:
                    raise IndexError(f"action {action} is not an index into MOVES")               # core.py:282
                self._actions[e, i] = action % NUM_MOVES
                self._order[e, t] = i
                ascending &= prev < i
                prev = i
        everyone = len(action_dict) == N
        dev_in = self._in_dev
        dev_in.copy_(self._in, non_blocking=True)                            # one H2D for actions + order + mask
        NR = N * self.R
        actions, order = dev_in[:NR].view(N, self.R), (None if ascending else dev_in[NR:2 * NR].view(N, self.R))
        mask = None
        if not everyone:
            self._mask_dev.copy_(dev_in[2 * NR:])
            mask = self._mask_dev
        if self.flat_obs:
            self.env.step_flat(actions, order=order, out=self._flat_dev, env_mask=mask)
        else:
            self.env.step(actions, order=order, env_mask=mask)
        data, rew, done = self._fetch()
        envs = list(action_dict.keys())
        obs = self._obs_dicts(data, envs)
        rewards = {e: {a: rew[e, int(a)] for a in obs[e]} for e in envs}
        dones = {e: {**{a: bool(done[e]) for a in obs[e]}, "__all__": bool(done[e])} for e in envs}
        infos = {e: {a: {} for a in obs[e]} for e in envs}
        self._pending = (obs, rewards, dones, infos)

    def try_reset(self, env_id: int):
        mask = np.zeros(self.num_envs, np.uint8)
        mask[env_id] = 1
        self.env.reset(env_mask=mask, with_obs=not self.flat_obs)
        self._num_agents = None                                              # *Train variants redraw the count
        if self.flat_obs:
            # the flat layout is rebuilt from the state for this env only (reset flavour); rows of the
            # other envs in the scratch tensor are not used
            scratch = self.env.build_obs_flat(nv.OBS_RESET)
            self._flat_dev[env_id].copy_(scratch[env_id])
            data = scratch[env_id].cpu().numpy()
            return {self._ids[i]: data[i] for i in range(int(self._agent_counts()[env_id]))}
        data, _, _ = self._fetch()
        return self._obs_dicts(data, [env_id])[env_id]

    def get_unwrapped(self):
        return [self.env]

    def stop(self) -> None:
        pass
