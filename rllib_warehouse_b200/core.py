"""`Warehouse` — the reference's RLlib MultiAgentEnv surface (warehouse/core.py:73-442) on top of
the batched CUDA environment.

Same constructor signature (core.py:78-86), attributes (core.py:111-118), `reset()` and
`step(action_dict)` contracts (per-agent observation / reward / done / info dicts keyed by
`str(i)`, `"__all__"` in dones), so `baseline/run.py`, `scripts/train.py` and `scripts/rollout.py`
work unchanged. One instance drives a 1-env `BatchedWarehouse`; this compatibility path pays one
device->host copy per step and is NOT the throughput path (that is `BatchedWarehouse` /
`WarehouseVectorEnv`, which keep everything on the GPU).
"""
import time
from typing import Dict, List, Tuple

import numpy as np
import torch

from . import spaces
from .batched import OBS_KEYS, Arena, BatchedWarehouse
from .config import WarehouseConfig

try:  # pragma: no cover - ray is absent in the build image
    from ray.rllib.env.multi_agent_env import MultiAgentEnv  # type: ignore
except Exception:  # noqa: BLE001
    class MultiAgentEnv:  # minimal stand-in with the same role (core.py:6,73)
        pass

__all__ = ["Warehouse"]

MOVES: List[List[int]] = [[x, y] for x in [-1, 0, 1] for y in [-1, 0, 1]]   # core.py:38
PICKUP_REWARD: float = 1.0     # core.py:40
DELIVERY_REWARD: float = 1.0   # core.py:41
ANIMATE_FRAMES_PER_STEP: int = 10      # core.py:69
ANIMATE_STEPS_PER_SECOND: float = 6.0  # core.py:70

_DEFAULT_DEVICE = "cuda:0"


class Warehouse(MultiAgentEnv):
    metadata = {"render.modes": ["human"]}   # core.py:74-76

    def __init__(self, num_agents: int, num_requests: int, area_dimension: int,
                 pickup_racks_arrangement: List[int], episode_duration: int,
                 pickup_wait_duration: int, *, device: str = None, seed: int = None,
                 random_num_agents: bool = False, max_num_agents: int = None) -> None:
        super().__init__()
        assert num_agents <= num_requests                                      # core.py:89
        self._config = WarehouseConfig(
            num_requests, area_dimension, tuple(pickup_racks_arrangement), episode_duration,
            pickup_wait_duration, max_num_agents or num_requests, random_num_agents)
        # The reference draws from the process-global np.random stream (core.py:196,215,339), so
        # `np.random.seed(s)` before construction makes it reproducible. Here the global stream
        # only supplies the 64-bit key of the device-side counter-based generator.
        if seed is None:
            seed = int(np.random.randint(0, 2**31 - 1)) | (int(np.random.randint(0, 2**31 - 1)) << 31)
        # One env, driven from the host through dicts: inputs and outputs live in page-locked host memory that
        # the kernels address directly — a step is one launch + one stream synchronisation, no copies.
        self._batched = BatchedWarehouse(self._config, 1, num_agents=num_agents,
                                         device=device or _DEFAULT_DEVICE, seed=seed, mapped_io=True)
        self._in = Arena([("actions", (1, num_requests), torch.int32), ("order", (1, num_requests), torch.int32)],
                         self._batched.device, mapped=True)
        self._num_agents = num_agents
        self._num_requests = num_requests
        # core.py:111-118
        self.num_agents: int = num_agents
        self.num_requests: int = num_requests
        self.animate_frames_per_step: int = ANIMATE_FRAMES_PER_STEP
        self.animate_steps_per_second: float = ANIMATE_STEPS_PER_SECOND
        self.reward_range = (0.0, 1.0)
        self.action_space = spaces.Discrete(len(MOVES))
        self.observation_space = spaces.observation_space(num_requests, area_dimension)
        self._viewer = None

    # ------------------------------------------------------------------------------------------
    def _obs_dicts(self, host=None) -> Dict[str, Dict[str, np.ndarray]]:
        # the kernel wrote the page-locked host buffer; the per-agent arrays are fresh copies, as in the
        # reference (callers may keep them across steps)
        host = self._batched.outputs_to_host() if host is None else host
        A = self.num_agents
        return {str(i): {k: host[k][0, i].copy() for k in OBS_KEYS} for i in range(A)}

    def reset(self) -> Dict[str, Dict[str, np.ndarray]]:
        """core.py:167-260 (and variants.py:69-71 for random agent counts)."""
        self._batched.reset()
        if self._config.random_num_agents:
            self.num_agents = self._num_agents = int(self._batched.state["num_agents"][0].item())
        return self._obs_dicts()

    def step(self, action_dict: Dict[str, int]) -> Tuple[
            Dict[str, Dict[str, np.ndarray]], Dict[str, float], Dict[str, bool], Dict[str, Dict]]:
        """core.py:262-442. Moves are resolved sequentially in `action_dict` iteration order
        (core.py:279); agents missing from the dict do not move."""
        R, A = self._num_requests, self.num_agents
        actions, order = self._in.host_views["actions"], self._in.host_views["order"]
        actions.fill(-1)
        order.fill(-1)
        ascending = True
        for t, (key, action) in enumerate(action_dict.items()):
            idx = int(key)
            action = int(action)
            if not 0 <= idx < A:
                raise IndexError(f"agent id {key!r} out of range")            # core.py:281
            MOVES[action]                                                      # core.py:282 IndexError
            actions[0, idx] = action % 9
            order[0, t] = idx
            ascending &= t == 0 or order[0, t - 1] < idx
        dev_in = self._in.to_device()
        self._batched.step(dev_in["actions"], order=None if ascending else dev_in["order"])
        host = self._batched.outputs_to_host()
        obs = self._obs_dicts(host)
        rew = host["rewards"][0].copy()
        done = bool(host["dones"][0])
        rewards = {str(i): rew[i] for i in range(A)}                            # core.py:435 (np.float32)
        dones = {str(i): done for i in range(A)}                                # core.py:438-440
        dones["__all__"] = done
        return obs, rewards, dones, {str(i): {} for i in range(A)}

    def render(self, mode: str = "human", animate: bool = False) -> None:
        """core.py:444-475. Copies env state (and the `_prev_*` mirrors, core.py:270-272) to the host and
        draws it like the reference: through gym's pyglet viewer when `gym.envs.classic_control.rendering`
        imports, else the frame's primitives are recorded (`self._viewer.frames`) and a text frame is
        printed: `.` floor, `p`/`P` idle/waiting pickup point, `d`/`D` idle/targeted delivery point,
        digits = free agents, letters a.. = delivering agents. With `animate`, 10 frames interpolated from
        the previous step's positions, paced to 6 steps per second (core.py:448-470)."""
        if mode != "human":
            raise NotImplementedError(f"render mode {mode!r}")                 # core.py:445-446
        from . import render as rd
        if self._viewer is None:
            self._viewer = rd.make_viewer(self._config.area_dimension)        # core.py:480-485
        if not self._batched.prev:
            # from now on every step keeps the previous state aside, as the reference always does; until the
            # next step the mirrors equal the current state (exactly the reference's situation after reset)
            self._batched.track_prev(True)
        cfg = self._config
        recorder = isinstance(self._viewer, rd.PrimitiveRecorder)
        if recorder:
            self._viewer.clear()
        rd.draw(self._viewer, cfg.area_dimension, cfg.pickup_racks_arrangement, self.render_state(), animate,
                self.animate_frames_per_step, self.animate_steps_per_second,
                sleep=(lambda s: None) if recorder else time.sleep)
        if recorder:
            print(self.render_text())

    def render_text(self) -> str:
        st = self.render_state()
        cfg, dim = self._config, self._config.area_dimension
        grid = [["." for _ in range(dim)] for _ in range(dim)]
        racks = cfg.pickup_racks_arrangement
        cells = [(x + ox, y + oy) for x in racks for y in racks for ox, oy in ((-1, -1), (0, -1), (-1, 0), (0, 0))]
        for p, (x, y) in enumerate(cells):                                     # core.py:171-175
            grid[y][x] = "P" if st["pickup_point_targets"][p] > -1 else "p"
        targeted = set(int(t) for t in st["agent_delivery_targets"] if t > -1)
        targeted |= set(int(t) for t in st["pickup_point_targets"] if t > -1)
        for d in range(cfg.num_delivery_points):                               # core.py:178-188
            v, side = 2 + d // 4, d % 4
            x, y = ((v, 0), (0, v), (v, dim - 1), (dim - 1, v))[side]
            grid[y][x] = "D" if d in targeted else "d"
        for i, ((x, y), t) in enumerate(zip(st["agent_positions"], st["agent_delivery_targets"])):
            grid[int(y)][int(x)] = chr(ord("a") + i) if t > -1 else str(i % 10)
        rows = ["".join(r) for r in reversed(grid)]                            # y = 0 at the bottom (core.py:14)
        return f"t={st['episode_time']}\n" + "\n".join(rows)

    def render_state(self) -> Dict[str, np.ndarray]:
        """Everything `render` reads (core.py:452-475), as the reference's int32 arrays: the current state
        and the `_prev_*` mirrors (equal to the current state until prev tracking has seen a step)."""
        st = self._batched.get_state()
        A = self.num_agents
        out = dict(agent_positions=st["agent_pos"][0, :A], agent_delivery_targets=st["agent_tgt"][0, :A],
                   pickup_point_targets=st["pickup_tgt"][0], pickup_point_timers=st["pickup_timer"][0],
                   episode_time=int(st["time"][0]))
        prev = self._batched.prev
        if prev:
            out.update(prev_agent_positions=prev["agent_pos"][0, :A].to(torch.int32).cpu().numpy(),
                       prev_agent_delivery_targets=prev["agent_tgt"][0, :A].to(torch.int32).cpu().numpy(),
                       prev_pickup_point_targets=prev["pickup_tgt"][0].to(torch.int32).cpu().numpy())
        else:
            out.update(prev_agent_positions=out["agent_positions"].copy(),
                       prev_agent_delivery_targets=out["agent_delivery_targets"].copy(),
                       prev_pickup_point_targets=out["pickup_point_targets"].copy())
        return out
