"""BatchedWarehouse — N independent warehouse environments resident in HBM.

Host-side mirror of `warehouse/core.py:73-442` for a whole batch: the same reset/step contract
(observations, rewards, dones), but tensors of shape [N, ...] instead of per-agent dicts, and
every bit of arithmetic done by the sm_100a kernels behind the C ABI (`include/wh_b200.h`).
PyTorch is used only to own device memory and streams.
"""
import ctypes as C

import numpy as np
import torch

from . import _native as nv
from .config import WarehouseConfig

OBS_KEYS = nv.OBS_KEYS


def _dev_tensor(x, dtype, device, shape=None):
    if x is None:
        return None
    if (isinstance(x, torch.Tensor) and x.dtype == dtype and x.is_contiguous()
            and (shape is None or tuple(x.shape) == tuple(shape))
            and (x.device == device or (x.device.type == "cpu" and x.is_pinned()))):
        # the per-step fast path: already what the kernel needs — a tensor on the device, or page-locked
        # host memory, which the kernels address directly (unified addressing; see Arena(mapped=True))
        return x
    if not isinstance(x, torch.Tensor):
        x = torch.from_numpy(np.ascontiguousarray(x))
    x = x.to(device=device, dtype=dtype).contiguous()
    if shape is not None:
        x = x.reshape(shape)
    return x


def _ptr(t):
    return None if t is None else t.data_ptr()


class Arena:
    """Several tensors as views into ONE device allocation (every view 256-byte aligned) with a pinned
    host twin of the same layout, so that a host-driven caller moves all of them with a single copy
    (`WarehouseVectorEnv`: one D2H per step instead of one per observation key).

    mapped=True (the one-env dict API of `core.Warehouse` / `solvers.WarehouseRandomGreedySolver`): there is
    no device copy at all — the views ARE the page-locked host buffer, and the kernels read / write it
    directly over PCIe (unified addressing: a cudaHostAlloc'ed pointer is valid on the device). For a
    handful of kilobytes per step that replaces two cudaMemcpyAsync round trips by nothing: a step is one
    launch + one stream synchronisation."""

    def __init__(self, specs, device, mapped=False):
        self.offsets, total = {}, 0
        self.mapped, self.device = bool(mapped), torch.device(device)
        for name, shape, dtype in specs:
            total = (total + 255) // 256 * 256
            nbytes = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
            self.offsets[name] = (total, nbytes, shape, dtype)
            total += nbytes
        self._host = None
        self.host_views = None
        if self.mapped:
            self.dev = torch.zeros(max(total, 256), dtype=torch.uint8).pin_memory()
        else:
            self.dev = torch.zeros(max(total, 256), dtype=torch.uint8, device=device)
        self.views = {k: self.dev[o:o + n].view(dt).view(sh) for k, (o, n, sh, dt) in self.offsets.items()}
        if self.mapped:
            self.host()

    def host(self):
        if self._host is None and self.mapped:
            self._host = self.dev
            self.host_views = {k: v.numpy() for k, v in self.views.items()}
        if self._host is None:
            self._host = torch.zeros(self.dev.shape, dtype=torch.uint8).pin_memory()
            self.host_views = {k: self._host[o:o + n].view(dt).view(sh).numpy()
                               for k, (o, n, sh, dt) in self.offsets.items()}
        return self._host

    def to_host(self):
        """device -> pinned host, one copy; returns numpy views (valid until the next call)."""
        if self.mapped:                       # the kernels wrote the host buffer themselves: wait for them
            torch.cuda.current_stream(self.device).synchronize()
            return self.host_views
        h = self.host()
        h.copy_(self.dev, non_blocking=True)
        torch.cuda.current_stream(self.dev.device).synchronize()
        return self.host_views

    def to_device(self):
        """pinned host (filled through `host_views`) -> device, one asynchronous copy."""
        if not self.mapped:
            self.dev.copy_(self.host(), non_blocking=True)
        return self.views


class _OnDevice:
    """`with torch.cuda.device(d)` costs several microseconds per launch; the current device already
    is `d` in the one-process-per-GPU setting, so only switch when it is not."""
    __slots__ = ("index", "ctx")

    def __init__(self, device):
        self.index, self.ctx = device.index if device.index is not None else torch.cuda.current_device(), None

    def __enter__(self):
        if torch.cuda.current_device() != self.index:
            self.ctx = torch.cuda.device(self.index)
            self.ctx.__enter__()
        return self

    def __exit__(self, *exc):
        if self.ctx is not None:
            ctx, self.ctx = self.ctx, None
            ctx.__exit__(*exc)
        return False


class BatchedWarehouse:
    """Structure-of-arrays state + observation tensors for `num_envs` environments on one GPU.

    num_agents: int (all envs) or None (= config.max_num_agents); per-env counts can be set through
    `reset(num_agents=...)` (replay) or drawn on device when `config.random_num_agents` (*Train).
    env_id0: global id of env 0 on this shard; the RNG is keyed by the GLOBAL env id, so results
    do not depend on how envs are split over GPUs.
    mapped_io: observations / rewards / dones live in page-locked HOST memory that the kernels write
    directly (for the one-env dict API: no device->host copy per step); `obs` etc. are then CPU tensors.
    """

    def __init__(self, config: WarehouseConfig, num_envs: int, num_agents=None, device="cuda:0",
                 seed: int = 0, env_id0: int = 0, auto_reset: bool = False, mapped_io: bool = False):
        if not torch.cuda.is_available():
            raise nv.NativeError("BatchedWarehouse needs a CUDA device: there is no CPU fallback")
        self.lib = nv.lib()
        self.config = config
        self.N = int(num_envs)
        self.R, self.P, self.D = config.num_requests, config.num_pickup_points, config.num_delivery_points
        self.device = torch.device(device)
        if self.device.index is None:          # "cuda" -> "cuda:<current>", so that device comparisons hold
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.seed, self.env_id0 = int(seed) & (2**64 - 1), int(env_id0)
        self.auto_reset = bool(auto_reset)
        self._on_device = _OnDevice(self.device)
        self._cfg = nv.make_config(config)
        N, R, P, dev = self.N, self.R, self.P, self.device
        A0 = config.max_num_agents if num_agents is None else int(num_agents)
        assert 1 <= A0 <= config.max_num_agents <= R                           # core.py:89, variants.py:24
        self.state = dict(
            agent_pos=torch.full((N, R, 2), -1, dtype=torch.int8, device=dev),
            agent_tgt=torch.full((N, R), -1, dtype=torch.int8, device=dev),
            pickup_tgt=torch.full((N, P), -1, dtype=torch.int8, device=dev),
            pickup_timer=torch.full((N, P), -1, dtype=torch.int16, device=dev),
            time=torch.zeros(N, dtype=torch.int32, device=dev),
            num_agents=torch.full((N,), A0, dtype=torch.int8, device=dev),
            episode=torch.full((N,), -1, dtype=torch.int32, device=dev),
            acc=torch.zeros((N, 4), dtype=torch.int32, device=dev),
        )
        i32, i8 = torch.int32, torch.int8
        # everything a step returns lives in one arena: `outputs_to_host()` is a single D2H copy
        self._out = Arena([
            ("num_agents", (N, R, 1), i32), ("self_position", (N, R, 2), i32),
            ("self_availability", (N, R, 1), i8), ("self_delivery_target", (N, R, 2), i32),
            ("other_positions", (N, R, R - 1, 2), i32), ("other_availabilities", (N, R, R - 1), i8),
            ("other_delivery_targets", (N, R, R - 1, 2), i32), ("requests", (N, R, R, 4), i32),
            ("rewards", (N, R), torch.float32), ("dones", (N,), torch.uint8)], dev, mapped=mapped_io)
        self.obs = {k: self._out.views[k] for k in OBS_KEYS}
        self.rewards = self._out.views["rewards"]
        self.dones = self._out.views["dones"]
        self.actions = torch.full((N, R), -1, dtype=torch.int32, device=dev)
        self.stats = torch.zeros(nv.NUM_STATS, dtype=torch.int64, device=dev)
        self._st = nv.State(**{k: self.state[k].data_ptr() for k in nv.STATE_KEYS})
        self._ob = nv.Obs(**{k: self.obs[k].data_ptr() for k in OBS_KEYS})
        self.launches = 0
        self.prev = None                         # render-only `_prev_*` mirrors (core.py:160-163), see track_prev()

    # ------------------------------------------------------------------------------------------
    def track_prev(self, on=True):
        """Keep the reference's render-only `_prev_*` state (core.py:270-272): every step first copies agent
        positions / delivery targets and pickup-point targets aside (three device-to-device copies, no
        synchronisation). Off by default — only `Warehouse.render(animate=True)` reads them."""
        if on and self.prev is None:
            self.prev = {k: self.state[k].clone() for k in ("agent_pos", "agent_tgt", "pickup_tgt")}
            self._pv = nv.Prev(**{k: v.data_ptr() for k, v in self.prev.items()})
        elif not on:
            self.prev = None

    def _save_prev(self):
        if self.prev is not None:
            with self._on_device:
                rc = self.lib.wh_save_prev(C.byref(self._cfg), C.byref(self._st), C.byref(self._pv), self.N, self._stream())
            nv.check(rc, "wh_save_prev")

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def outputs_to_host(self):
        """All observation keys + rewards + dones of the last reset/step as numpy views of ONE pinned
        host buffer, filled by a single device->host copy (meant for small N: the per-env dict API)."""
        return self._out.to_host()

    def obs_bytes_per_env(self):
        return sum(t[0].numel() * t.element_size() for t in self.obs.values())

    # ------------------------------------------------------------------------------------------
    def reset(self, agent_pos=None, init_pickups=None, init_targets=None, num_agents=None,
              env_mask=None, with_obs=True):
        """core.py:167-260. With `agent_pos` etc. the reference's recorded draws are replayed."""
        dev, N, R = self.device, self.N, self.R
        i8 = torch.int8
        agent_pos = _dev_tensor(agent_pos, i8, dev, (N, R, 2))
        init_pickups = _dev_tensor(init_pickups, i8, dev, (N, R))
        init_targets = _dev_tensor(init_targets, i8, dev, (N, R))
        num_agents = _dev_tensor(num_agents, i8, dev, (N,))
        env_mask = _dev_tensor(env_mask, torch.uint8, dev, (N,))
        keep = (agent_pos, init_pickups, init_targets, num_agents, env_mask)  # alive until launch returns
        with self._on_device:
            rc = self.lib.wh_reset(C.byref(self._cfg), C.byref(self._st), N, self.env_id0, self.seed,
                                   _ptr(agent_pos), _ptr(init_pickups), _ptr(init_targets),
                                   _ptr(num_agents), _ptr(env_mask),
                                   C.byref(self._ob) if with_obs else None, self._stream())
        nv.check(rc, "wh_reset")
        self.launches += 1
        del keep
        self._save_prev()                        # core.py:203-213: after a reset the mirrors equal the fresh state
        return self.obs

    def step(self, actions, order=None, spawn_pickups=None, spawn_targets=None, with_obs=True, env_mask=None):
        """core.py:262-442. actions [N,R] int (-1 = agent absent); order [N,R] = action-dict order.
        env_mask [N] (non-zero = step this env): the other envs keep state, observation, reward, done."""
        dev, N, R = self.device, self.N, self.R
        actions = _dev_tensor(actions, torch.int32, dev, (N, R))
        order = _dev_tensor(order, torch.int32, dev, (N, R))
        spawn_pickups = _dev_tensor(spawn_pickups, torch.int8, dev, (N, R))
        spawn_targets = _dev_tensor(spawn_targets, torch.int8, dev, (N, R))
        env_mask = _dev_tensor(env_mask, torch.uint8, dev, (N,))
        flags = nv.FLAG_AUTO_RESET if (self.auto_reset and spawn_pickups is None) else 0
        self._save_prev()                        # core.py:270-272
        with self._on_device:
            rc = self.lib.wh_step(C.byref(self._cfg), C.byref(self._st), N, self.env_id0, self.seed,
                                  _ptr(actions), _ptr(order), _ptr(spawn_pickups), _ptr(spawn_targets),
                                  self.rewards.data_ptr(), self.dones.data_ptr(), self.stats.data_ptr(),
                                  C.byref(self._ob) if with_obs else None, flags, _ptr(env_mask), self._stream())
        nv.check(rc, "wh_step")
        self.launches += 1
        return self.obs, self.rewards, self.dones

    def step_flat(self, actions, order=None, out=None, env_mask=None):
        """`step` whose observations are RLlib-flattened float32 [N, R, 9R+1], written by the step
        kernel itself (no dict-keyed tensors are produced)."""
        dev, N, R = self.device, self.N, self.R
        actions = _dev_tensor(actions, torch.int32, dev, (N, R))
        order = _dev_tensor(order, torch.int32, dev, (N, R))
        env_mask = _dev_tensor(env_mask, torch.uint8, dev, (N,))
        if out is None:
            if getattr(self, "_flat", None) is None:
                self._flat = torch.empty((N, R, 9 * R + 1), dtype=torch.float32, device=dev)
            out = self._flat
        flags = nv.FLAG_AUTO_RESET if self.auto_reset else 0
        self._save_prev()                        # core.py:270-272
        with self._on_device:
            rc = self.lib.wh_step_flat(C.byref(self._cfg), C.byref(self._st), N, self.env_id0, self.seed,
                                       _ptr(actions), _ptr(order), self.rewards.data_ptr(),
                                       self.dones.data_ptr(), self.stats.data_ptr(), out.data_ptr(),
                                       flags, _ptr(env_mask), self._stream())
        nv.check(rc, "wh_step_flat")
        self.launches += 1
        return out, self.rewards, self.dones

    def greedy_step(self, random_action_prob=0.0, solver_seed=0, with_obs=True, want_actions=True):
        """One run.py:42-62 loop iteration for all envs in a single kernel: greedy solver
        (solvers.py:27-58) evaluated from the resident state, then step + observation build."""
        thr = int(float(random_action_prob) * 4294967296.0)
        flags = nv.FLAG_AUTO_RESET if self.auto_reset else 0
        self._save_prev()                        # core.py:270-272
        with self._on_device:
            rc = self.lib.wh_greedy_step(C.byref(self._cfg), C.byref(self._st), self.N, self.env_id0,
                                         self.seed, int(solver_seed), thr,
                                         self.actions.data_ptr() if want_actions else None,
                                         self.rewards.data_ptr(), self.dones.data_ptr(),
                                         self.stats.data_ptr(), C.byref(self._ob) if with_obs else None,
                                         flags, self._stream())
        nv.check(rc, "wh_greedy_step")
        self.launches += 1
        return self.obs, self.rewards, self.dones

    def greedy_rollout(self, steps: int, random_action_prob=0.0, solver_seed=0, with_obs=True):
        """`steps` iterations of the run.py:42-62 loop (greedy solver -> step) in ONE kernel launch: the
        state stays in registers between the steps and no per-step observation is written. Leaves the
        state, `dones` and the statistics exactly as `steps` calls of `greedy_step` would; returns the
        per-agent reward SUMS of these steps [N, R] (also in `self.rewards`). with_obs: rebuild the
        resident observations from the final state afterwards (one more launch)."""
        thr = int(float(random_action_prob) * 4294967296.0)
        flags = nv.FLAG_AUTO_RESET if self.auto_reset else 0
        with self._on_device:
            rc = self.lib.wh_greedy_rollout(C.byref(self._cfg), C.byref(self._st), self.N, self.env_id0,
                                            self.seed, int(solver_seed), thr, int(steps),
                                            self.rewards.data_ptr(), self.dones.data_ptr(),
                                            self.stats.data_ptr(), flags, self._stream())
        nv.check(rc, "wh_greedy_rollout")
        self.launches += 1
        if with_obs:
            self.build_obs(nv.OBS_STEP)
        return self.rewards

    def multi_step(self, steps: int, actions=None, random_action_prob=0.0, solver_seed=0, with_obs=True,
                   per_step=False, out=None, kernel=None):
        """`steps` consecutive `env.step` calls in ONE kernel launch (wh_multi_step); the state stays in
        registers between them. actions: int [steps, N, R] open-loop actions, or None = the in-kernel greedy
        solver (one run.py:42-62 iteration per step). Every step's observations are written (with_obs).

        per_step=False: observations land in the resident tensors (each step overwrites the previous one,
        exactly what `steps` single launches leave behind); returns (obs, reward SUMS [N,R], last dones [N]).
        per_step=True: returns per-step tensors — obs dict of [steps, N, ...], rewards [steps, N, R],
        dones [steps, N] (allocated here, or pass `out=(obs_dict, rewards, dones)` to reuse buffers).
        kernel: None / "auto" = chosen from the launch size; "throughput", "low_occupancy", "ws1", "ws2" force one
        (WH_FLAG_MULTI_KERNEL; all produce identical results)."""
        dev, N, R, T = self.device, self.N, self.R, int(steps)
        if actions is not None:
            actions = _dev_tensor(actions, torch.int32, dev, (T, N, R))
        thr = int(float(random_action_prob) * 4294967296.0)
        flags = (nv.FLAG_AUTO_RESET if self.auto_reset else 0) | nv.flag_multi_kernel(kernel)
        if per_step:
            flags |= nv.FLAG_PER_STEP_OUT
            if out is None:
                obs = ({k: torch.empty((T,) + tuple(v.shape), dtype=v.dtype, device=dev) for k, v in self.obs.items()}
                       if with_obs else None)
                out = (obs, torch.empty((T, N, R), dtype=torch.float32, device=dev),
                       torch.empty((T, N), dtype=torch.uint8, device=dev))
            obs, rewards, dones = out
            ob = nv.Obs(**{k: obs[k].data_ptr() for k in OBS_KEYS}) if with_obs else None
        else:
            obs, rewards, dones = (self.obs if with_obs else None), self.rewards, self.dones
            ob = self._ob if with_obs else None
        with self._on_device:
            rc = self.lib.wh_multi_step(C.byref(self._cfg), C.byref(self._st), N, self.env_id0, self.seed, T,
                                        _ptr(actions), int(solver_seed), thr, rewards.data_ptr(), dones.data_ptr(),
                                        self.stats.data_ptr(), C.byref(ob) if ob is not None else None, flags,
                                        self._stream())
        nv.check(rc, "wh_multi_step")
        self.launches += 1
        return obs, rewards, dones

    def build_obs(self, flavour=nv.OBS_STEP):
        with self._on_device:
            rc = self.lib.wh_build_obs(C.byref(self._cfg), C.byref(self._st), self.N, int(flavour),
                                       C.byref(self._ob), self._stream())
        nv.check(rc, "wh_build_obs")
        self.launches += 1
        return self.obs

    def build_obs_flat(self, flavour=None, out=None):
        """Observations in RLlib's flattened float32 layout [N, R, 9R+1] (alphabetical key order of
        the core.py:119-148 Dict space), written by one kernel straight from the resident state.
        flavour: OBS_STEP (default) or OBS_RESET, as for build_obs."""
        N, R = self.N, self.R
        if out is None:
            if getattr(self, "_flat", None) is None:
                self._flat = torch.empty((N, R, 9 * R + 1), dtype=torch.float32, device=self.device)
            out = self._flat
        with self._on_device:
            rc = self.lib.wh_build_obs_flat(C.byref(self._cfg), C.byref(self._st), N,
                                            nv.OBS_STEP if flavour is None else int(flavour),
                                            out.data_ptr(), self._stream())
        nv.check(rc, "wh_build_obs_flat")
        self.launches += 1
        return out

    @staticmethod
    def flatten_obs(obs):
        """Reference flattening of the dict observations (RLlib Dict preprocessor order)."""
        keys = sorted(obs.keys())
        n, r = obs["requests"].shape[:2]
        return torch.cat([obs[k].reshape(n, r, -1).to(torch.float32) for k in keys], dim=2)

    def greedy_actions(self, obs=None, random_action_prob=0.0, solver_seed=0, is_random=None,
                       random_actions=None, out=None):
        """solvers.py:27-58 on observation tensors (defaults to the resident observations)."""
        dev, N, R = self.device, self.N, self.R
        if obs is None:
            ob = self._ob
            keep = None
        else:
            keep = {k: _dev_tensor(obs[k], self.obs[k].dtype, dev, self.obs[k].shape) for k in OBS_KEYS}
            ob = nv.Obs(**{k: keep[k].data_ptr() for k in OBS_KEYS})
        is_random = _dev_tensor(is_random, torch.uint8, dev, (N, R))
        random_actions = _dev_tensor(random_actions, torch.int32, dev, (N, R))
        out = self.actions if out is None else out
        thr = int(float(random_action_prob) * 4294967296.0)
        with self._on_device:
            rc = self.lib.wh_greedy(C.byref(self._cfg), C.byref(ob), self.state["num_agents"].data_ptr(),
                                    self.state["episode"].data_ptr(), self.state["time"].data_ptr(),
                                    N, self.env_id0, int(solver_seed), thr, _ptr(is_random),
                                    _ptr(random_actions), out.data_ptr(), self._stream())
        nv.check(rc, "wh_greedy")
        self.launches += 1
        del keep
        return out

    # ------------------------------------------------------------------------------------------
    def get_state(self):
        """State widened to the reference's int32 numpy arrays (core.py:153-165)."""
        torch.cuda.synchronize(self.device)
        return {k: v.to(torch.int32).cpu().numpy() for k, v in self.state.items()}

    def load_state(self, **arrays):
        """Inject state (reference widths accepted); used by tests and checkpoint restore."""
        for k, v in arrays.items():
            t = self.state[k]
            t.copy_(_dev_tensor(v, t.dtype, self.device, t.shape))

    def state_dict(self):
        return {k: v.clone() for k, v in self.state.items()}

    def load_state_dict(self, sd):
        self.load_state(**sd)

    def stats_dict(self, stats=None):
        """Episode statistics (scripts/train.py:18-23 custom metrics) from the int64 stats vector."""
        s = (self.stats if stats is None else stats).cpu().numpy().astype(np.int64)
        out = dict(episodes=int(s[0]), return_sum=int(s[1]), pickups=int(s[2]), deliveries=int(s[3]),
                   expired=int(s[4]))
        tot_avg_num = 0.0
        for n in range(1, self.R + 1):
            ep, ret = int(s[8 + 2 * (n - 1)]), int(s[9 + 2 * (n - 1)])
            if ep:
                out[f"avg_agent_reward_{n}"] = ret / n / ep
                tot_avg_num += ret / n
        if out["episodes"]:
            out["avg_agent_reward_all"] = tot_avg_num / out["episodes"]
        return out


class StepGraph:
    """CUDA-graph capture of `steps` consecutive environment steps for launch-bound batch sizes
    (e.g. BASELINE configs[1]: 4 096 Small envs, where one step kernel takes ~3 us but a Python
    launch ~10 us). policy="greedy": the fused solver+step kernel, no inputs. policy="actions":
    every captured step reads `self.actions` ([steps, N, R] int32), which the caller overwrites in
    place before `replay()`. Observations / rewards / dones land in the env's resident tensors
    (the values after the LAST captured step; per-step rewards are accumulated in
    `self.reward_sum` [N, R])."""

    def __init__(self, env: BatchedWarehouse, steps: int = 1, policy: str = "greedy", with_obs: bool = True,
                 random_action_prob: float = 0.0, solver_seed: int = 0):
        assert policy in ("greedy", "actions")
        self.env, self.steps, self.policy = env, int(steps), policy
        dev = env.device
        self.actions = torch.zeros((self.steps, env.N, env.R), dtype=torch.int32, device=dev)
        self.reward_sum = torch.zeros((env.N, env.R), dtype=torch.float32, device=dev)

        def body():
            self.reward_sum.zero_()
            for t in range(self.steps):
                if policy == "greedy":
                    env.greedy_step(random_action_prob, solver_seed, with_obs=with_obs, want_actions=False)
                else:
                    env.step(self.actions[t], with_obs=with_obs)
                self.reward_sum.add_(env.rewards)

        # capture never executes, so no warm-up run is needed (and it would advance the env)
        self.graph = torch.cuda.CUDAGraph()
        launches0 = env.launches
        with torch.cuda.device(dev):
            stream = torch.cuda.Stream(device=dev)
            stream.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.graph(self.graph, stream=stream):
                body()
        self.launches_per_replay = env.launches - launches0
        env.launches = launches0

    def replay(self):
        self.graph.replay()
        self.env.launches += self.launches_per_replay
        return self.env.obs, self.env.rewards, self.env.dones
