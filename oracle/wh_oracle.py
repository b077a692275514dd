"""ctypes front-end of the C oracle (oracle/wh_oracle.c) — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may
import this module; the product package (rllib_warehouse_b200) must never do so.
"""
import ctypes as C
import hashlib
import os
import re
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libwh_oracle.so")
MAX_RACKS = 8
NUM_STATS = 80

OBS_KEYS = [
    "num_agents", "self_position", "self_availability", "self_delivery_target",
    "other_positions", "other_availabilities", "other_delivery_targets", "requests",
]


class Config(C.Structure):
    _fields_ = [
        ("num_requests", C.c_int32), ("area_dimension", C.c_int32), ("num_racks", C.c_int32),
        ("racks", C.c_int32 * MAX_RACKS), ("episode_duration", C.c_int32),
        ("pickup_wait_duration", C.c_int32), ("max_num_agents", C.c_int32),
        ("random_num_agents", C.c_int32),
    ]


class State(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in (
        "agent_pos", "agent_tgt", "pickup_tgt", "pickup_timer", "time", "num_agents", "episode", "acc")]


class Obs(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in OBS_KEYS]


# variants.py:25-32, 40-47, 55-62
VARIANTS = {
    "small": dict(num_requests=4, area_dimension=12, racks=[4, 8], episode=200, wait=200),
    "medium": dict(num_requests=9, area_dimension=16, racks=[4, 8, 12], episode=200, wait=200),
    "large": dict(num_requests=16, area_dimension=20, racks=[4, 8, 12, 16], episode=200, wait=200),
}


def _host_tag():
    """Identifies the CPU the library was compiled for: it is built with -march=native, and the
    build container's .so travels to the GPU box, whose host may be a different CPU."""
    try:
        txt = open("/proc/cpuinfo").read()
        model = re.search(r"model name\s*:\s*(.*)", txt)
        flags = re.search(r"flags\s*:\s*(.*)", txt)
        return (model.group(1) if model else "?") + " | " + hashlib.sha1((flags.group(1) if flags else "").encode()).hexdigest()
    except OSError:
        return "unknown"


def build(force=False):
    tag_path, tag = LIB_PATH + ".host", _host_tag()
    stale = (not os.path.exists(LIB_PATH)
             or os.path.getmtime(LIB_PATH) < os.path.getmtime(os.path.join(HERE, "wh_oracle.c"))
             or not os.path.exists(tag_path) or open(tag_path).read() != tag)
    if force or stale:
        subprocess.check_call(["make", "-C", HERE, "-s", "-B"], env={**os.environ, "CC": "gcc"})
        with open(tag_path, "w") as f:
            f.write(tag)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB_PATH)
        _lib.who_rollout.restype = C.c_int64
    return _lib


def make_config(num_requests, area_dimension, racks, episode=200, wait=200, max_num_agents=None,
                random_num_agents=False):
    cfg = Config()
    cfg.num_requests, cfg.area_dimension, cfg.num_racks = num_requests, area_dimension, len(racks)
    for i, r in enumerate(racks):
        cfg.racks[i] = r
    cfg.episode_duration, cfg.pickup_wait_duration = episode, wait
    cfg.max_num_agents = max_num_agents or num_requests
    cfg.random_num_agents = int(random_num_agents)
    return cfg


def variant_config(size, random_num_agents=False):
    return make_config(random_num_agents=random_num_agents, **VARIANTS[size])


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _i32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.int32)


class OracleEnv:
    """N independent reference-semantics environments held in numpy (int32, reference widths)."""

    def __init__(self, cfg, n_envs, num_agents=None, seed=0, env_id0=0):
        self.cfg, self.N, self.seed, self.env_id0 = cfg, int(n_envs), int(seed), int(env_id0)
        R = self.R = cfg.num_requests
        self.P = lib().who_num_pickup_points(C.byref(cfg))
        self.D = lib().who_num_delivery_points(C.byref(cfg))
        N, P = self.N, self.P
        self.state = dict(
            agent_pos=np.full((N, R, 2), -1, np.int32), agent_tgt=np.full((N, R), -1, np.int32),
            pickup_tgt=np.full((N, P), -1, np.int32), pickup_timer=np.full((N, P), -1, np.int32),
            time=np.zeros(N, np.int32),
            num_agents=np.full(N, cfg.max_num_agents if num_agents is None else num_agents, np.int32),
            episode=np.full(N, -1, np.int32), acc=np.zeros((N, 4), np.int32),
        )
        self.obs = dict(
            num_agents=np.zeros((N, R, 1), np.int32), self_position=np.zeros((N, R, 2), np.int32),
            self_availability=np.zeros((N, R, 1), np.int8),
            self_delivery_target=np.zeros((N, R, 2), np.int32),
            other_positions=np.zeros((N, R, R - 1, 2), np.int32),
            other_availabilities=np.zeros((N, R, R - 1), np.int8),
            other_delivery_targets=np.zeros((N, R, R - 1, 2), np.int32),
            requests=np.zeros((N, R, R, 4), np.int32),
        )
        self.rewards = np.zeros((N, R), np.float32)
        self.dones = np.zeros(N, np.uint8)
        self.stats = np.zeros(NUM_STATS, np.int64)
        self.actions = np.full((N, R), -1, np.int32)

    def _st(self):
        s = State()
        for k, _ in State._fields_:
            setattr(s, k, self.state[k].ctypes.data)
        return s

    def _ob(self):
        o = Obs()
        for k in OBS_KEYS:
            setattr(o, k, self.obs[k].ctypes.data)
        return o

    def load_state(self, **arrays):
        for k, v in arrays.items():
            self.state[k][...] = np.asarray(v).reshape(self.state[k].shape)

    def reset(self, agent_pos=None, init_pickups=None, init_targets=None, num_agents=None, env_mask=None):
        agent_pos, init_pickups, init_targets = _i32(agent_pos), _i32(init_pickups), _i32(init_targets)
        num_agents = _i32(num_agents)
        env_mask = None if env_mask is None else np.ascontiguousarray(env_mask, np.uint8)
        st = self._st()
        rc = lib().who_reset(C.byref(self.cfg), C.byref(st), C.c_int64(self.N), C.c_int64(self.env_id0),
                             C.c_uint64(self.seed), _p(agent_pos), _p(init_pickups), _p(init_targets),
                             _p(num_agents), _p(env_mask))
        assert rc == 0, rc
        return self.build_obs(flavour=1)

    def step(self, actions, order=None, spawn_pickups=None, spawn_targets=None, with_obs=True):
        actions, order = _i32(actions), _i32(order)
        spawn_pickups, spawn_targets = _i32(spawn_pickups), _i32(spawn_targets)
        st = self._st()
        rc = lib().who_step(C.byref(self.cfg), C.byref(st), C.c_int64(self.N), C.c_int64(self.env_id0),
                            C.c_uint64(self.seed), _p(actions), _p(order), _p(spawn_pickups),
                            _p(spawn_targets), _p(self.rewards), _p(self.dones), _p(self.stats))
        assert rc == 0, rc
        if with_obs:
            self.build_obs(flavour=0)
        return self.obs, self.rewards, self.dones

    def build_obs(self, flavour):
        st, ob = self._st(), self._ob()
        rc = lib().who_build_obs(C.byref(self.cfg), C.byref(st), C.c_int64(self.N), C.c_int(flavour),
                                 C.byref(ob))
        assert rc == 0, rc
        return self.obs

    def greedy(self, obs=None, rand_prob=0.0, solver_seed=0, is_random=None, random_actions=None):
        if obs is not None:
            saved, self.obs = self.obs, obs
        ob = self._ob()
        if obs is not None:
            self.obs = saved
        thr = int(rand_prob * 4294967296.0)
        is_random = None if is_random is None else np.ascontiguousarray(is_random, np.uint8)
        random_actions = _i32(random_actions)
        rc = lib().who_greedy(C.byref(self.cfg), C.byref(ob), _p(self.state["num_agents"]),
                              _p(self.state["episode"]), _p(self.state["time"]), C.c_int64(self.N),
                              C.c_int64(self.env_id0), C.c_uint64(solver_seed), C.c_uint64(thr),
                              _p(is_random), _p(random_actions), _p(self.actions))
        assert rc == 0, rc
        return self.actions

    def rollout(self, n_steps, n_threads, policy="greedy", actions=None, auto_reset=True):
        st, ob = self._st(), self._ob()
        actions = _i32(actions)
        n_sets = 1 if actions is None or actions.ndim < 3 else actions.shape[0]
        n = lib().who_rollout(C.byref(self.cfg), C.byref(st), C.byref(ob), C.c_int64(self.N),
                              C.c_int64(self.env_id0), C.c_uint64(self.seed),
                              C.c_int(1 if policy == "greedy" else 0), _p(actions), _p(self.actions),
                              _p(self.rewards), _p(self.dones), _p(self.stats), C.c_int(n_steps),
                              C.c_int(n_threads), C.c_int(int(auto_reset)), C.c_int(n_sets))
        assert n >= 0, n
        return int(n)


def philox(c0, c1, c2, c3, seed):
    out = (C.c_uint32 * 4)()
    lib().who_philox(C.c_uint32(c0), C.c_uint32(c1), C.c_uint32(c2), C.c_uint32(c3), C.c_uint64(seed), out)
    return list(out)
