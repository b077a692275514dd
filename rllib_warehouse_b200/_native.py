"""ctypes binding of the C ABI in include/wh_b200.h (libwh_b200.so, hand-written sm_100a kernels).

There is deliberately NO fallback: if the shared library is missing or a call fails, we raise."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("WH_B200_LIB") or os.path.join(HERE, "lib", "libwh_b200.so")   # env override: kernel tuning builds
MAX_RACKS = 8
NUM_STATS = 80
OBS_STEP, OBS_RESET = 0, 1
FLAG_AUTO_RESET = 1
FLAG_COMPACT_IO = 2
FLAG_NO_PDL = 4
FLAG_PER_STEP_OUT = 8
# wh_multi_step kernel selection (WH_FLAG_MULTI_KERNEL(k), bits 4-6 of flags)
MULTI_KERNELS = {"auto": 0, "throughput": 1, "low_occupancy": 2, "ws1": 3, "ws2": 4}


def flag_multi_kernel(name):
    return (MULTI_KERNELS[name or "auto"] & 7) << 4

OBS_KEYS = (
    "num_agents", "self_position", "self_availability", "self_delivery_target",
    "other_positions", "other_availabilities", "other_delivery_targets", "requests",
)
STATE_KEYS = ("agent_pos", "agent_tgt", "pickup_tgt", "pickup_timer", "time", "num_agents", "episode", "acc")

# every symbol include/wh_b200.h declares
SYMBOLS = (
    "wh_version", "wh_error_string", "wh_num_pickup_points", "wh_num_delivery_points",
    "wh_reset", "wh_step", "wh_step_flat", "wh_build_obs", "wh_build_obs_flat", "wh_greedy", "wh_greedy_step", "wh_greedy_rollout", "wh_multi_step", "wh_save_prev", "wh_stats_allreduce",
    "wh_env_create", "wh_env_destroy", "wh_env_reset", "wh_env_step_host", "wh_env_step_host_compact", "wh_env_greedy_step_host",
    "wh_env_obs_ptrs", "wh_env_state_ptrs", "wh_env_stats_host", "wh_env_launch_count",
)


class NativeError(RuntimeError):
    pass


class Config(C.Structure):
    _fields_ = [
        ("num_requests", C.c_int32), ("area_dimension", C.c_int32), ("num_racks", C.c_int32),
        ("racks", C.c_int32 * MAX_RACKS), ("episode_duration", C.c_int32),
        ("pickup_wait_duration", C.c_int32), ("max_num_agents", C.c_int32),
        ("random_num_agents", C.c_int32),
    ]


class State(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in STATE_KEYS]


class Prev(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ("agent_pos", "agent_tgt", "pickup_tgt")]


class Obs(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in OBS_KEYS]


_lib = None


def lib():
    """Loads libwh_b200.so; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NativeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). rllib_warehouse_b200 has no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        L.wh_error_string.restype = C.c_char_p
        L.wh_env_launch_count.restype = C.c_int64
        vp, i64, u64, ci = C.c_void_p, C.c_int64, C.c_uint64, C.c_int
        L.wh_reset.argtypes = [vp, vp, i64, i64, u64, vp, vp, vp, vp, vp, vp, vp]
        L.wh_step.argtypes = [vp, vp, i64, i64, u64, vp, vp, vp, vp, vp, vp, vp, vp, ci, vp, vp]
        L.wh_step_flat.argtypes = [vp, vp, i64, i64, u64, vp, vp, vp, vp, vp, vp, ci, vp, vp]
        L.wh_greedy_step.argtypes = [vp, vp, i64, i64, u64, u64, u64, vp, vp, vp, vp, vp, ci, vp]
        L.wh_greedy_rollout.argtypes = [vp, vp, i64, i64, u64, u64, u64, ci, vp, vp, vp, ci, vp]
        L.wh_multi_step.argtypes = [vp, vp, i64, i64, u64, ci, vp, u64, u64, vp, vp, vp, vp, ci, vp]
        L.wh_build_obs.argtypes = [vp, vp, i64, ci, vp, vp]
        L.wh_build_obs_flat.argtypes = [vp, vp, i64, ci, vp, vp]
        L.wh_greedy.argtypes = [vp, vp, vp, vp, vp, i64, i64, u64, u64, vp, vp, vp, vp]
        L.wh_stats_allreduce.argtypes = [vp, vp, vp]
        L.wh_save_prev.argtypes = [vp, vp, vp, i64, vp]
        L.wh_env_create.argtypes = [vp, i64, ci, i64, u64, ci, vp]
        L.wh_env_destroy.argtypes = [vp]
        L.wh_env_destroy.restype = None
        L.wh_env_reset.argtypes = [vp]
        L.wh_env_step_host.argtypes = [vp, vp, vp, vp, vp]
        L.wh_env_step_host_compact.argtypes = [vp, vp, vp, vp]
        L.wh_env_greedy_step_host.argtypes = [vp, vp, vp]
        L.wh_env_obs_ptrs.argtypes = [vp, vp]
        L.wh_env_state_ptrs.argtypes = [vp, vp]
        L.wh_env_stats_host.argtypes = [vp, vp]
        L.wh_env_launch_count.argtypes = [vp]
        _lib = L
    return _lib


def check(rc, what):
    if rc != 0:
        raise NativeError(f"{what} failed: [{rc}] {lib().wh_error_string(rc).decode()}")


def make_config(cfg) -> Config:
    c = Config()
    c.num_requests, c.area_dimension = cfg.num_requests, cfg.area_dimension
    racks = cfg.pickup_racks_arrangement
    if len(racks) > MAX_RACKS:
        raise NativeError("at most 8 racks per axis are supported")
    c.num_racks = len(racks)
    for i, r in enumerate(racks):
        c.racks[i] = r
    c.episode_duration, c.pickup_wait_duration = cfg.episode_duration, cfg.pickup_wait_duration
    c.max_num_agents, c.random_num_agents = cfg.max_num_agents, int(cfg.random_num_agents)
    return c
