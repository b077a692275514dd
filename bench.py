#!/usr/bin/env python
"""bench.py — agent-steps/s of the batched warehouse env.step on N B200s, beside the CPU reference.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [--steps K] [--warmup W]      # CPU arm (oracle port, all host threads)
    torchrun --nproc-per-node N ... bench.py --gpus N ...          # one rank per GPU

Workload (BASELINE.json configs[3]): WarehouseLarge, 16 agents, 262 144 envs per GPU, contiguous
global env-id shards, no data-path collective (weak scaling: per-GPU work is fixed as N grows;
`--scaling strong` splits a fixed total of 262 144 envs over the GPUs instead). A "step" is
one `env.step` of every env: the fused move/collision/expiry/pickup/respawn/delivery kernel with
the observation build, on int32 actions already resident in HBM (uniform-random; a pool of 200
action tensors = one full episode of distinct actions, cycled). Observations (2.2 GB per step in total) are larger than L2, so no flush
is needed between iterations. `e2e` is the same step through the host-buffer C ABI
(`wh_env_step_host`): actions come from pinned host memory every step and rewards + dones go
back to pinned host memory every step; observations stay in HBM for an on-device policy.
Prints ONE JSON line (rank 0).
"""
import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# NCCL writes its banner / debug lines to stdout by default; stdout must carry exactly one JSON line
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

VARIANT_AGENTS = {"small": 4, "medium": 9, "large": 16}
# compulsory I/O at API dtypes + narrow state read+write, per env-step (A = R)
ALG_BYTES_PER_ENV_STEP = {"small": 711, "medium": 3071, "large": 9147}
SOLVER_BYTES_PER_AGENT = {"small": 85, "medium": 165, "large": 277}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--variant", default="large", choices=list(VARIANT_AGENTS))
    ap.add_argument("--envs", type=int, default=262144, help="envs per GPU (weak, default) / total envs (strong)")
    ap.add_argument("--scaling", default="weak", choices=["strong", "weak"])
    ap.add_argument("--policy", default="random", choices=["random", "greedy", "greedy_fused"])
    ap.add_argument("--e2e-steps", type=int, default=100)
    ap.add_argument("--e2e-chunks", type=int, default=8)
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline leg")
    ap.add_argument("--cpu-envs", type=int, default=8192, help="sample size (envs) of the CPU legs")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary measurements (solver kernel, configs[2])")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--seed", type=int, default=20261018)
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port (plain C restatement of the reference, pthreads over envs)
# ------------------------------------------------------------------------------------------------
def cpu_rollout_rate(variant, n_envs, steps, threads, seed, policy="random", warmup=5):
    import numpy as np
    from oracle import wh_oracle as wo
    env = wo.OracleEnv(wo.variant_config(variant), n_envs, seed=seed)
    env.reset()
    rng = np.random.Generator(np.random.PCG64(seed))
    actions = rng.integers(0, 9, size=(16, n_envs, env.R)).astype(np.int32)   # 16 distinct action sets, cycled
    env.rollout(warmup, threads, policy=policy, actions=actions)
    t0 = time.perf_counter()
    agent_steps = env.rollout(steps, threads, policy=policy, actions=actions)
    dt = time.perf_counter() - t0
    return agent_steps / dt, dt, agent_steps


def _python_port_worker(job):
    """One reference-style env object (numpy port of core.py) stepped in a Python loop."""
    variant, seconds, seed = job
    import numpy as np
    from oracle import ref_port as rp
    from oracle import wh_oracle as wo
    v = wo.VARIANTS[variant]
    A = v["num_requests"]
    env = rp.PortWarehouse(A, v["num_requests"], v["area_dimension"], v["racks"], rng=np.random.default_rng(seed))
    env.reset()
    rng = np.random.default_rng(seed + 1)
    n, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        for _ in range(50):
            acts = rng.integers(0, 9, size=A)
            _, _, done = env.step(list(enumerate(acts.tolist())))
            n += 1
            if done:
                env.reset()
    return n * A, time.perf_counter() - t0


def python_port_baseline(variant, seconds, seed):
    """The reference's own cost class: Python + numpy, one env object per process (what RLlib's
    MultiAgentEnv->BaseEnv vectorisation loops over), on every host core."""
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    with mp.get_context("spawn").Pool(cores) as pool:
        res = pool.map(_python_port_worker, [(variant, seconds, seed + 17 * i) for i in range(cores)])
    rate = sum(n / dt for n, dt in res)
    return {"value": rate, "unit": "agent-steps/s", "cores": cores, "kind": "port",
            "sample": f"oracle/ref_port.py (numpy port at the reference's granularity), 1 {variant} env per process x "
                      f"{cores} processes, {seconds:.0f}s each, random actions"}


def cpu_baseline(args, budget_s):
    threads = os.cpu_count() or 1
    n = args.cpu_envs
    rate, dt, _ = cpu_rollout_rate(args.variant, n, 4, threads, args.seed, "random", warmup=1)
    steps = max(8, int(budget_s * rate / (n * VARIANT_AGENTS[args.variant])))
    steps = min(steps, 4000)
    rate, dt, agent_steps = cpu_rollout_rate(args.variant, n, steps, threads, args.seed, "random")
    out = {
        "value": rate, "unit": "agent-steps/s", "cores": threads, "kind": "port",
        "sample": f"oracle/wh_oracle.c (C port of core.py step+obs), {n} {args.variant} envs x {steps} steps, "
                  f"random actions, {threads} pthreads, {dt:.1f}s",
    }
    try:
        out["python_port"] = python_port_baseline(args.variant, 4.0, args.seed)
    except Exception as e:  # noqa: BLE001
        out["python_port"] = {"error": repr(e)}
    return out


def run_reference(args):
    """--impl reference: the CPU implementation of the same env.step on all host threads; each
    step is one env.step over a bounded sample of the workload (args.cpu_envs envs)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    from oracle import wh_oracle as wo
    threads = os.cpu_count() or 1
    n, A = args.cpu_envs, VARIANT_AGENTS[args.variant]
    env = wo.OracleEnv(wo.variant_config(args.variant), n, seed=args.seed)
    env.reset()
    rng = np.random.Generator(np.random.PCG64(args.seed))
    actions = rng.integers(0, 9, size=(16, n, env.R)).astype(np.int32)        # 16 distinct action sets, cycled
    env.rollout(args.warmup, threads, policy="random", actions=actions)
    t0 = time.perf_counter()
    agent_steps = env.rollout(args.steps, threads, policy="random", actions=actions)
    dt = time.perf_counter() - t0
    rate = agent_steps / dt
    sample = (f"oracle/wh_oracle.c (C port of the reference step+obs; the Python reference cannot travel to "
              f"the GPU box), bounded sample: {n} {args.variant} envs per step instead of the config's "
              f"envs_total, random actions, {threads} pthreads")
    emit({
        "impl": "reference", "metric": "agent_steps_per_sec", "value": rate, "unit": "agent-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": workload_config(args, max(1, args.gpus)),
        "cpu_baseline": {"value": rate, "unit": "agent-steps/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": "agent-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


def workload_config(args, world):
    per_gpu = args.envs if args.scaling == "weak" else args.envs // world
    return {
        "workload": f"warehouse-{args.variant}-{args.envs}-envs-{'per-gpu' if args.scaling == 'weak' else 'total'}",
        "variant": args.variant, "agents_per_env": VARIANT_AGENTS[args.variant],
        "envs_total": per_gpu * world, "envs_per_gpu": per_gpu, "policy": args.policy, "episode_steps": 200,
        "step": "one env.step of every env: fused move/collision/expiry/pickup/respawn/delivery + observation build",
        "l2": "observations written per step (2.2 GB per GPU for large) exceed the 126 MB L2; no flush needed",
        "parallelism": f"env-sharded x{world}, no per-step communication",
    }


# ------------------------------------------------------------------------------------------------
# clocks sampler (NVML) — runs during the timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
               0x4: "sw_power_cap", 0x80: "hw_power_brake", 0x2: "applications_clocks_setting"}

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz, self.ok = [], set(), None, False
        self._stop = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.01)

    def __enter__(self):
        if self.ok:
            self.t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self.ok:
            self.t.join()

    def summary(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "note": "NVML unavailable"}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from rllib_warehouse_b200 import VARIANTS, BatchedWarehouse
    from rllib_warehouse_b200 import _native as nv

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    A = VARIANT_AGENTS[args.variant]
    n_local = args.envs if args.scaling == "weak" else args.envs // world
    n_total = n_local * world
    cfg = VARIANTS[args.variant]
    env = BatchedWarehouse(cfg, n_local, device=dev, seed=args.seed, env_id0=rank * n_local, auto_reset=True)
    env.reset()
    R = env.R
    gen = torch.Generator(device=dev)
    gen.manual_seed(args.seed + rank)
    n_pool = 200 if args.policy == "random" else 1
    pool = [torch.randint(0, 9, (n_local, R), dtype=torch.int32, device=dev, generator=gen) for _ in range(n_pool)]

    launches_per_step = {"random": 1, "greedy": 2, "greedy_fused": 1}[args.policy]

    def one_step(i):
        if args.policy == "random":
            env.step(pool[i % n_pool])
        elif args.policy == "greedy":
            env.step(env.greedy_actions())
        else:
            env.greedy_step(want_actions=False)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        one_step(i)
    barrier()
    launches0 = env.launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        ev0.record()
        for i in range(args.steps):
            one_step(i)
        stats = env.stats.clone()
        if world > 1:
            dist.all_reduce(stats)           # end-of-rollout episode statistics over NCCL
        ev1.record()
        barrier()
    gpu_launches = env.launches - launches0
    ms = ev0.elapsed_time(ev1)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = n_total * A * args.steps / (ms * 1e-3)

    # ---- dominant kernel (fused step+obs) timed per launch with CUDA events on its stream ----
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(min(args.steps, 200))]
    for i, (a, b) in enumerate(evs):
        acts = pool[i % n_pool] if args.policy != "greedy" else env.greedy_actions()
        a.record()
        if args.policy == "greedy_fused":
            env.greedy_step(want_actions=False)
        else:
            env.step(acts)
        b.record()
    torch.cuda.synchronize()
    kms = sorted(a.elapsed_time(b) for a, b in evs)
    k_iso_ms = sum(kms) / len(kms)
    # Average launch duration of the dominant kernel. When the timed region holds exactly one k_step
    # launch per step and nothing else (random / greedy_fused policies), that is the CUDA-event time of
    # the region divided by its launches (an upper bound on the true kernel time: it still contains the
    # launch gaps). An event pair around every single launch (k_iso_ms) adds the event/launch overhead of a
    # non-back-to-back launch (~1 % for the 0.38 ms Large step, ~15 % for a 35 us Small step) and
    # defeats programmatic dependent launch; it is reported beside it. Two kernels per step (policy
    # "greedy": solver + step) can only be separated by per-launch events.
    k_avg_ms = ms / args.steps if launches_per_step == 1 else k_iso_ms
    peaks = {}
    peak_src = "fallback 6650 GB/s (B200_PROFILING.md)"
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        peak = float(peaks["hbm_gbs"])
        peak_src = "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
    except Exception:  # noqa: BLE001
        peak = 6650.0
    alg_bytes = ALG_BYTES_PER_ENV_STEP[args.variant] * n_local
    achieved = alg_bytes / (k_avg_ms * 1e-3) / 1e9
    traffic = None
    try:   # ncu-measured DRAM bytes per launch for this exact launch shape (profiles/, round 1)
        tj = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))[args.variant]
        if tj["envs_per_launch"] == n_local:
            traffic = tj["traffic_bytes"]
    except Exception:  # noqa: BLE001
        pass
    roofline = {
        "bound": "hbm", "kernel": "wh::k_step (fused step + observation build)", "achieved": achieved,
        "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
        "alg_bytes_per_launch": alg_bytes, "kernel_ms_avg": k_avg_ms,
        "kernel_ms_source": ("timed region (CUDA events) / launches in it" if launches_per_step == 1
                             else "one CUDA-event pair per launch"),
        "kernel_ms_isolated_avg": k_iso_ms, "kernel_ms_isolated_median": kms[len(kms) // 2],
        "frac_isolated": alg_bytes / (k_iso_ms * 1e-3) / 1e9 / peak,
        "peak_source": peak_src,
    }

    out = {
        "metric": "agent_steps_per_sec", "value": value, "unit": "agent-steps/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "int32",
        "data": "synthetic", "config": workload_config(args, world),
        "roofline": roofline, "gpu_launches": gpu_launches, "launches_per_step": launches_per_step,
        "clocks": clocks.summary(),
    }

    # ---- e2e through the host-buffer C ABI: pinned actions in, rewards + dones out, every step ----
    if not args.no_e2e:
        L = nv.lib()
        ccfg = nv.make_config(cfg)

        def run_e2e(compact, chunks):
            h = C.c_void_p()
            nv.check(L.wh_env_create(C.byref(ccfg), n_local, local, rank * n_local, args.seed, chunks,
                                     C.byref(h)), "wh_env_create")
            nv.check(L.wh_env_reset(h), "wh_env_reset")
            adt, rdt = (torch.int8, torch.uint8) if compact else (torch.int32, torch.float32)
            host_actions = [torch.randint(0, 9, (n_local, R), dtype=adt).pin_memory() for _ in range(4)]
            host_rewards = torch.zeros((n_local, R), dtype=rdt).pin_memory()
            host_dones = torch.zeros(n_local, dtype=torch.uint8).pin_memory()

            def call(i):
                if compact:
                    rc = L.wh_env_step_host_compact(h, host_actions[i & 3].data_ptr(), host_rewards.data_ptr(),
                                                    host_dones.data_ptr())
                else:
                    rc = L.wh_env_step_host(h, host_actions[i & 3].data_ptr(), host_rewards.data_ptr(),
                                            host_dones.data_ptr(), None)
                nv.check(rc, "wh_env_step_host")

            for i in range(5):
                call(i)
            barrier()
            t0 = time.perf_counter()
            for i in range(args.e2e_steps):
                call(i)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            tt = torch.tensor([dt], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt = float(tt.item())
            launches = int(L.wh_env_launch_count(h))
            res = {
                "value": n_total * A * args.e2e_steps / dt, "unit": "agent-steps/s",
                "h2d_bytes_per_step": n_local * R * host_actions[0].element_size(),
                "d2h_bytes_per_step": n_local * R * host_rewards.element_size() + n_local,
                "steps": args.e2e_steps, "ms_per_step": 1e3 * dt / args.e2e_steps, "chunks": chunks,
                "api": ("wh_env_step_host_compact (int8 actions / uint8 rewards on the wire)" if compact else
                        "wh_env_step_host (int32 actions / float32 rewards, the reference's dtypes)")
                       + "; C ABI, pinned host buffers, observations stay in HBM",
                "reward_checksum": float(host_rewards.sum()), "gpu_launches": launches,
            }
            L.wh_env_destroy(h)
            return res

        ref_dtypes = run_e2e(False, args.e2e_chunks)
        compact = run_e2e(True, args.e2e_chunks)
        best, other = (compact, ref_dtypes) if compact["value"] >= ref_dtypes["value"] else (ref_dtypes, compact)
        out["e2e"] = best
        out["e2e_alt"] = other
    if world == 1 and not args.no_extras:
        out["extras"] = extras(args, dev, peak)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(args, args.cpu_seconds)
    out["stats"] = {k: v for k, v in env.stats_dict(stats).items() if not k.startswith("avg_agent_reward_") or k.endswith("_all")}
    if rank == 0:
        emit(out)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _time_steps(fn, steps, warmup):
    import torch
    for i in range(warmup):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def extras(args, dev, peak):
    """Secondary lines (not the headline): the solver kernel's own roofline, the run.py-style loop
    (solver kernel + step kernel), the single fused kernel, and BASELINE configs[2]
    (Medium, 65 536 envs, batched greedy solver)."""
    import torch
    from rllib_warehouse_b200 import VARIANTS, BatchedWarehouse
    res = {}
    for name, variant, n in (("large_262144", "large", args.envs if args.variant == "large" else 262144),
                             ("configs2_medium_65536", "medium", 65536)):
        A = VARIANT_AGENTS[variant]
        env = BatchedWarehouse(VARIANTS[variant], n, device=dev, seed=args.seed + 1, auto_reset=True)
        env.reset()
        steps, warm = 100, 10
        ms_solver = _time_steps(lambda i: env.greedy_actions(), steps, warm)
        ms_loop = _time_steps(lambda i: env.step(env.greedy_actions()), steps, warm)
        ms_fused = _time_steps(lambda i: env.greedy_step(want_actions=False), steps, warm)
        rnd = torch.randint(0, 9, (n, env.R), dtype=torch.int32, device=dev)
        ms_flat = _time_steps(lambda i: env.step_flat(rnd), steps, warm)
        ms_roll = _time_steps(lambda i: env.greedy_rollout(50, with_obs=False), 6, 2) / 50
        sb = SOLVER_BYTES_PER_AGENT[variant] * n * A
        fb = ALG_BYTES_PER_ENV_STEP[variant] * n
        res[name] = {
            "solver_kernel": {"ms": ms_solver, "achieved_GBs": sb / ms_solver / 1e6, "frac": sb / ms_solver / 1e6 / peak,
                              "alg_bytes_per_launch": sb},
            "solver_plus_step": {"ms": ms_loop, "agent_steps_per_sec": n * A / (ms_loop * 1e-3), "launches_per_step": 2},
            "fused_greedy_step": {"ms": ms_fused, "agent_steps_per_sec": n * A / (ms_fused * 1e-3),
                                  "frac": fb / ms_fused / 1e6 / peak, "launches_per_step": 1},
            # baseline evaluation (run.py --envs N --rollout-kernel): 50 solver+step iterations per launch,
            # state in registers, NO observations written — a different (lighter) operation than env.step
            "greedy_rollout_kernel_no_obs": {"ms_per_step": ms_roll, "agent_steps_per_sec": n * A / (ms_roll * 1e-3),
                                             "steps_per_launch": 50},
            # RLlib-flattened float32 observations from the step kernel: 4(9R+1) B/agent instead of 33R+4
            "step_flat_f32_obs": {"ms": ms_flat, "agent_steps_per_sec": n * A / (ms_flat * 1e-3),
                                  "alg_bytes_per_launch": fb + n * A * (4 * (9 * A + 1) - (33 * A + 4)),
                                  "frac": (fb + n * A * (4 * (9 * A + 1) - (33 * A + 4))) / ms_flat / 1e6 / peak},
        }
        del env
        torch.cuda.empty_cache()
    # Proxy for BASELINE configs[4] (RLlib PPO rollouts; ray is not installed): sampling with an
    # on-device torch policy (2x256 MLP, bf16, argmax) fed by the step kernel's flattened observations.
    # The policy GEMMs are library code (cuBLAS through torch); the env side is this repo's kernels.
    try:
        n = 65536
        env = BatchedWarehouse(VARIANTS["large"], n, device=dev, seed=args.seed + 3, auto_reset=True)
        env.reset()
        F = 9 * env.R + 1
        torch.manual_seed(0)
        policy = torch.nn.Sequential(torch.nn.Linear(F, 256), torch.nn.ReLU(), torch.nn.Linear(256, 256),
                                     torch.nn.ReLU(), torch.nn.Linear(256, 9)).to(dev, torch.bfloat16)
        flat = [env.build_obs_flat(1)]

        def sample(i):
            with torch.no_grad():
                logits = policy(flat[0].view(-1, F).to(torch.bfloat16))
                actions = logits.argmax(dim=-1).view(n, env.R).to(torch.int32)
            flat[0], _, _ = env.step_flat(actions)

        ms_pol = _time_steps(sample, 50, 5)
        res["configs4_proxy_torch_policy_rollout"] = {
            "ms": ms_pol, "agent_steps_per_sec": n * 16 / (ms_pol * 1e-3), "envs": n,
            "note": "large, 65 536 envs; per step: wh_step_flat + bf16 MLP 145-256-256-9 over 1 048 576 agent rows + argmax",
        }
        del env, policy, flat
        torch.cuda.empty_cache()
    except Exception as e:  # noqa: BLE001
        res["configs4_proxy_torch_policy_rollout"] = {"error": repr(e)}
    # BASELINE configs[1] shape (Small, 4 096 envs): launch-bound eagerly, so also as a CUDA graph
    from rllib_warehouse_b200 import StepGraph
    env = BatchedWarehouse(VARIANTS["small"], 4096, device=dev, seed=args.seed + 2, auto_reset=True)
    env.reset()
    ms_eager = _time_steps(lambda i: env.greedy_step(want_actions=False), 400, 20)
    graph = StepGraph(env, steps=50, policy="greedy")
    ms_graph = _time_steps(lambda i: graph.replay(), 20, 3) / 50
    res["configs1_small_4096"] = {
        "fused_greedy_step_eager": {"ms": ms_eager, "agent_steps_per_sec": 4096 * 4 / (ms_eager * 1e-3)},
        "fused_greedy_step_cuda_graph": {"ms": ms_graph, "agent_steps_per_sec": 4096 * 4 / (ms_graph * 1e-3),
                                         "frac": ALG_BYTES_PER_ENV_STEP["small"] * 4096 / ms_graph / 1e6 / peak,
                                         "note": "50 steps per graph replay; the whole working set fits in L2"},
    }
    # BASELINE configs[0]: ONE Small env driven exactly like baseline/run.py:20-62 — the reference-named
    # dict API (WarehouseSmall(4), WarehouseRandomGreedySolver p=0), one 200-step episode. Per step: one
    # H2D + launch + D2H for the solver and the same for the env (packed pinned buffers), dict building
    # in Python. The reference's own single-core Python step + solver for this case: ~385 us.
    try:
        from rllib_warehouse_b200 import WarehouseRandomGreedySolver, WarehouseSmall
        wenv = WarehouseSmall(4)
        solver = WarehouseRandomGreedySolver(wenv.num_agents, wenv.num_requests, 0.0, wenv.action_space)
        think = step = 0.0
        for ep in range(3):                     # episode 0 = warm-up
            obs, done, n_steps = wenv.reset(), False, 0
            think = step = 0.0
            while not done:
                t0 = time.perf_counter()
                acts = solver.compute_action(obs)
                t1 = time.perf_counter()
                obs, rew, dones, _ = wenv.step(acts)
                t2 = time.perf_counter()
                think, step, n_steps, done = think + t1 - t0, step + t2 - t1, n_steps + 1, dones["__all__"]
        res["configs0_small_single_env_dict_api"] = {
            "steps": n_steps, "us_per_env_step": 1e6 * step / n_steps, "us_per_solver_call": 1e6 * think / n_steps,
            "agent_steps_per_sec": 4 * n_steps / (think + step),
            "note": "reference plumbing check, not a throughput path: 1 env, 4 agents, per-agent dicts on the host",
        }
    except Exception as e:  # noqa: BLE001
        res["configs0_small_single_env_dict_api"] = {"error": repr(e)}
    return res


_REAL_STDOUT = None


def _protect_stdout():
    """stdout must carry exactly ONE JSON line, but NCCL / torch print banners ("NCCL version ...")
    to fd 1 from C++. Point fd 1 at stderr for the whole run and keep the real stdout for emit()."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def emit(obj):
    line = (json.dumps(obj) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(line.decode())
        sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_REAL_STDOUT, line)


def main():
    _protect_stdout()
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
