"""Timing of the UNMODIFIED reference (`oracle/_ref`, see make_ref.py) on the host cores —
MEASUREMENT INFRASTRUCTURE ONLY: only bench.py's `cpu_baseline` / `--impl reference` legs and tests/
may import this module.

Two legs, both driving the reference's own `Warehouse.step` (warehouse/core.py:262-442) through its
public API with per-agent action dicts, exactly as `baseline/run.py:42-62` and RLlib's sampler do:

  * single process: one Python loop over n env objects — this is what RLlib's own
    `MultiAgentEnv -> BaseEnv` vectorisation (`num_envs_per_worker`) does inside one rollout worker;
  * all cores: one process per host core (the `num_workers` axis of RLlib), each looping over its own
    env objects, started together behind a barrier; rate = all agent-steps / (last end - first start).

A "step" is one `env.step` of EVERY env object of the sample; episodes that end are reset (the cost
class RLlib pays too), actions are uniform-random and pre-generated outside the timed loop.
"""
import multiprocessing as mp
import os
import time

VARIANT_CLASS = {"small": "WarehouseSmall", "medium": "WarehouseMedium", "large": "WarehouseLarge"}
VARIANT_AGENTS = {"small": 4, "medium": 9, "large": 16}


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:  # pragma: no cover
        return os.cpu_count() or 1


def _make_envs(variant, n_envs, seed):
    import numpy as np
    from oracle import make_ref
    ref, _ = make_ref.import_reference()
    np.random.seed(seed)                      # the reference draws from the process-global stream
    A = VARIANT_AGENTS[variant]
    envs = [getattr(ref, VARIANT_CLASS[variant])(A) for _ in range(n_envs)]
    for e in envs:
        e.reset()
    rng = np.random.Generator(np.random.PCG64(seed + 1))
    acts = rng.integers(0, 9, size=(16, n_envs, A))
    action_dicts = [[{str(i): int(acts[s, e, i]) for i in range(A)} for e in range(n_envs)] for s in range(16)]
    return envs, action_dicts, A


def _run_steps(envs, action_dicts, steps, s0=0):
    for s in range(steps):
        row = action_dicts[(s0 + s) & 15]
        for k, env in enumerate(envs):
            _, _, dones, _ = env.step(row[k])
            if dones["__all__"]:
                env.reset()


def time_in_process(variant, n_envs, steps, warmup, seed=0):
    """One process, one loop over n_envs reference env objects. Returns (agent_steps, seconds)."""
    envs, action_dicts, A = _make_envs(variant, n_envs, seed)
    _run_steps(envs, action_dicts, warmup)
    t0 = time.perf_counter()
    _run_steps(envs, action_dicts, steps, warmup)
    dt = time.perf_counter() - t0
    return n_envs * A * steps, dt


def _worker(variant, n_envs, steps, warmup, seed, barrier, q):
    try:
        envs, action_dicts, A = _make_envs(variant, n_envs, seed)
        _run_steps(envs, action_dicts, warmup)
        barrier.wait(timeout=600)
        t0 = time.perf_counter()              # CLOCK_MONOTONIC: comparable across processes
        _run_steps(envs, action_dicts, steps, warmup)
        t1 = time.perf_counter()
        q.put((n_envs * A * steps, t0, t1, None))
    except Exception as e:  # noqa: BLE001
        q.put((0, 0.0, 0.0, repr(e)))


def time_all_cores(variant, envs_per_proc, steps, warmup, seed=0, procs=None):
    """`procs` processes (default: every host core), each stepping envs_per_proc reference env objects
    `steps` times after `warmup` steps. Returns (agent_steps, seconds, procs)."""
    procs = procs or host_cores()
    ctx = mp.get_context("spawn")
    barrier, q = ctx.Barrier(procs), ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(variant, envs_per_proc, steps, warmup, seed + 101 * i, barrier, q))
          for i in range(procs)]
    for p in ps:
        p.start()
    res, deadline = [], time.time() + 1800
    while len(res) < procs:
        try:
            res.append(q.get(timeout=1.0))
        except Exception:  # noqa: BLE001  (queue.Empty)
            dead = [p for p in ps if not p.is_alive() and p.exitcode not in (0, None)]
            if dead or time.time() > deadline:   # a worker died before reporting: do not wait for it
                for p in ps:
                    if p.is_alive():
                        p.terminate()
                raise RuntimeError(f"reference worker exited with {[p.exitcode for p in dead]} before reporting")
    for p in ps:
        p.join(timeout=60)
    errs = [r[3] for r in res if r[3]]
    if errs:
        raise RuntimeError("reference worker failed: " + errs[0])
    total = sum(r[0] for r in res)
    dt = max(r[2] for r in res) - min(r[1] for r in res)
    return total, dt, procs


def calibrate(variant, seed=0):
    """Seconds per env.step of ONE reference env object in this process (a few hundred steps)."""
    n, steps = 4, 60
    _, dt = time_in_process(variant, n, steps, 10, seed)
    return dt / (n * steps)
