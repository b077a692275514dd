#!/usr/bin/env python
"""BASELINE configs[1] shape in one launch: wh_multi_step on 4 096 Small envs (200 steps per launch, observations
written every step), every wh_multi_step kernel side by side (WH_FLAG_MULTI_KERNEL).
`python tools/multi_small.py [variant] [envs] [steps] [kernel ...]`; with one kernel name it is the command line
for ncu (what bounds the launch-sized regime)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rllib_warehouse_b200 import BatchedWarehouse, VARIANTS

variant = sys.argv[1] if len(sys.argv) > 1 else "small"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
T = int(sys.argv[3]) if len(sys.argv) > 3 else 200
kernels = sys.argv[4:] or ["throughput", "low_occupancy", "ws1", "ws2", "auto"]
cfg = VARIANTS[variant]
R = cfg.num_requests
alg = {"small": 711, "medium": 3071, "large": 9147}[variant] * n            # algorithmic bytes per step


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps / T * 1e3                                  # us per step


for k in kernels:
    env = BatchedWarehouse(cfg, n, seed=1, auto_reset=True)
    env.reset()
    T_ps = max(2, min(T, int(16e9 // alg)))                                    # per-step slices stay below ~16 GB
    acts = torch.randint(0, 9, (T, n, R), dtype=torch.int32, device="cuda")
    res = {"variant": variant, "envs": n, "steps_per_launch": T, "steps_per_launch_per_step_slices": T_ps, "kernel": k}
    for name, kw, steps in (("greedy_resident", {}, T), ("greedy_per_step_slices", {"per_step": True}, T_ps),
                            ("open_loop_resident", {"actions": acts}, T),
                            ("open_loop_per_step_slices", {"actions": acts[:T_ps], "per_step": True}, T_ps)):
        if "per_step" in kw:
            kw["out"] = env.multi_step(steps, kernel=k, **kw)
        us = timed(lambda: env.multi_step(steps, kernel=k, **kw)) * T / steps
        res[name + "_us_per_step"] = round(us, 3)
        res[name + "_frac_of_hbm_peak"] = round(alg / (us * 1e-6) / 6545.6e9, 4)
        kw.pop("out", None)
        torch.cuda.empty_cache()
    print(json.dumps(res), flush=True)
