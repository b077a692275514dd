#!/usr/bin/env python
"""wh_greedy_rollout (no observations) timing: `python tools/rollout_ab.py` -> us per step for Small / Medium batches."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rllib_warehouse_b200 import BatchedWarehouse, VARIANTS
for variant, n in (("small", 262144), ("small", 4096), ("medium", 65536), ("medium", 262144)):
    env = BatchedWarehouse(VARIANTS[variant], n, seed=1, auto_reset=True)
    env.reset()
    for _ in range(2):
        env.greedy_rollout(50, with_obs=False)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(5):
        env.greedy_rollout(50, with_obs=False)
    b.record(); torch.cuda.synchronize()
    print(json.dumps({"variant": variant, "envs": n, "us_per_step": round(a.elapsed_time(b) / 250 * 1e3, 3)}), flush=True)
