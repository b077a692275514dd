#!/usr/bin/env python
"""wh_multi_step timing: `python tools/multi_bench.py variant envs steps_per_launch` (greedy, observations every step)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rllib_warehouse_b200 import BatchedWarehouse, VARIANTS
v, n, T = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
env = BatchedWarehouse(VARIANTS[v], n, seed=1, auto_reset=True)
env.reset()
for _ in range(2):
    env.multi_step(T)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); a.record()
for _ in range(4):
    env.multi_step(T)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 4 / T
bytes_ = {"small": 711, "medium": 3071, "large": 9147}[v] * n
print(f"{os.environ.get('WH_B200_LIB', 'default').split('/')[-1]} {v} {n} envs T={T}: {ms * 1e3:.2f} us/step, {bytes_ / ms / 1e6 / 6545.6:.4f} of the HBM peak (algorithmic bytes)")
