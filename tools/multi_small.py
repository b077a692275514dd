#!/usr/bin/env python
"""BASELINE configs[1] shape in one launch: wh_multi_step on 4 096 Small envs (greedy, 200 steps per launch,
observations written every step). `python tools/multi_small.py [envs] [steps]`; used under ncu to see what
bounds the launch-sized regime."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rllib_warehouse_b200 import BatchedWarehouse, VARIANTS
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
T = int(sys.argv[2]) if len(sys.argv) > 2 else 200
env = BatchedWarehouse(VARIANTS["small"], n, seed=1, auto_reset=True)
env.reset()
for _ in range(3):
    env.multi_step(T)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); a.record()
for _ in range(5):
    env.multi_step(T)
b.record(); torch.cuda.synchronize()
print(f"{n} small envs: {a.elapsed_time(b) / 5 / T * 1e3:.3f} us per step")
