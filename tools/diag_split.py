#!/usr/bin/env python
"""Where does the fused step kernel's time go? Times, per variant: the fused step+obs kernel, the
step alone (no observation build), and the observation build alone (k_obs). CUDA events, 200 launches."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rllib_warehouse_b200 import BatchedWarehouse, VARIANTS

def timeit(fn, n=200, warm=20):
    for _ in range(warm): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n

for v in sys.argv[1:] or ["small", "medium", "large"]:
    N = 262144
    env = BatchedWarehouse(VARIANTS[v], N, device="cuda:0", seed=1, auto_reset=True)
    env.reset()
    acts = [torch.randint(0, 9, (N, env.R), dtype=torch.int32, device="cuda:0") for _ in range(16)]
    i = [0]
    def fused(): i[0] += 1; env.step(acts[i[0] % 16])
    def noobs(): i[0] += 1; env.step(acts[i[0] % 16], with_obs=False)
    def obs(): env.build_obs()
    def flat(): i[0] += 1; env.step_flat(acts[i[0] % 16])
    def gfused(): env.greedy_step(want_actions=False)
    R = env.R
    flat_bytes = N * (R * 4 * (9 * R + 1) + 4 * R + 4 * R + R + 1 + 2 * (3 * R + 3 * env.P + 5))
    r = dict(variant=v, fused_ms=timeit(fused), step_only_ms=timeit(noobs), obs_only_ms=timeit(obs),
             greedy_fused_ms=timeit(gfused), flat_ms=timeit(flat))
    r["flat_frac_of_6545.6"] = flat_bytes / (r["flat_ms"] * 1e-3) / 1e9 / 6545.6
    print(json.dumps(r))
