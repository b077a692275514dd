"""Measures this GPU's write-only (fill) and copy bandwidth with CUDA events — context for the
roofline of the store-dominated step+obs kernel (93 % of its algorithmic bytes are writes)."""
import json
import torch

dev = torch.device("cuda:0")
n = 2 * 1024**3  # bytes
a = torch.empty(n, dtype=torch.uint8, device=dev)
b = torch.empty(n, dtype=torch.uint8, device=dev)


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best

t_fill = timeit(lambda: a.fill_(7))
t_copy = timeit(lambda: b.copy_(a))
print(json.dumps({"fill_GBs": n / t_fill / 1e6, "copy_GBs_read_plus_write": 2 * n / t_copy / 1e6,
                  "bytes": n, "fill_ms": t_fill, "copy_ms": t_copy}))
