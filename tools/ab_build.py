#!/usr/bin/env python
"""Kernel-tuning builds: `python tools/ab_build.py name=-DFOO=1,-DBAR=2 ...` compiles one
libwh_b200 variant per argument into rllib_warehouse_b200/lib/ab/<name>.so (git-ignored, shipped to
the GPU box by gpurun); `WH_B200_LIB=<path>` selects one at run time (tools/ab_run.sh)."""
import os
import sys
from concurrent.futures import ThreadPoolExecutor

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rllib_warehouse_b200 import build as b  # noqa: E402

AB = os.path.join(b.LIB_DIR, "ab")


def one(arg):
    name, _, flags = arg.partition("=")
    out = os.path.join(AB, name + ".so")
    b.build(force=True, extra_flags=[f for f in flags.split(",") if f], out=out)
    regs = [l for l in open(out + ".log") if "k_step" in l or "registers" in l]
    return name, out, regs


if __name__ == "__main__":
    os.makedirs(AB, exist_ok=True)
    with ThreadPoolExecutor(8) as ex:
        for name, out, regs in ex.map(one, sys.argv[1:]):
            print(name, out)
