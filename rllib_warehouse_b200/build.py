"""In-tree build of libwh_b200.so: explicit nvcc for sm_100a (no JIT cache, the .so travels with
the repo snapshot to the GPU box)."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIB_DIR, "libwh_b200.so")
SOURCES = ["wh_b200.cu"]
DEPS = ["wh_b200.cu", "wh_kernels.cuh", os.path.join("..", "..", "include", "wh_b200.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC", "-Xptxas", "-v", "-ldl",
]


def nvcc_path():
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def up_to_date():
    if not os.path.exists(LIB):
        return False
    t = os.path.getmtime(LIB)
    return all(os.path.getmtime(os.path.join(CSRC, d)) <= t for d in DEPS)


def build(force=False, verbose=False, extra_flags=(), out=None):
    """extra_flags / out: kernel-tuning builds (tools/ab_build.py) next to the default library."""
    if out is None and not force and up_to_date():
        return LIB
    os.makedirs(LIB_DIR, exist_ok=True)
    out = out or LIB
    cmd = [nvcc_path(), *NVCC_FLAGS, *extra_flags, "-o", out, *SOURCES]
    res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    log = res.stdout + res.stderr
    with open(os.path.join(LIB_DIR, "build.log") if out == LIB else out + ".log", "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + log[-4000:])
    if verbose:
        print(log)
    return out


if __name__ == "__main__":
    print(build(force=True, verbose=True))
