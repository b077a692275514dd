#!/usr/bin/env python
"""oracle/_ref recipe — TEST / MEASUREMENT INFRASTRUCTURE ONLY (never imported by the product path).

The reference (ffahleraz/rllib-warehouse) is pure Python, so there is nothing to compile: the
"build" of `oracle/_ref/` is a verbatim file copy of the reference's hot-path modules from where
they lie under `/root/reference` —

    warehouse/{__init__,core,variants}.py   (core.py:73-442 Warehouse, variants.py:19-98)
    baseline/{solvers,run}.py               (solvers.py:18-58, run.py:15-99)

— into the git-ignored directory `oracle/_ref/` (listed in `.gitignore`, NOT in `.gpurunignore`,
so it never enters the history but travels to the GPU box like the built `.so` files), plus a
MANIFEST.json with the SHA-256 of every copied file so that a run can prove the files are the
unmodified reference. `gym` / `ray` are absent from this image; the stand-ins the copy imports
under are `oracle/stubs/` (tracked; ~100 lines, only the names core.py:5-6,118-148 touch).

Used by: `bench.py --impl reference` / the `cpu_baseline` leg (kind "reference": the reference's
own `Warehouse.step` timed on the box's host cores), and `tests/test_zz_reference_files.py`
(the reference's own `baseline/run.py` file executed unmodified against this repo's `warehouse` /
`solvers` shims on the GPU). `/root/reference` itself is never read on the GPU box.

    python oracle/make_ref.py        # (re)creates oracle/_ref/ ; no-op message if the reference is not mounted
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("WH_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")
STUBS = os.path.join(HERE, "stubs")
FILES = (
    "warehouse/__init__.py", "warehouse/core.py", "warehouse/variants.py", "warehouse/py.typed",
    "baseline/solvers.py", "baseline/run.py",
)


def _sha(path):
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def available():
    return os.path.isfile(os.path.join(OUT, "MANIFEST.json"))


def make(force=False):
    """Copies the reference's hot-path files into oracle/_ref/. Returns the manifest, or None when
    the reference is not mounted (the GPU box: there the copy made in the build container is used)."""
    if not os.path.isdir(os.path.join(REF, "warehouse")):
        return json.load(open(os.path.join(OUT, "MANIFEST.json"))) if available() else None
    manifest = {"source": REF, "files": {}}
    for rel in FILES:
        src, dst = os.path.join(REF, rel), os.path.join(OUT, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if force or not os.path.exists(dst) or _sha(src) != _sha(dst):
            shutil.copyfile(src, dst)
        manifest["files"][rel] = _sha(dst)
    with open(os.path.join(OUT, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    return manifest


def verify():
    """True iff every file under oracle/_ref still has the SHA-256 recorded when it was copied."""
    if not available():
        return False
    m = json.load(open(os.path.join(OUT, "MANIFEST.json")))
    return all(os.path.isfile(os.path.join(OUT, rel)) and _sha(os.path.join(OUT, rel)) == h
               for rel, h in m["files"].items())


def import_reference():
    """Imports the copied reference (`warehouse` package + `solvers` module) under the stub gym / ray,
    without leaving either in sys.modules / sys.path: this repo ships its own `warehouse` and
    `solvers` shims under the same names. Returns (warehouse_module, WarehouseRandomGreedySolver)."""
    if not available():
        raise FileNotFoundError("oracle/_ref is missing: run `python oracle/make_ref.py` where /root/reference is mounted")
    names = ("warehouse", "solvers", "gym", "ray")
    mine = lambda k: k in names or k.startswith(tuple(n + "." for n in names))   # noqa: E731
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if mine(k)}
    paths = [STUBS, OUT, os.path.join(OUT, "baseline")]
    sys.path[:0] = paths
    try:
        import warehouse as ref_pkg
        from solvers import WarehouseRandomGreedySolver as ref_solver
        assert os.path.realpath(ref_pkg.__file__).startswith(os.path.realpath(OUT)), ref_pkg.__file__
    finally:
        for p in paths:
            sys.path.remove(p)
        for k in [k for k in sys.modules if mine(k)]:
            del sys.modules[k]
        sys.modules.update(saved)
    return ref_pkg, ref_solver


if __name__ == "__main__":
    m = make(force=True)
    if m is None:
        print(f"{REF} is not mounted and oracle/_ref does not exist: nothing to do")
        sys.exit(1)
    print(f"oracle/_ref: {len(m['files'])} files from {m['source']}; verified={verify()}")
