"""BASELINE configs[0]: the reference's OWN `baseline/run.py` file (run.py:15-99), unmodified, executed
against this repo's drop-in shims on the GPU.

The file comes from `oracle/_ref/` — a verbatim, SHA-256-manifested copy of the reference made by
`oracle/make_ref.py` in the build container (git-ignored; it travels to the GPU box with the snapshot,
`/root/reference` does not exist there). Two pairings:

  * reference run.py + reference solvers.py (numpy on the host)  ->  this repo's `warehouse` package;
  * reference run.py  ->  this repo's `warehouse` package AND this repo's `solvers` shim (CUDA solver).
"""
import os
import re
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")
RUN_PY = os.path.join(REF, "baseline", "run.py")
STUBS = os.path.join(ROOT, "oracle", "stubs")


def _need_ref():
    sys.path.insert(0, ROOT)
    from oracle import make_ref
    if not make_ref.available():
        pytest.skip("oracle/_ref absent: run `python oracle/make_ref.py` where /root/reference is mounted")
    assert make_ref.verify(), "oracle/_ref differs from its manifest: not the unmodified reference"


def _check(res, steps=200):
    assert res.returncode == 0, res.stdout[-1500:] + res.stderr[-3000:]
    assert f"=== Done ({steps} steps) ===" in res.stdout, res.stdout[-500:]
    m = re.search(r"Total: ([0-9.]+), Per Agent: ([0-9.]+)", res.stdout)
    assert m, res.stdout[-500:]
    return float(m.group(1))


@pytest.mark.parametrize("size,agents,prob", [("small", 4, 0.0), ("medium", 5, 0.1), ("large", 16, 0.0)])
def test_reference_run_py_and_reference_solver_on_the_cuda_env(size, agents, prob):
    """`python oracle/_ref/baseline/run.py small 4 0.0`: script dir first on sys.path, so `solvers` is the
    reference's own solvers.py; `warehouse` resolves to this repo's package (PYTHONPATH). The file itself
    asserts observation_space.contains() for every observation of every step (run.py:36-37,58-59)."""
    _need_ref()
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([ROOT, STUBS]))
    res = subprocess.run([sys.executable, RUN_PY, size, str(agents), str(prob)], capture_output=True, text=True,
                         timeout=600, env=env, cwd=ROOT)
    total = _check(res)
    assert total >= 1.0          # the greedy baseline always earns something in 200 steps


def test_reference_run_py_on_both_shims():
    """The same unmodified file, but `solvers` resolves to this repo's shim (CUDA solver kernel)."""
    _need_ref()
    code = (
        "import runpy, sys\n"
        f"sys.path[:0] = [{os.path.join(ROOT, 'baseline')!r}, {ROOT!r}]\n"
        "import solvers, warehouse\n"
        f"assert solvers.__file__.startswith({os.path.join(ROOT, 'baseline')!r}), solvers.__file__\n"
        f"assert warehouse.__file__.startswith({os.path.join(ROOT, 'warehouse')!r}), warehouse.__file__\n"
        "sys.argv = ['run.py', 'small', '4', '0.0']\n"
        f"runpy.run_path({RUN_PY!r}, run_name='__main__')\n"
    )
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert _check(res) >= 1.0
