"""bench.py contract checks that need no GPU: the reference arm prints exactly one JSON line with
the agreed keys; the CUDA arm refuses to run without a device (no CPU fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _reference_line(extra_env=None):
    env = dict(os.environ, **(extra_env or {}))
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "3",
                          "--warmup", "1", "--cpu-envs", "256", "--cpu-seconds", "2"], capture_output=True, text=True,
                         timeout=600, env=env)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "agent_steps_per_sec" and d["unit"] == "agent-steps/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["value"] > 0
    assert d["config"]["workload"].startswith("warehouse-large-262144") and d["config"]["envs_per_gpu"] == 262144
    # the bounded CPU sample is stated, not hidden behind the config's 262 144 envs
    assert d["config"]["cpu_sample_envs_per_step"] >= 1 and "bounded sample" in d["config"]["workload"]
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "agent-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    return d


def test_reference_arm_json_line():
    """With oracle/_ref present (made by oracle/make_ref.py where the reference is mounted) the arm times
    the UNMODIFIED reference: kind "reference", single-process and C-port legs beside it."""
    sys.path.insert(0, ROOT)
    from oracle import make_ref
    if not make_ref.available():
        if make_ref.make() is None:
            import pytest
            pytest.skip("neither /root/reference nor oracle/_ref on this box")
    assert make_ref.verify()
    d = _reference_line()
    cb = d["cpu_baseline"]
    assert cb["kind"] == "reference" and "unmodified reference" in cb["sample"]
    assert cb["single_process"]["kind"] == "reference" and cb["single_process"]["cores"] == 1
    assert cb["c_port"]["kind"] == "port" and cb["c_port"]["value"] > cb["value"]     # C beats Python + numpy


def test_reference_arm_without_the_reference_copy():
    d = _reference_line({"WH_BENCH_NO_REF": "1"})
    assert d["cpu_baseline"]["kind"] == "port" and "absent" in d["cpu_baseline"]["note"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, timeout=120, env=env)
    assert res.returncode == 0 and res.stdout.strip() == ""


def test_cuda_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("CUDA present")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=300)
    assert res.returncode != 0 and res.stdout.strip() == ""
    assert "no CPU fallback" in res.stderr or "CUDA" in res.stderr


import pytest  # noqa: E402


@pytest.mark.gpu
def test_cuda_arm_json_contract_on_a_small_workload():
    """The CUDA arm end to end on a reduced workload (the full one is the driver's): one JSON line with the
    contract's keys — roofline (bound, achieved, peak, frac, traffic), the three e2e consumers with their copied
    bytes, cpu_baseline with a kind, clocks, and exactly one kernel launch per timed step."""
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "6", "--warmup", "3", "--envs", "8192",
                          "--e2e-steps", "6", "--cpu-seconds", "2", "--cpu-envs", "256", "--no-extras"],
                         capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stderr[-3000:]
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["metric"] == "agent_steps_per_sec" and d["unit"] == "agent-steps/s" and d["n_gpus"] == 1
    assert d["steps"] == 6 and d["warmup"] == 3 and d["gpu_launches"] == 6 and d["launches_per_step"] == 1
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None and d["dtype"] == "int32"
    assert d["config"]["workload"] == "warehouse-large-8192-envs-per-gpu" and d["config"]["envs_total"] == 8192
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and r["peak"] > 1000 and 0 < r["frac"] < 1.3
    assert abs(r["achieved"] - r["alg_bytes_per_launch"] / (r["kernel_ms_avg"] * 1e-3) / 1e9) < 1e-6 * r["achieved"]
    assert r["alg_bytes_per_launch"] == 9147 * 8192 and "traffic" in r and "traffic_source" in r
    for k, h2d, d2h in (("e2e", 8192 * 16 * 4, 8192 * 16 * 4 + 8192), ("e2e_alt", 8192 * 16, 8192 * 16 + 8192)):
        assert d[k]["value"] > 0 and d[k]["h2d_bytes_per_step"] == h2d and d[k]["d2h_bytes_per_step"] == d2h
        assert d[k]["value"] != d["value"] and d[k]["gpu_launches"] >= 6
    assert d["e2e_host_obs"]["d2h_bytes_per_step"] > 8192 * 8512 and d["e2e_copy_pipeline"]["chunks"] == 8
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] > 0
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"} and "collective_ms" in d
