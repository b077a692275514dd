"""CPU-side checks: the C-ABI library loads and exports every declared symbol, argument/config
validation happens before any CUDA call, host-side helpers, and the N>1 statistics reduction over
gloo (world_size 2)."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from rllib_warehouse_b200 import _native as nv
    L = nv.lib()
    header = open(os.path.join(ROOT, "include", "wh_b200.h")).read()
    declared = set(re.findall(r"\b(wh_[a-z_0-9]+)\s*\(", header))
    assert declared == set(nv.SYMBOLS), declared ^ set(nv.SYMBOLS)
    for s in declared:
        assert hasattr(L, s), s
    assert L.wh_version() >= 100
    assert L.wh_error_string(0) == b"ok"


def test_config_and_argument_validation_needs_no_gpu():
    from rllib_warehouse_b200 import LARGE, WarehouseConfig
    from rllib_warehouse_b200 import _native as nv
    L = nv.lib()
    cfg = nv.make_config(LARGE)
    assert L.wh_num_pickup_points(C.byref(cfg)) == 64 and L.wh_num_delivery_points(C.byref(cfg)) == 64
    st, ob = nv.State(), nv.Obs()
    # NULL state pointers -> WH_E_ARG, no launch
    assert L.wh_step(C.byref(cfg), C.byref(st), 4, 0, 0, None, None, None, None, None, None, None, None, 0, None, None) == 10002
    assert L.wh_build_obs(C.byref(cfg), C.byref(st), 4, 0, C.byref(ob), None) == 10002
    # unsupported geometry -> WH_E_CONFIG
    bad = nv.make_config(WarehouseConfig(16, 40, (4, 8, 12, 16)))      # D = 144 > 64
    assert L.wh_reset(C.byref(bad), C.byref(st), 4, 0, 0, None, None, None, None, None, None, None) == 10001
    bad = nv.make_config(WarehouseConfig(40, 20, (4, 8, 12, 16)))      # R > 32
    assert L.wh_reset(C.byref(bad), C.byref(st), 4, 0, 0, None, None, None, None, None, None, None) == 10001
    assert b"unsupported" in L.wh_error_string(10001)


def test_flags_match_the_header_and_multi_kernel_selection_is_validated():
    """The Python binding's flag constants are the header's; wh_multi_step rejects an unknown kernel selection, and a
    warp-specialised kernel without observations or for a non-variant geometry, before any CUDA call."""
    from rllib_warehouse_b200 import LARGE, WarehouseConfig
    from rllib_warehouse_b200 import _native as nv
    header = open(os.path.join(ROOT, "include", "wh_b200.h")).read()
    for name, val in (("AUTO_RESET", nv.FLAG_AUTO_RESET), ("COMPACT_IO", nv.FLAG_COMPACT_IO), ("NO_PDL", nv.FLAG_NO_PDL),
                      ("PER_STEP_OUT", nv.FLAG_PER_STEP_OUT)):
        m = re.search(r"#define\s+WH_FLAG_%s\s+(\d+)" % name, header)
        assert m and int(m.group(1)) == val, name
    assert "#define WH_FLAG_MULTI_KERNEL(k) (((k) & 7) << 4)" in header
    assert [nv.flag_multi_kernel(k) for k in ("auto", "throughput", "low_occupancy", "ws1", "ws2")] == [0, 16, 32, 48, 64]
    assert nv.flag_multi_kernel(None) == 0
    L = nv.lib()
    cfg = nv.make_config(LARGE)
    st, ob = nv.State(), nv.Obs()
    for f in nv.STATE_KEYS:
        setattr(st, f, 16)                       # non-NULL dummies: validation only, nothing is launched
    for f in nv.OBS_KEYS:
        setattr(ob, f, 16)
    args = lambda c, obs, flags: (C.byref(c), C.byref(st), 4, 0, 0, 3, None, 0, 0, 16, 16, None, obs, flags, None)  # noqa: E731
    assert L.wh_multi_step(*args(cfg, C.byref(ob), 5 << 4)) == 10002                       # kernel 5 does not exist
    assert L.wh_multi_step(*args(cfg, None, nv.flag_multi_kernel("ws2"))) == 10002          # ws kernels need obs
    odd = nv.make_config(WarehouseConfig(6, 14, (3, 7, 11), 30, 10, 6))                     # runtime-geometry kernels
    assert L.wh_multi_step(*args(odd, C.byref(ob), nv.flag_multi_kernel("ws1"))) == 10002
    assert L.wh_multi_step(*args(cfg, C.byref(ob), 128)) == 10002                          # unknown flag bit


def test_no_cpu_fallback():
    from rllib_warehouse_b200 import SMALL, BatchedWarehouse
    from rllib_warehouse_b200 import _native as nv
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(nv.NativeError):
        BatchedWarehouse(SMALL, 4)


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under the product package, the shims or the
    driver scripts may import, load or execute it (bench.py's CPU legs and smoke() are the only
    sanctioned users outside tests/)."""
    roots = [os.path.join(ROOT, d) for d in ("rllib_warehouse_b200", "warehouse", "baseline", "scripts", "include")]
    pat = re.compile(r"(^|\s)(from|import)\s+oracle\b|libwh_oracle|wh_oracle|ref_port")
    for root in roots:
        for dirpath, _, files in os.walk(root):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h")):
                    src = open(os.path.join(dirpath, f)).read()
                    assert not pat.search(src), os.path.join(dirpath, f)


def test_spaces_and_variant_constants():
    from rllib_warehouse_b200 import LARGE, MEDIUM, SMALL, spaces
    assert (SMALL.num_pickup_points, SMALL.num_delivery_points, SMALL.null_position) == (16, 32, 6)
    assert (MEDIUM.num_pickup_points, MEDIUM.num_delivery_points, MEDIUM.null_position) == (36, 48, 8)
    assert (LARGE.num_pickup_points, LARGE.num_delivery_points, LARGE.null_position) == (64, 64, 10)
    sp = spaces.observation_space(4, 12)
    ob = {"num_agents": np.array([3], np.int32), "self_position": np.array([1, 2], np.int32),
          "self_availability": np.array([1], np.int8), "self_delivery_target": np.array([6, 6], np.int32),
          "other_positions": np.zeros((3, 2), np.int32), "other_availabilities": np.zeros(3, np.int8),
          "other_delivery_targets": np.zeros((3, 2), np.int32), "requests": np.zeros((4, 4), np.int32)}
    assert sp.contains(ob)
    ob["self_position"] = np.array([1, 13], np.int32)
    assert not sp.contains(ob)
    assert spaces.Discrete(9).contains(8) and not spaces.Discrete(9).contains(9)


def test_shard_range_covers_everything():
    from rllib_warehouse_b200.parallel import shard_range
    for n, w in [(262144, 8), (1000, 3), (5, 8)]:
        r = [shard_range(n, k, w) for k in range(w)]
        assert r[0][0] == 0 and r[-1][1] == n and all(r[i][1] == r[i + 1][0] for i in range(w - 1))


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    from rllib_warehouse_b200 import _native as nv
    from rllib_warehouse_b200.parallel import allreduce_stats, shard_range, stats_to_metrics
    from oracle import wh_oracle as wo
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    n_total = 64
    lo, hi = shard_range(n_total, rank, world)
    env = wo.OracleEnv(wo.variant_config("small", random_num_agents=True), hi - lo, seed=11, env_id0=lo)
    env.reset()
    env.rollout(200, 1, policy="greedy", auto_reset=True)
    local = torch.from_numpy(env.stats.copy())
    total = allreduce_stats(local)
    q.put((rank, total.tolist(), stats_to_metrics(total, 4)))
    dist.destroy_process_group()


def test_stats_allreduce_gloo_world2():
    """Sharded rollout on 2 ranks (CPU stand-in for the per-GPU shards) + all-reduce == 1 rank."""
    import torch.multiprocessing as mp
    from oracle import wh_oracle as wo
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    whole = wo.OracleEnv(wo.variant_config("small", random_num_agents=True), 64, seed=11)
    whole.reset()
    whole.rollout(200, 1, policy="greedy", auto_reset=True)
    for _, tot, metrics in res:
        assert tot == whole.stats.tolist()
        assert metrics["episodes"] == 64 and "avg_agent_reward_all" in metrics


def _load_script(name):
    import importlib.util
    spec = importlib.util.spec_from_file_location(f"_script_{name}", os.path.join(ROOT, "scripts", f"{name}.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_rollout_cli_matches_the_reference(tmp_path):
    """scripts/rollout.py keeps the reference's command line (rollout.py:90-113: trial_dir num_agents
    [-i ITERATION] [-r], render off by default) and its checkpoint choice + messages (rollout.py:33-56)."""
    ro = _load_script("rollout")
    a = ro.parse(["/some/trial", "4", "-i", "100", "-r"])
    assert (a.trial_dir, a.num_agents, a.iteration, a.render) == ("/some/trial", 4, 100, True)
    a = ro.parse(["/some/trial", "2"])
    assert a.iteration == -1 and a.render is False and a.run == "SAC" and a.num_episodes == 1
    for n in (200, 400, 1000):
        (tmp_path / f"checkpoint_{n}").mkdir()
    it, path, msg = ro.pick_checkpoint(str(tmp_path), -1)
    assert it == 1000 and path.endswith(os.path.join("checkpoint_1000", "checkpoint-1000")) and "lastest" in msg
    it, _, msg = ro.pick_checkpoint(str(tmp_path), 400)
    assert it == 400 and "selected checkpoint at iteration 400" in msg
    it, _, msg = ro.pick_checkpoint(str(tmp_path), 450)
    assert it == 400 and "doesn't exist, loading the closest one at 400" in msg
    it, _, _ = ro.pick_checkpoint(str(tmp_path), 300)          # tie: the lower one, as min() over the sorted list does
    assert it == 200


def test_train_script_registers_the_vector_env_creator():
    """scripts/train.py: same env ids as the reference (train.py:29-33); the creator yields the vectorised
    BaseEnv adapter when asked to (needs CUDA to construct, so only the wiring is checked here) and the
    experiment specs carry the env_config it reads; the episode metrics are the reference's (train.py:18-23)."""
    import yaml
    tr = _load_script("train")
    assert set(tr.ENV_IDS) == {"WarehouseSmall-v0", "WarehouseMedium-v0", "WarehouseLarge-v0"}
    for size in ("small", "medium", "large"):
        spec = yaml.safe_load(open(os.path.join(ROOT, "scripts", "experiments", f"warehouse-{size}-ppo",
                                                f"warehouse-{size}-ppo.yaml")))
        (name, body), = spec.items()
        assert body["run"] == "PPO" and body["env"] in tr.ENV_IDS and tr.ENV_IDS[body["env"]] == size
        ec = body["config"]["env_config"]
        assert ec["vector"] is True and ec["flat_obs"] is True and ec["num_envs"] >= 1

    class Episode:
        agent_rewards = {("0", "p"): 3.0, ("1", "p"): 5.0}
        custom_metrics = {}
    tr.episode_metrics({"episode": Episode})
    assert Episode.custom_metrics == {"avg_agent_reward_all": [4.0], "avg_agent_reward_2": [4.0]}
    if not torch.cuda.is_available():
        from rllib_warehouse_b200 import _native as nv
        with pytest.raises(nv.NativeError):                     # the creator really builds the CUDA vector env
            tr.make_env("small", {"num_envs": 8}, vector=True)


def test_reference_copy_is_verbatim_and_tamper_evident(tmp_path, monkeypatch):
    """oracle/_ref (oracle/make_ref.py) is a verbatim copy of the reference's hot-path files: the manifest
    hashes match the mounted reference where it is available, and any edit of a copied file is detected."""
    import shutil
    from oracle import make_ref
    if not make_ref.available() and make_ref.make() is None:
        pytest.skip("neither /root/reference nor oracle/_ref on this box")
    assert make_ref.verify()
    ref = os.environ.get("WH_REFERENCE", "/root/reference")
    if os.path.isdir(os.path.join(ref, "warehouse")):
        for rel in make_ref.FILES:
            assert open(os.path.join(ref, rel), "rb").read() == open(os.path.join(make_ref.OUT, rel), "rb").read(), rel
    clone = tmp_path / "_ref"
    shutil.copytree(make_ref.OUT, clone)
    monkeypatch.setattr(make_ref, "OUT", str(clone))
    assert make_ref.verify()
    with open(clone / "warehouse" / "core.py", "a") as f:
        f.write("\n# edited\n")
    assert not make_ref.verify()
