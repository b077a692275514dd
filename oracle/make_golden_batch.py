#!/usr/bin/env python
"""Full-size reference digests — TEST INFRASTRUCTURE ONLY (never imported by the product path).

BASELINE.json configs[1] asks for "smallest variant, 4096 parallel envs, random actions
(bit-exact step validation vs reference)", SURVEY.md §8d config 2 spells it out: env e is the
UNMODIFIED reference run with `np.random.seed(BASE + e)`, actions come from a separate
`Generator(PCG64(SEED_A))` tensor [T, N, A], T = 200 (so the step-200 mass expiry is covered) and
every state tensor, observation key, reward and done flag must be bit-identical at every step.

The reference cannot travel to the GPU box and 4096 x 200 full outputs would be ~0.5 GB, so this
script records, per batch:
  * the replayable semantic draws (reset: accepted agent cells, initial request pickup / delivery
    ids; step: respawned pickup / delivery ids — the protocol of oracle/make_golden.py),
  * for every step and every output array a CRC-32 of the whole [N, ...] array (canonical dtype:
    int32 for state / observations / actions, float32 rewards, uint8 dones), plus the final state.
`tests/golden_util.check_batch_digests` replays the draws on any env-like object (C oracle, CUDA
path) and compares the CRCs. Batches (committed under tests/golden/batch_*.npz):
  small_random   4096 envs x 200 steps, A = 4, random actions          (configs[1])
  medium_greedy  1024 envs x 200 steps, A = 9, reference greedy solver (configs[2] replay subset)
  large_random    512 envs x 200 steps, A = 16, random actions         (configs[3] replay subset)
  small_train_greedy / large_train_random   2048 / 256 envs of the *Train variants: per-env random agent
                 count (rows >= A of every [N,R,...] array are padding and hold -1 / 0 in the digest)

    python oracle/make_golden_batch.py      # needs /root/reference; ~2 min on 8 cores
"""
import multiprocessing as mp
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402  (imports the unmodified reference under the stubs)

BASE_SEED = 20260000
SEED_A = 4242
T = 200
BATCHES = {
    "small_random": dict(size="small", n=4096, policy="random"),
    "medium_greedy": dict(size="medium", n=1024, policy="greedy"),
    "large_random": dict(size="large", n=512, policy="random"),
    # *Train variants (variants.py:65-98): the agent count is redrawn per env, rows >= A are padding
    "small_train_greedy": dict(size="small", n=2048, policy="greedy", train=True),
    "large_train_random": dict(size="large", n=256, policy="random", train=True),
}
STATE_KEYS = ("agent_pos", "agent_tgt", "pickup_tgt", "pickup_timer", "time", "num_agents")
OUT_KEYS = STATE_KEYS + tuple("obs_" + k for k in mg.OBS_KEYS) + ("actions", "rewards", "dones")


def crc(a, dtype):
    return zlib.crc32(np.ascontiguousarray(a, dtype=dtype).tobytes())


def canon_dtype(k):
    return np.float32 if k == "rewards" else (np.uint8 if k == "dones" else np.int32)


def _one_env(args):
    size, e, policy, actions, train = args
    cls, train_cls = mg.VARIANTS[size]
    np.random.seed(BASE_SEED + e)
    with mg.ChoiceRecorder() as rec:
        env = train_cls() if train else cls(cls.max_num_agents)
        rec.take()
        obs = env.reset()
        A = env.num_agents                      # *Train: redrawn by reset (variants.py:69-74)
        R = env.num_requests
        init_p, init_t = rec.take()
        solver = mg.WarehouseRandomGreedySolver(A, R, 0.0, env.action_space)
        pos0 = np.full((R, 2), -1, np.int8)
        pos0[:A] = env._agent_positions
        out = dict(reset_agent_pos=pos0, reset_num_agents=np.int8(A),
                   reset_init_pickups=init_p.astype(np.int8), reset_init_targets=init_t.astype(np.int8))
        reset_rec = dict(mg.snap_state(env, R))
        reset_rec.update(mg.snap_obs(obs, A, R))
        steps = []
        for t in range(T):
            if policy == "greedy":
                ad = solver.compute_action(obs)
                act = mg.pad([int(ad[str(i)]) for i in range(A)], R)
            else:
                act = mg.pad(actions[t][:A], R)
                ad = {str(i): int(act[i]) for i in range(A)}
            obs, rew, dones, _ = env.step(ad)
            sp, st = rec.take()
            r = dict(spawn_pickups=mg.pad(sp, R).astype(np.int8), spawn_targets=mg.pad(st, R).astype(np.int8),
                     actions=act.astype(np.int8))
            r.update(mg.snap_state(env, R))
            r.update(mg.snap_obs(obs, A, R))
            r["rewards"] = np.array([rew[str(i)] for i in range(A)] + [0.0] * (R - A), np.float32)
            r["dones"] = np.uint8(dones["__all__"])
            steps.append(r)
    out["reset"] = mg.shrink(reset_rec)
    out["steps"] = mg.shrink(mg.stack(steps))
    return out


def make_batch(name, size, n, policy, pool, train=False):
    cls, _ = mg.VARIANTS[size]
    A = cls.max_num_agents
    actions = np.random.Generator(np.random.PCG64(SEED_A)).integers(0, 9, size=(T, n, A)).astype(np.int32)
    res = pool.map(_one_env, [(size, e, policy, actions[:, e] if policy == "random" else None, train)
                              for e in range(n)], chunksize=16)
    fx = dict(mg.env_dims(cls(A)))
    fx.update(n=np.int32(n), T=np.int32(T), A=np.int32(A), base_seed=np.int64(BASE_SEED),
              seed_actions=np.int64(SEED_A), policy=np.array(policy), out_keys=np.array(OUT_KEYS),
              train=np.int8(train))
    for k in ("reset_agent_pos", "reset_init_pickups", "reset_init_targets", "reset_num_agents"):
        fx[k] = np.stack([r[k] for r in res])
    for k in ("spawn_pickups", "spawn_targets"):
        fx[k] = np.stack([r["steps"][k] for r in res], axis=1)               # [T, n, R]
    reset_keys = STATE_KEYS + tuple("obs_" + k for k in mg.OBS_KEYS)
    fx["reset_keys"] = np.array(reset_keys)
    fx["reset_crc"] = np.array([crc(np.stack([r["reset"][k] for r in res]), np.int32) for k in reset_keys],
                               np.uint32)
    table = np.zeros((T, len(OUT_KEYS)), np.uint32)
    for j, k in enumerate(OUT_KEYS):
        full = np.stack([r["steps"][k] for r in res], axis=1)               # [T, n, ...]
        for t in range(T):
            table[t, j] = crc(full[t], canon_dtype(k))
        if k in STATE_KEYS:
            fx["final_" + k] = full[-1]
        if k == "rewards":
            fx["return_per_env"] = full.sum(axis=(0, 2)).astype(np.float32)
    fx["step_crc"] = table
    fx["actions_crc_all"] = np.uint32(crc(actions, np.int32))                # guards the PCG64 stream
    path = os.path.join(mg.OUT, f"batch_{name}.npz")
    np.savez_compressed(path, **fx)
    print(name, os.path.getsize(path), "bytes; mean return/env", float(fx["return_per_env"].mean()))


def main():
    os.makedirs(mg.OUT, exist_ok=True)
    with mp.Pool(os.cpu_count()) as pool:
        for name, kw in BATCHES.items():
            if len(sys.argv) > 1 and name not in sys.argv[1:]:
                continue
            make_batch(name, pool=pool, **kw)


if __name__ == "__main__":
    main()
