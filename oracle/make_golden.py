#!/usr/bin/env python
"""Golden-vector generator — TEST INFRASTRUCTURE ONLY (never imported by the product path).

Runs the UNMODIFIED reference (`/root/reference/warehouse/core.py`, `variants.py`,
`baseline/solvers.py`) under the stub `gym`/`ray` packages in `oracle/stubs/` and records
inputs, replayable RNG draws and every output into small `.npz` fixtures under `tests/golden/`.
The reference has no tests or known-answer vectors of its own, so these fixtures
— produced by executing the reference itself — are what pins the oracle (`oracle/ref_port.py`,
`oracle/wh_oracle.c`) and, through it, the CUDA path.

Only runnable where `/root/reference` exists (the build container). The fixtures it writes are
committed; the GPU box never needs the reference.

    python oracle/make_golden.py            # regenerates tests/golden/*.npz

Replay protocol: the reference draws from the process-global legacy
`np.random` stream (core.py:196-197, 215-220, 339-350; variants.py:74). We record the *semantic*
draws instead of emulating MT19937:
  reset : accepted agent cells [A,2]; initial request pickup ids [R] + delivery ids [R]
  step  : respawned pickup ids [k] + delivery ids [k]  (k = R - #active, padded with -1 to R)
  Train : the redrawn num_agents
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("WH_REFERENCE", "/root/reference")


def _import_reference():
    """Imports the reference's `warehouse` package and `baseline/solvers.py` under the stub
    gym/ray WITHOUT leaving them in sys.modules / sys.path: this repo ships its own `warehouse`
    and `solvers` modules under the same names (the drop-in shims), and the two must not shadow
    each other inside one test process."""
    names = ("warehouse", "solvers", "gym", "ray")
    mine = lambda k: k in names or k.startswith(tuple(n + "." for n in names))
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if mine(k)}
    paths = [os.path.join(HERE, "stubs"), REF, os.path.join(REF, "baseline")]
    sys.path[:0] = paths
    try:
        import warehouse as ref_pkg
        from solvers import WarehouseRandomGreedySolver as ref_solver
        assert os.path.realpath(ref_pkg.__file__).startswith(os.path.realpath(REF)), ref_pkg.__file__
    finally:
        for p in paths:
            sys.path.remove(p)
        for k in [k for k in sys.modules if mine(k)]:
            del sys.modules[k]
        sys.modules.update(saved)
    return ref_pkg, ref_solver


refwh, WarehouseRandomGreedySolver = _import_reference()

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")

VARIANTS = {
    "small": (refwh.WarehouseSmall, refwh.WarehouseSmallTrain),
    "medium": (refwh.WarehouseMedium, refwh.WarehouseMediumTrain),
    "large": (refwh.WarehouseLarge, refwh.WarehouseLargeTrain),
}

OBS_KEYS = [
    "num_agents", "self_position", "self_availability", "self_delivery_target",
    "other_positions", "other_availabilities", "other_delivery_targets", "requests",
]


class ChoiceRecorder:
    """Wraps np.random.choice and logs every call's (population, size, result)."""

    def __init__(self):
        self.log = []
        self._orig = np.random.choice

    def __enter__(self):
        def wrapped(a, size=None, replace=True, p=None):
            out = self._orig(a, size, replace, p)
            self.log.append(np.array(out, dtype=np.int64).reshape(-1))
            return out

        np.random.choice = wrapped
        return self

    def __exit__(self, *exc):
        np.random.choice = self._orig

    def take(self):
        log, self.log = self.log, []
        return log


def env_dims(env):
    return dict(
        R=env._num_requests, P=env._num_pickup_points, D=env._num_delivery_points,
        dim=env._area_dimension, racks=np.array(env._pickup_racks_arrangement, np.int32),
        episode=env._episode_duration, wait=env._pickup_wait_duration,
    )


def snap_state(env, R):
    """State padded to R agent rows (rows >= A hold -1)."""
    A = env._num_agents
    pos = np.full((R, 2), -1, np.int32)
    pos[:A] = env._agent_positions
    tgt = np.full(R, -1, np.int32)
    tgt[:A] = env._agent_delivery_targets
    return dict(
        agent_pos=pos, agent_tgt=tgt,
        pickup_tgt=env._pickup_point_targets.astype(np.int32).copy(),
        pickup_timer=env._pickup_point_timers.astype(np.int32).copy(),
        time=np.int32(env._episode_time), num_agents=np.int32(A),
    )


def snap_obs(obs, A, R):
    """Obs dict-of-dicts -> fixed [R, ...] arrays per key; rows >= A are -1 (not produced)."""
    shapes = {
        "num_agents": (1,), "self_position": (2,), "self_availability": (1,),
        "self_delivery_target": (2,), "other_positions": (R - 1, 2),
        "other_availabilities": (R - 1,), "other_delivery_targets": (R - 1, 2),
        "requests": (R, 4),
    }
    out = {}
    for k in OBS_KEYS:
        arr = np.full((R,) + shapes[k], -1, np.int32)
        for i in range(A):
            v = obs[str(i)][k]
            assert v.shape == shapes[k], (k, v.shape)
            assert v.dtype == (np.int8 if "availab" in k else np.int32), (k, v.dtype)
            arr[i] = v
        out["obs_" + k] = arr
    return out


def pad(v, n, fill=-1):
    out = np.full(n, fill, np.int32)
    out[: len(v)] = v
    return out


def stack(records):
    keys = records[0].keys()
    return {k: np.stack([r[k] for r in records]) for k in keys}


def shrink(d):
    """int8 is enough for everything except timers/time (<= 32767)."""
    out = {}
    for k, v in d.items():
        v = np.asarray(v)
        if v.dtype.kind == "f":
            out[k] = v.astype(np.float32)
        elif v.dtype.kind in "iu" and v.size and -128 <= v.min() and v.max() <= 127:
            out[k] = v.astype(np.int8)
        elif v.dtype.kind in "iu":
            out[k] = v.astype(np.int16) if (v.size == 0 or (-32768 <= v.min() and v.max() <= 32767)) else v
        else:
            out[k] = v
    return out


def run_episode(size, A, seed, policy, train=False, T=None, rand_prob=0.0):
    """One seeded episode of the reference. policy: 'random' | 'greedy'."""
    fixed_cls, train_cls = VARIANTS[size]
    np.random.seed(seed)
    act_rng = np.random.Generator(np.random.PCG64(seed + 7919))
    with ChoiceRecorder() as rec:
        env = train_cls() if train else fixed_cls(A)
        rec.take()
        obs = env.reset()
        A = env.num_agents
        dims = env_dims(env)
        R = dims["R"]
        init_pick, init_tgt = rec.take()
        out = dict(dims)
        out.update(
            {"reset_" + k: v for k, v in snap_state(env, R).items()},
            reset_init_pickups=init_pick, reset_init_targets=init_tgt,
        )
        out.update({"reset_" + k: v for k, v in snap_obs(obs, A, R).items()})
        solver = WarehouseRandomGreedySolver(A, R, rand_prob, env.action_space)
        steps = []
        T = T or dims["episode"]
        for _ in range(T):
            if policy == "greedy":
                ad = solver.compute_action(obs)
                actions = pad([int(ad[str(i)]) for i in range(A)], R)
            else:
                actions = pad(act_rng.integers(0, 9, size=A), R)
                ad = {str(i): int(actions[i]) for i in range(A)}
            obs, rew, dones, infos = env.step(ad)
            sp, st = rec.take()
            r = dict(actions=actions, spawn_pickups=pad(sp, R), spawn_targets=pad(st, R))
            r.update(snap_state(env, R))
            r.update(snap_obs(obs, A, R))
            r["rewards"] = np.array([rew[str(i)] for i in range(A)] + [0.0] * (R - A), np.float32)
            assert all(type(rew[str(i)]) is np.float32 for i in range(A))
            assert set(dones) == {str(i) for i in range(A)} | {"__all__"}
            assert len({bool(v) for v in dones.values()}) == 1
            r["done"] = np.int8(dones["__all__"])
            assert infos == {str(i): {} for i in range(A)}
            steps.append(r)
        out.update(stack(steps))
    return shrink(out)


def pickup_cells(env):
    return {tuple(p) for p in env._pickup_point_positions.tolist()}


def run_single_steps(size, n_cases, seed):
    """Random injected states (co-location, border cells, pickup cells, near-expiry timers),
    random action-dict ORDER and ABSENT agents; one reference step each."""
    fixed_cls, _ = VARIANTS[size]
    rng = np.random.Generator(np.random.PCG64(seed))
    np.random.seed(seed)
    cases = []
    dims = None
    with ChoiceRecorder() as rec:
        for _ in range(n_cases):
            Rmax = fixed_cls.max_num_agents
            A = int(rng.integers(1, Rmax + 1))
            env = fixed_cls(A)
            env.reset()
            rec.take()
            dims = env_dims(env)
            R, P, D, dim, wait = dims["R"], dims["P"], dims["D"], dims["dim"], dims["wait"]
            # --- inject a random state -------------------------------------------------
            pos = rng.integers(0, dim, size=(A, 2))
            mode = rng.integers(0, 4)
            if mode == 0:        # crowd agents into a 4x4 patch: many collisions / co-location
                pos = rng.integers(0, 4, size=(A, 2)) + rng.integers(0, dim - 3, size=(1, 2))
            elif mode == 1:      # put agents onto pickup cells (some on active requests)
                pc = env._pickup_point_positions
                pos = pc[rng.integers(0, P, size=A)]
            for a in range(1, A):  # explicit co-location
                if rng.random() < 0.15:
                    pos[a] = pos[rng.integers(0, a)]
            tgt = np.where(rng.random(A) < 0.5, -1, rng.integers(0, D, size=A))
            for a in range(A):   # sometimes place a delivering agent next to / on its target
                if tgt[a] >= 0 and rng.random() < 0.4:
                    d = env._delivery_point_positions[tgt[a]]
                    pos[a] = np.clip(d + rng.integers(-1, 2, size=2), 0, dim - 1)
            active = rng.permutation(P)[:R]
            ptgt = np.full(P, -1, np.int64)
            ptim = np.full(P, -1, np.int64)
            ptgt[active] = rng.integers(0, D, size=R)
            ptim[active] = np.where(rng.random(R) < 0.3, 1, rng.integers(1, wait + 1, size=R))
            env._agent_positions = pos.astype(np.int32)
            env._agent_delivery_targets = tgt.astype(np.int32)
            env._pickup_point_targets = ptgt.astype(np.int32)
            env._pickup_point_timers = ptim.astype(np.int32)
            env._episode_time = int(rng.choice([0, 1, 50, 198, 199, 200, 250]))
            pre = {"pre_" + k: v for k, v in snap_state(env, R).items()}
            # --- action dict: random order, random absences ----------------------------
            actions = np.full(R, -1, np.int64)
            present = [a for a in range(A) if rng.random() > 0.15]
            for a in present:
                actions[a] = rng.integers(0, 9)
            order = list(present)
            if rng.random() < 0.6:
                rng.shuffle(order)
            ad = {str(a): int(actions[a]) for a in order}
            obs, rew, dones, _ = env.step(ad)
            sp, st = rec.take()
            r = dict(pre)
            r.update(actions=actions.astype(np.int32), order=pad(order, R),
                     spawn_pickups=pad(sp, R), spawn_targets=pad(st, R))
            r.update(snap_state(env, R))
            r.update(snap_obs(obs, A, R))
            r["rewards"] = np.array([rew[str(i)] for i in range(A)] + [0.0] * (R - A), np.float32)
            r["done"] = np.int8(dones["__all__"])
            cases.append(r)
    out = dict(dims)
    out.update(stack(cases))
    return shrink(out)


def run_solver_cases(size, n_cases, seed, rand_prob):
    """Reference solver on reference obs with eps-random actions: records the uniform draws'
    outcome (is_random mask) and gym-sampled actions so the solver can be replayed exactly."""
    fixed_cls, _ = VARIANTS[size]
    np.random.seed(seed)
    A = fixed_cls.max_num_agents
    env = fixed_cls(A)
    obs = env.reset()
    R = env.num_requests
    solver = WarehouseRandomGreedySolver(A, R, rand_prob, env.action_space)
    uni_log, samp_log = [], []
    orig_uniform, orig_sample = np.random.uniform, env.action_space.sample

    def uniform(*a, **k):
        v = orig_uniform(*a, **k)
        uni_log.append(v)
        return v

    def sample():
        v = orig_sample()
        samp_log.append(v)
        return v

    np.random.uniform = uniform
    env.action_space.sample = sample
    recs = []
    try:
        for _ in range(n_cases):
            uni_log.clear()
            samp_log.clear()
            ad = solver.compute_action(obs)
            u = np.array(uni_log)
            is_rand = (u < rand_prob)
            ra = np.full(A, -1, np.int32)
            ra[is_rand] = samp_log
            r = dict(snap_obs(obs, A, R))
            r.update(is_random=is_rand.astype(np.int8), random_actions=ra, uniforms=u,
                     actions=np.array([int(ad[str(i)]) for i in range(A)], np.int32))
            recs.append(r)
            obs, _, dones, _ = env.step(ad)
            if dones["__all__"]:
                obs = env.reset()
    finally:
        np.random.uniform = orig_uniform
    out = dict(env_dims(env), rand_prob=np.float64(rand_prob))
    out.update(stack(recs))
    out = shrink(out)
    out["uniforms"] = np.stack([r["uniforms"] for r in recs]).astype(np.float64)
    return out


def scenario(env, pos, tgt, ptgt_pairs, timers, time, action_items, rec):
    """Hand-built state -> one reference step. ptgt_pairs: {pickup_idx: delivery_idx}."""
    R, P = env._num_requests, env._num_pickup_points
    A = env._num_agents
    env._agent_positions = np.array(pos, np.int32).reshape(A, 2)
    env._agent_delivery_targets = np.array(tgt, np.int32)
    pt = np.full(P, -1, np.int32)
    tm = np.full(P, -1, np.int32)
    for i, (p, d) in enumerate(ptgt_pairs.items()):
        pt[p] = d
        tm[p] = timers[i] if timers is not None else env._pickup_wait_duration
    env._pickup_point_targets, env._pickup_point_timers = pt, tm
    env._episode_time = time
    pre = {"pre_" + k: v for k, v in snap_state(env, R).items()}
    rec.take()
    actions = np.full(R, -1, np.int32)
    order = []
    for a, act in action_items:
        actions[a] = act
        order.append(a)
    obs, rew, dones, _ = env.step({str(a): int(act) for a, act in action_items})
    sp, st = rec.take()
    r = dict(pre)
    r.update(actions=actions, order=pad(order, R), spawn_pickups=pad(sp, R), spawn_targets=pad(st, R))
    r.update(snap_state(env, R))
    r.update(snap_obs(obs, A, R))
    r["rewards"] = np.array([rew[str(i)] for i in range(A)] + [0.0] * (R - A), np.float32)
    r["done"] = np.int8(dones["__all__"])
    return r


def run_quirk_scenarios():
    """Named hand-built cases for the quirk list in SURVEY.md §8a (Small variant, R=4, dim=12).
    Small pickup cells: racks [4,8] -> {3,4}x{3,4}, {3,4}x{7,8}, {7,8}x{3,4}, {7,8}x{7,8}.
    MOVES[a] = (a//3-1, a%3-1): 0=(-1,-1) 1=(-1,0) 2=(-1,1) 3=(0,-1) 4=stay 5=(0,1) 6=(1,-1) 7=(1,0) 8=(1,1)."""
    np.random.seed(1234)
    names, recs = [], []
    req = {0: 3, 5: 7, 10: 11, 15: 20}  # 4 active requests, pickup idx -> delivery idx
    with ChoiceRecorder() as rec:
        def mk(A):
            e = refwh.WarehouseSmall(A)
            e.reset()
            return e

        def add(name, env, *a, **k):
            names.append(name)
            recs.append(scenario(env, *a, rec=rec, **k))

        # 1. dict order matters: two agents want the same free cell (5,5)
        add("order_ascending", mk(2), [[5, 4], [5, 6]], [-1, -1], req, None, 10, [(0, 5), (1, 3)])
        add("order_reversed", mk(2), [[5, 4], [5, 6]], [-1, -1], req, None, 10, [(1, 3), (0, 5)])
        # 2. per-axis clamp => wall sliding: (0,5) action 2 (-1,+1) -> (0,6); corner stays
        add("wall_slide", mk(2), [[0, 5], [11, 11]], [-1, -1], req, None, 10, [(0, 2), (1, 8)])
        # 3. anti-swap: 0 moves (5,5)->(6,5)? blocked (occupied); swap attempt both directions
        add("anti_swap", mk(2), [[5, 5], [6, 5]], [-1, -1], req, None, 10, [(0, 7), (1, 1)])
        # chain-follow: 0 vacates (6,5)->(7,5)... later agent 1 moves into the vacated cell
        add("chain_follow_ok", mk(2), [[6, 5], [5, 5]], [-1, -1], req, None, 10, [(0, 7), (1, 7)])
        add("chain_follow_blocked", mk(2), [[6, 5], [5, 5]], [-1, -1], req, None, 10, [(1, 7), (0, 7)])
        # anti-cross: 0 goes (5,5)->(6,6) diagonal; 1 at (5,6) wants (6,5) crossing it; and reverse order
        add("anti_cross_a", mk(2), [[5, 5], [5, 6]], [-1, -1], req, None, 10, [(0, 8), (1, 6)])
        add("anti_cross_b", mk(2), [[5, 5], [6, 5]], [-1, -1], req, None, 10, [(0, 8), (1, 2)])
        add("anti_cross_c", mk(2), [[5, 6], [5, 5]], [-1, -1], req, None, 10, [(0, 6), (1, 8)])
        # 4. co-location: 0,1 share (5,5); 0 leaves -> occ False -> 2 may enter although 1 remains
        add("coloc_leave_enter", mk(3), [[5, 5], [5, 5], [5, 6]], [-1, -1, -1], req, None, 10,
            [(0, 7), (2, 3)])
        # ... but if 1 'stays' (action 4) in between it re-marks the cell and blocks 2
        add("coloc_stay_remarks", mk(3), [[5, 5], [5, 5], [5, 6]], [-1, -1, -1], req, None, 10,
            [(0, 7), (1, 4), (2, 3)])
        # ... absent != stay: 1 absent, same as coloc_leave_enter but with explicit absence noted
        add("coloc_absent", mk(3), [[5, 5], [5, 5], [5, 6]], [-1, -1, -1], req, None, 10,
            [(2, 3), (0, 7)])
        # co-located swap-back: 0 leaves (5,5)->(6,5); 1 (co-located) follows; 2 at (6,5)?? occupied start
        add("coloc_follow", mk(3), [[5, 5], [5, 5], [7, 5]], [-1, -1, -1], req, None, 10,
            [(0, 7), (1, 7), (2, 1)])
        # 5. expiry precedes pickup: agent steps onto pickup 0 (cell (3,3)) whose timer hits 0 now
        add("expiry_before_pickup", mk(1), [[2, 3]], [-1], req, [1, 200, 200, 200], 10, [(0, 7)])
        add("pickup_timer_2", mk(1), [[2, 3]], [-1], req, [2, 200, 200, 200], 10, [(0, 7)])
        # mass expiry: all R requests expire in one step -> k = R respawns
        add("mass_expiry", mk(2), [[1, 1], [10, 10]], [-1, -1], req, [1, 1, 1, 1], 199, [(0, 4), (1, 4)])
        # 6/7. pickup then same-step status: agent moves onto active pickup -> reward 1, availability 0
        add("pickup_basic", mk(2), [[2, 3], [10, 10]], [-1, -1], req, None, 10, [(0, 7), (1, 4)])
        # busy agent walks over an active pickup: no pickup
        add("pickup_busy_agent", mk(1), [[2, 3]], [5], req, None, 10, [(0, 7)])
        # agent standing still on an active request picks it up (respawn-under-agent follow-up)
        add("pickup_standing", mk(1), [[3, 3]], [-1], req, None, 10, [(0, 4)])
        # forced co-location on a pickup cell: both get target + reward (core.py:320-331)
        add("pickup_coloc_forced", mk(2), [[3, 3], [3, 3]], [-1, -1], req, None, 10, [(0, 4), (1, 4)])
        # delivery: delivery idx 3 -> v=2, side 3 -> (11,2); agent next to it moves on
        add("delivery_basic", mk(1), [[10, 2]], [3], req, None, 10, [(0, 7)])
        add("delivery_not_yet", mk(1), [[9, 2]], [3], req, None, 10, [(0, 7)])
        # delivery idx 0 -> (2,0); agent already standing on it delivers without moving
        add("delivery_standing", mk(1), [[2, 0]], [0], req, None, 10, [(0, 4)])
        # 9. other_delivery_targets deletes row 1 (not row i): 3 delivering agents, distinct targets
        add("obs_row1_quirk", mk(4), [[1, 1], [1, 5], [5, 1], [9, 9]], [0, 5, 10, -1], req, None, 10,
            [(0, 4), (1, 4), (2, 4), (3, 4)])
        add("obs_row1_quirk_A1", mk(1), [[1, 1]], [7], req, None, 10, [(0, 4)])
        # done flag at time == episode_duration, and stepping past done
        add("done_at_200", mk(1), [[1, 1]], [-1], req, None, 199, [(0, 4)])
        add("past_done", mk(1), [[1, 1]], [-1], req, None, 200, [(0, 5)])
        add("not_done_198", mk(1), [[1, 1]], [-1], req, None, 198, [(0, 5)])
        # empty action dict: nobody moves, world still advances
        add("empty_actions", mk(3), [[1, 1], [2, 2], [5, 5]], [-1, 4, -1], req, [1, 5, 5, 5], 10, [])
    out = dict(env_dims(refwh.WarehouseSmall(4)))
    out.update(stack(recs))
    out = shrink(out)
    out["names"] = np.array(names)
    return out


def main():
    os.makedirs(OUT, exist_ok=True)
    # full seeded episodes (T=200 so the step-200 mass expiry is covered), random + greedy policies
    for size, Amax in (("small", 4), ("medium", 9), ("large", 16)):
        eps = {}
        plan = [(Amax, "random", 11), (Amax, "greedy", 12), (max(1, Amax // 2), "greedy", 13),
                (1, "random", 14)]
        for i, (A, policy, seed) in enumerate(plan):
            ep = run_episode(size, A, 1000 * (1 + len(size)) + seed, policy, T=210)
            for k, v in ep.items():
                eps[f"ep{i}_{k}"] = v
            eps[f"ep{i}_policy"] = np.array(policy)
        # *Train variant: random agent count redrawn at construction and on reset
        for j in range(3):
            ep = run_episode(size, None, 5000 + 17 * j + len(size), "greedy", train=True, T=60)
            for k, v in ep.items():
                eps[f"train{j}_{k}"] = v
        np.savez_compressed(os.path.join(OUT, f"episodes_{size}.npz"), **eps)
        ss = run_single_steps(size, 400, 77 + len(size))
        np.savez_compressed(os.path.join(OUT, f"single_steps_{size}.npz"), **ss)
        sv = run_solver_cases(size, 120, 99 + len(size), 0.3)
        np.savez_compressed(os.path.join(OUT, f"solver_{size}.npz"), **sv)
    np.savez_compressed(os.path.join(OUT, "quirks_small.npz"), **run_quirk_scenarios())
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
