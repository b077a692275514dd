"""numpy restatement of the reference hot path — TEST INFRASTRUCTURE ONLY.

A second, independent CPU oracle (the first is the C port, oracle/wh_oracle.c): it follows
`warehouse/core.py:167-442` and `baseline/solvers.py:27-58` at the granularity the reference
itself works at — one environment object, numpy arrays per step, Python loop over agents — so it
also stands in for the reference's per-step cost class in bench.py's Python-level CPU number
(the reference package itself cannot travel to the GPU box).

It is pinned against the same fixtures recorded from the unmodified reference
(tests/test_port_golden.py). RNG: the reference's global-stream draws are REPLAYED through the
arguments; when they are omitted a local `np.random.Generator` supplies
distribution-equivalent draws (used only for timing).

Only tests/ and bench.py's cpu legs may import this module.
"""
import numpy as np

OBS_KEYS = (
    "num_agents", "self_position", "self_availability", "self_delivery_target",
    "other_positions", "other_availabilities", "other_delivery_targets", "requests",
)
STATE_KEYS = ("agent_pos", "agent_tgt", "pickup_tgt", "pickup_timer", "time", "num_agents")


class PortWarehouse:
    """One environment (core.py:73-442)."""

    def __init__(self, num_agents, num_requests, area_dimension, racks, episode=200, wait=200, rng=None):
        assert num_agents <= num_requests                                      # core.py:89
        self.A, self.R, self.dim = int(num_agents), int(num_requests), int(area_dimension)
        self.racks = [int(r) for r in racks]
        self.P = 4 * len(self.racks) ** 2                                      # core.py:96
        self.D = 4 * (self.dim - 4)                                            # core.py:97
        self.episode, self.wait = int(episode), int(wait)
        self.null = self.dim // 2                                              # core.py:107
        self.rng = rng or np.random.default_rng(0)
        # core.py:171-175
        self.pickup_cells = np.array(
            [c for x in self.racks for y in self.racks for c in ((x - 1, y - 1), (x, y - 1), (x - 1, y), (x, y))],
            dtype=np.int32)
        # core.py:178-188
        self.delivery_cells = np.array(
            [c for v in range(2, self.dim - 2) for c in ((v, 0), (0, v), (v, self.dim - 1), (self.dim - 1, v))],
            dtype=np.int32)
        self._pickup_lookup = {}
        for p, c in enumerate(map(tuple, self.pickup_cells.tolist())):
            self._pickup_lookup.setdefault(c, p)                               # first match (argmax)
        self.pos = np.zeros((self.A, 2), np.int32)
        self.tgt = np.full(self.A, -1, np.int32)
        self.ptgt = np.full(self.P, -1, np.int32)
        self.ptim = np.full(self.P, -1, np.int32)
        self.time = 0

    # -- core.py:167-221 ------------------------------------------------------------------------
    def reset(self, agent_pos=None, init_pickups=None, init_targets=None):
        self.time = 0
        if agent_pos is None:
            cells = []
            while len(cells) < self.A:                                         # core.py:192-201
                c = (int(self.rng.integers(1, self.dim - 1)), int(self.rng.integers(1, self.dim - 1)))
                if c not in self._pickup_lookup:
                    cells.append(c)
            agent_pos = cells
            init_pickups = self.rng.permutation(self.P)[: self.R]              # core.py:215-220
            init_targets = self.rng.permutation(self.D)[: self.R]
        self.pos = np.array(agent_pos, np.int32).reshape(-1, 2)[: self.A].copy()
        self.tgt = np.full(self.A, -1, np.int32)
        self.ptgt = np.full(self.P, -1, np.int32)
        self.ptim = np.full(self.P, -1, np.int32)
        ip = np.asarray(init_pickups).reshape(-1)
        it = np.asarray(init_targets).reshape(-1)
        ok = ip >= 0
        self.ptgt[ip[ok]] = it[ok]
        self.ptim[ip[ok]] = self.wait                                          # core.py:221
        return self.observations(reset_flavour=True)

    # -- core.py:262-368, 435-440 ---------------------------------------------------------------
    def step(self, action_items, spawn_pickups=None, spawn_targets=None):
        """action_items: iterable of (agent_index, action) in action-dict iteration order."""
        self.time += 1
        dim = self.dim
        occ = np.zeros((dim, dim), bool)                                       # core.py:275-276
        occ[self.pos[:, 0], self.pos[:, 1]] = True
        forbidden = set()
        for idx, action in action_items:                                       # core.py:279-300
            px, py = int(self.pos[idx, 0]), int(self.pos[idx, 1])
            x, y = px + action // 3 - 1, py + action % 3 - 1
            if not 0 <= x < dim:
                x = px
            if not 0 <= y < dim:
                y = py
            if occ[x, y] or (px, py, x, y) in forbidden:
                continue
            occ[px, py], occ[x, y] = False, True
            forbidden.add((x, y, px, py))
            if x != px and y != py:
                forbidden.add((x, py, px, y))
                forbidden.add((px, y, x, py))
            self.pos[idx] = (x, y)
        # core.py:303-306
        self.ptim[self.ptgt > -1] -= 1
        expired = self.ptim == 0
        self.ptgt[expired] = -1
        self.ptim[expired] = -1
        # core.py:309-335
        cand = np.array([self._pickup_lookup.get((int(x), int(y)), -1) for x, y in self.pos], np.int64)
        picks = (cand >= 0) & (self.tgt == -1) & (self.ptgt[np.maximum(cand, 0)] > -1)
        served = cand[picks]
        self.tgt[picks] = self.ptgt[served]
        self.ptgt[served] = -1
        self.ptim[served] = -1
        rewards = np.zeros(self.A, np.float32)
        rewards[picks] += 1.0
        # core.py:338-351
        inactive = np.nonzero(self.ptgt == -1)[0]
        k = self.R - self.P + len(inactive)
        if spawn_pickups is None:
            sp = self.rng.permutation(inactive)[:k]
            st = self.rng.permutation(self.D)[:k]
        else:
            sp = np.asarray(spawn_pickups).reshape(-1)
            st = np.asarray(spawn_targets).reshape(-1)
            keep = sp >= 0
            sp, st = sp[keep], st[keep]
        self.ptim[sp] = self.wait
        self.ptgt[sp] = st
        # core.py:354-368
        busy = np.nonzero(self.tgt > -1)[0]
        arrived = busy[(self.delivery_cells[self.tgt[busy]] == self.pos[busy]).all(axis=1)]
        self.tgt[arrived] = -1
        rewards[arrived] += 1.0
        return self.observations(reset_flavour=False), rewards, self.time >= self.episode

    # -- core.py:224-260 / 371-432 --------------------------------------------------------------
    def observations(self, reset_flavour):
        A, R, null = self.A, self.R, self.null
        ppos = np.full((R, 2), null, np.int32)
        ppos[:A] = self.pos
        avail = np.zeros(R, np.int8)
        tpos = np.full((R, 2), null, np.int32)
        if not reset_flavour:                                                  # core.py:233-236 vs 381-407
            busy = self.tgt > -1
            avail[:A][~busy] = 1
            tpos[:A][busy] = self.delivery_cells[self.tgt[busy]]
        waiting = self.ptgt > -1                                               # core.py:409-418
        requests = np.hstack((self.pickup_cells[waiting], self.delivery_cells[self.ptgt[waiting]])).astype(np.int32)
        num_agents = np.full(1, A, np.int32)
        out = {}
        for i in range(A):
            rows = np.arange(R) != i
            trows = rows if reset_flavour else (np.arange(R) != 1)             # core.py:256 vs core.py:428
            out[i] = {
                "num_agents": num_agents, "self_position": ppos[i], "self_availability": avail[i:i + 1],
                "self_delivery_target": tpos[i], "other_positions": ppos[rows],
                "other_availabilities": avail[rows], "other_delivery_targets": tpos[trows],
                "requests": requests,
            }
        return out


def greedy_actions(observations, num_agents, num_requests, is_random=None, random_actions=None):
    """solvers.py:27-58 for one env; observations[i] are per-agent dicts."""
    acts = np.full(num_agents, -1, np.int32)
    for i in range(num_agents):
        o = observations[i]
        me = o["self_position"].astype(np.int64)
        if o["self_availability"][0] == 0:
            target = o["self_delivery_target"].astype(np.int64)
        else:
            req = o["requests"][:num_requests, :2].astype(np.int64)
            target = req[np.argmin(np.abs(req - me).sum(axis=1))]             # first minimum
        step = np.clip(target - me, -1, 1)
        acts[i] = (step[0] + 1) * 3 + (step[1] + 1)
        if is_random is not None and is_random[i]:
            acts[i] = random_actions[i]
    return acts


class PortEnv:
    """N PortWarehouse objects behind the batch interface the golden checkers drive."""

    def __init__(self, kw, n, num_agents):
        self.kw, self.N = kw, n
        self.R = kw["num_requests"]
        self.envs = [self._make(num_agents or self.R) for _ in range(n)]
        self._sync()

    def _make(self, A):
        k = self.kw
        return PortWarehouse(A, k["num_requests"], k["area_dimension"], k["racks"], k["episode"], k["wait"])

    def _sync(self, obs=None):
        N, R = self.N, self.R
        P = self.envs[0].P
        st = dict(agent_pos=np.full((N, R, 2), -1, np.int32), agent_tgt=np.full((N, R), -1, np.int32),
                  pickup_tgt=np.zeros((N, P), np.int32), pickup_timer=np.zeros((N, P), np.int32),
                  time=np.zeros(N, np.int32), num_agents=np.zeros(N, np.int32))
        for e, w in enumerate(self.envs):
            st["agent_pos"][e, : w.A] = w.pos
            st["agent_tgt"][e, : w.A] = w.tgt
            st["pickup_tgt"][e], st["pickup_timer"][e] = w.ptgt, w.ptim
            st["time"][e], st["num_agents"][e] = w.time, w.A
        self.state = st
        if obs is not None:
            shapes = {"num_agents": (1,), "self_position": (2,), "self_availability": (1,),
                      "self_delivery_target": (2,), "other_positions": (R - 1, 2),
                      "other_availabilities": (R - 1,), "other_delivery_targets": (R - 1, 2), "requests": (R, 4)}
            self.obs = {k: np.full((N, R) + shapes[k], -1, np.int32) for k in OBS_KEYS}
            for e, od in enumerate(obs):
                for i, o in od.items():
                    for k in OBS_KEYS:
                        self.obs[k][e, i] = o[k]
        return self.obs if obs is not None else None

    def load_state(self, **a):
        n_ag = np.asarray(a["num_agents"]).reshape(-1)
        self.envs = [self._make(int(n_ag[e])) for e in range(self.N)]
        for e, w in enumerate(self.envs):
            w.pos = np.asarray(a["agent_pos"][e][: w.A], np.int32).copy()
            w.tgt = np.asarray(a["agent_tgt"][e][: w.A], np.int32).copy()
            w.ptgt = np.asarray(a["pickup_tgt"][e], np.int32).copy()
            w.ptim = np.asarray(a["pickup_timer"][e], np.int32).copy()
            w.time = int(np.asarray(a["time"]).reshape(-1)[e])
        self._sync()

    def reset(self, agent_pos=None, init_pickups=None, init_targets=None, num_agents=None):
        if num_agents is not None:
            self.envs = [self._make(int(np.asarray(num_agents).reshape(-1)[e])) for e in range(self.N)]
        obs = [w.reset(None if agent_pos is None else agent_pos[e], None if init_pickups is None else init_pickups[e],
                       None if init_targets is None else init_targets[e]) for e, w in enumerate(self.envs)]
        return self._sync(obs)

    def step(self, actions, order=None, spawn_pickups=None, spawn_targets=None):
        obs, rew, dones = [], np.zeros((self.N, self.R), np.float32), np.zeros(self.N, np.uint8)
        for e, w in enumerate(self.envs):
            ids = [int(i) for i in (order[e] if order is not None else range(w.A)) if 0 <= int(i) < w.A]
            items = [(i, int(actions[e][i])) for i in ids if int(actions[e][i]) >= 0]
            o, r, d = w.step(items, None if spawn_pickups is None else spawn_pickups[e],
                             None if spawn_targets is None else spawn_targets[e])
            obs.append(o)
            rew[e, : w.A] = r
            dones[e] = d
        return self._sync(obs), rew, dones
