#!/usr/bin/env python
"""Golden vectors for the render bridge — TEST INFRASTRUCTURE ONLY.

Runs the UNMODIFIED reference `Warehouse.render` (warehouse/core.py:444-617) with a recording stand-in for
`gym.envs.classic_control.rendering` (pyglet is absent here), so that every primitive the reference hands
its viewer — polygons with vertices and colour, circles with radius, centre and colour, per frame — is
captured together with the env state it was drawn from (current state and the `_prev_*` mirrors). The
fixtures (`tests/golden/render_{small,medium,large}.npz`) pin `rllib_warehouse_b200/render.py`.

    python oracle/make_golden_render.py        # needs /root/reference

Frame encoding: float64 [n_primitives, 12] rows — polygon: [0, x0,y0,x1,y1,x2,y2,x3,y3, r,g,b];
circle: [1, radius, cx, cy, resolution, 0,0,0,0, r,g,b].
"""
import contextlib
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402  (imports the reference under the stub gym / ray)

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
FRAMES = []


class _Geom:
    def __init__(self, row):
        self.row = row

    def add_attr(self, attr):
        self.row[2:4] = [float(attr.translation[0]), float(attr.translation[1])]


class _Transform:
    def __init__(self, translation=(0.0, 0.0)):
        self.translation = translation


class _Viewer:
    def __init__(self, width, height):
        self.size, self.rows = (width, height), []

    def draw_polygon(self, v, color=(0, 0, 0)):
        assert len(v) == 4
        self.rows.append([0.0] + [float(c) for p in v for c in p] + [float(c) for c in color])

    def draw_circle(self, radius, res, color=(0, 0, 0)):
        row = [1.0, float(radius), 0.0, 0.0, float(res), 0.0, 0.0, 0.0, 0.0] + [float(c) for c in color]
        self.rows.append(row)
        return _Geom(row)

    def render(self):
        FRAMES.append(np.array(self.rows, dtype=np.float64))
        self.rows = []


@contextlib.contextmanager
def recording_gym():
    """`from gym.envs.classic_control import rendering` inside core.py:478 resolves to the recorder."""
    mods = {n: types.ModuleType(n) for n in ("gym", "gym.envs", "gym.envs.classic_control",
                                             "gym.envs.classic_control.rendering")}
    mods["gym"].envs = mods["gym.envs"]
    mods["gym.envs"].classic_control = mods["gym.envs.classic_control"]
    mods["gym.envs.classic_control"].rendering = mods["gym.envs.classic_control.rendering"]
    mods["gym.envs.classic_control.rendering"].Viewer = _Viewer
    mods["gym.envs.classic_control.rendering"].Transform = _Transform
    saved = {n: sys.modules.get(n) for n in mods}
    sys.modules.update(mods)
    try:
        yield
    finally:
        for n, m in saved.items():
            if m is None:
                sys.modules.pop(n, None)
            else:
                sys.modules[n] = m


def record(size, A, seed, render_steps):
    import time
    np.random.seed(seed)
    env = mg.VARIANTS[size][0](A)
    solver = mg.WarehouseRandomGreedySolver(env.num_agents, env.num_requests, 0.1, env.action_space)
    obs = env.reset()
    out = {"viewport": np.array([env._viewport_dimension_PX]), "dim": np.array([env._area_dimension]),
           "racks": np.array(env._pickup_racks_arrangement), "frames_per_step": np.array([env.animate_frames_per_step])}
    sleep, time.sleep = time.sleep, (lambda s: None)           # core.py:469: do not actually pace the animation
    try:
        with recording_gym():
            k = 0
            for t in range(max(render_steps) + 1):
                if t in render_steps:
                    for animate in (False, True):
                        del FRAMES[:]
                        env.render(animate=animate)
                        tag = f"c{k}_"
                        out[tag + "animate"] = np.array([int(animate)])
                        out[tag + "agent_positions"] = env._agent_positions.copy()
                        out[tag + "agent_delivery_targets"] = env._agent_delivery_targets.copy()
                        out[tag + "pickup_point_targets"] = env._pickup_point_targets.copy()
                        out[tag + "prev_agent_positions"] = env._prev_agent_positions.copy()
                        out[tag + "prev_agent_delivery_targets"] = env._prev_agent_delivery_targets.copy()
                        out[tag + "prev_pickup_point_targets"] = env._prev_pickup_point_targets.copy()
                        out[tag + "n_frames"] = np.array([len(FRAMES)])
                        for f, rows in enumerate(FRAMES):
                            out[tag + f"frame{f}"] = rows
                        k += 1
                obs, _, _, _ = env.step(solver.compute_action(obs))
            out["n_cases"] = np.array([k])
    finally:
        time.sleep = sleep
    return out


if __name__ == "__main__":
    for size, A, seed, steps in (("small", 3, 5, (0, 1, 17, 40)), ("medium", 9, 6, (0, 9, 33)), ("large", 16, 7, (0, 25))):
        d = record(size, A, seed, steps)
        path = os.path.join(OUT, f"render_{size}.npz")
        np.savez_compressed(path, **d)
        print(path, int(d["n_cases"][0]), "render calls,", os.path.getsize(path), "bytes")
