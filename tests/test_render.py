"""Render bridge (reference: warehouse/core.py:444-617).

CPU: `rllib_warehouse_b200.render.frame_primitives` / `animation_frames` produce, for the states recorded
while the UNMODIFIED reference rendered an episode, exactly the primitives the reference handed its viewer
(`tests/golden/render_*.npz`, made by `oracle/make_golden_render.py`): same order, vertices, radii,
centres and colours, for still frames and for the 10-frame interpolation from the `_prev_*` state.
GPU: `Warehouse.render_state()` — current state plus the `_prev_*` mirrors kept by `wh_save_prev` —
equals the oracle's state before / after every step of an episode, and `render()` draws from it.
"""
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def encode(prims):
    rows = []
    for p in prims:
        if p[0] == "polygon":
            rows.append([0.0] + [float(c) for v in p[1] for c in v] + list(p[2]))
        else:
            rows.append([1.0, float(p[1]), float(p[2][0]), float(p[2][1]), 30.0, 0.0, 0.0, 0.0, 0.0] + list(p[3]))
    return np.array(rows, dtype=np.float64)


@pytest.mark.parametrize("size", ["small", "medium", "large"])
def test_frame_primitives_match_what_the_reference_draws(size):
    from rllib_warehouse_b200 import render as rd
    d = np.load(os.path.join(HERE, "golden", f"render_{size}.npz"))
    dim, racks = int(d["dim"][0]), [int(r) for r in d["racks"]]
    assert rd.viewport_px(dim) == int(d["viewport"][0])                       # core.py:103-105
    n_cases, stills, animated = int(d["n_cases"][0]), 0, 0
    for k in range(n_cases):
        tag = f"c{k}_"
        state = {key: d[tag + key] for key in ("agent_positions", "agent_delivery_targets", "pickup_point_targets",
                                                "prev_agent_positions", "prev_agent_delivery_targets",
                                                "prev_pickup_point_targets")}
        want = [d[tag + f"frame{f}"] for f in range(int(d[tag + "n_frames"][0]))]
        rec = rd.PrimitiveRecorder(rd.viewport_px(dim), rd.viewport_px(dim))
        slept = []
        n = rd.draw(rec, dim, racks, state, bool(d[tag + "animate"][0]), int(d["frames_per_step"][0]), 6.0,
                    sleep=slept.append)
        assert n == len(want) == len(rec.frames)
        for f, (got, ref) in enumerate(zip(rec.frames, want)):
            got = encode(got)
            assert got.shape == ref.shape, (k, f, got.shape, ref.shape)
            assert np.allclose(got, ref, rtol=0, atol=1e-4), (k, f, np.abs(got - ref).max())
        if d[tag + "animate"][0]:
            animated += 1
            assert n == 10 and all(0 < s <= 1.0 / 60.0 for s in slept)          # paced to 6 steps/s x 10 frames
        else:
            stills += 1
    assert stills >= 2 and animated >= 2


def test_static_cell_tables():
    """core.py:171-188 in closed form."""
    from rllib_warehouse_b200 import render as rd
    p = rd.pickup_cells([4, 8])
    assert p.shape == (16, 2) and p[0].tolist() == [3, 3] and p[3].tolist() == [4, 4] and p[4].tolist() == [3, 7]
    dl = rd.delivery_cells(12)
    assert dl.shape == (32, 2) and dl[:4].tolist() == [[2, 0], [0, 2], [2, 11], [11, 2]]


@pytest.mark.gpu
@pytest.mark.parametrize("size,A", [("small", 3), ("large", 16)])
def test_render_state_tracks_the_oracle_through_an_episode(size, A, capsys):
    """`render_state()` == the oracle's state (current AND previous-step mirrors, core.py:270-272) at every
    step of an episode driven through the dict API; `render()` / `render(animate=True)` draw exactly the
    primitives of that state (1 and 10 frames) and print the text frame."""
    import rllib_warehouse_b200 as wh
    from oracle import wh_oracle as wo
    from rllib_warehouse_b200 import render as rd
    cls = {"small": wh.WarehouseSmall, "large": wh.WarehouseLarge}[size]
    env = cls(A, seed=123)
    cpu = wo.OracleEnv(wo.variant_config(size), 1, num_agents=A, seed=123)
    env.reset(); cpu.reset()
    env.render()                                             # turns prev tracking on, like the reference's first frame
    assert isinstance(env._viewer, rd.PrimitiveRecorder) and len(env._viewer.frames) == 1
    rng = np.random.Generator(np.random.PCG64(4))
    cfg = env._config
    for t in range(60):
        before = {k: v.copy() for k, v in cpu.state.items()}
        acts = rng.integers(0, 9, size=A)
        env.step({str(i): int(acts[i]) for i in range(A)})
        a = np.full((1, cpu.R), -1, np.int32); a[0, :A] = acts
        cpu.step(a)
        st = env.render_state()
        assert np.array_equal(st["agent_positions"], cpu.state["agent_pos"][0, :A])
        assert np.array_equal(st["agent_delivery_targets"], cpu.state["agent_tgt"][0, :A])
        assert np.array_equal(st["pickup_point_targets"], cpu.state["pickup_tgt"][0])
        assert np.array_equal(st["pickup_point_timers"], cpu.state["pickup_timer"][0])
        assert st["episode_time"] == int(cpu.state["time"][0]) == t + 1
        assert np.array_equal(st["prev_agent_positions"], before["agent_pos"][0, :A])
        assert np.array_equal(st["prev_agent_delivery_targets"], before["agent_tgt"][0, :A])
        assert np.array_equal(st["prev_pickup_point_targets"], before["pickup_tgt"][0])
        if t % 20 == 7:
            env.render(animate=True)
            frames = env._viewer.frames
            assert len(frames) == env.animate_frames_per_step == 10
            want = [rd.frame_primitives(cfg.area_dimension, cfg.pickup_racks_arrangement, p, tg, pk)
                    for p, tg, pk in rd.animation_frames(st, 10)]
            assert all(np.allclose(encode(g), encode(w)) for g, w in zip(frames, want))
            # frame 0 of the animation still shows the PREVIOUS cell of the last agent (its discs are drawn last)
            assert np.allclose(encode(frames[0])[-1][2:4], (before["agent_pos"][0, A - 1] + 1.5) * 30)
    text = capsys.readouterr().out
    assert "t=0" in text and "t=8" in text and "t=48" in text          # one text frame per render() call
    with pytest.raises(NotImplementedError):
        env.render(mode="rgb_array")
