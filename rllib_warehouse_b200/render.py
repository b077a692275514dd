"""Render bridge — draws one environment's state the way `Warehouse.render` does
(warehouse/core.py:444-617), from state copied off the GPU.

The reference draws with gym's pyglet viewer (`gym.envs.classic_control.rendering.Viewer`): a grey frame,
a white floor, one square per pickup point (blue while a request waits there), a double square per
delivery point (blue rim while some agent carries an item for it), three concentric discs per agent
(the innermost, blue one only while it carries an item), and — with `animate=True` — 10 interpolated
frames from the previous step's positions, drawn with the PREVIOUS step's targets (core.py:448-470).

Here the frame is first built as a plain list of primitives (`frame_primitives`, a pure function of the
state arrays — this is what the parity test compares with the primitives the reference hands its viewer),
then sent to whatever viewer is available: gym's, if `gym.envs.classic_control.rendering` imports, else a
`PrimitiveRecorder` that keeps the last frames (and the caller prints the text frame).
"""
import time
from typing import List, Sequence, Tuple

import numpy as np

# core.py:50-67
AGENT_RADIUS = 0.38
BORDER_WIDTH = 1.0
PIXELS_PER_METER = 30
GREY_DARK, GREY_LIGHT, BLUE, WHITE = (0.5, 0.5, 0.5), (0.8, 0.8, 0.8), (0.0, 0.0, 1.0), (1.0, 1.0, 1.0)
CIRCLE_RESOLUTION = 30

Polygon = Tuple[str, List[Tuple[float, float]], Tuple[float, float, float]]
Circle = Tuple[str, float, Tuple[float, float], Tuple[float, float, float]]


def viewport_px(area_dimension: int) -> int:
    return int(area_dimension + 2 * BORDER_WIDTH) * PIXELS_PER_METER          # core.py:103-105


def pickup_cells(racks: Sequence[int]) -> np.ndarray:
    """core.py:171-175: for x in racks, for y in racks: (x-1,y-1) (x,y-1) (x-1,y) (x,y)."""
    return np.array([(x + ox, y + oy) for x in racks for y in racks
                     for ox, oy in ((-1, -1), (0, -1), (-1, 0), (0, 0))], dtype=np.int32)


def delivery_cells(dim: int) -> np.ndarray:
    """core.py:178-188: for v in 2..dim-3: (v,0) (0,v) (v,dim-1) (dim-1,v)."""
    return np.array([c for v in range(2, dim - 2) for c in ((v, 0), (0, v), (v, dim - 1), (dim - 1, v))],
                    dtype=np.int32)


def _square(x: float, y: float, inset: float):
    """Axis-aligned square of the unit cell at (x, y), shrunk by `inset` on every side, in pixels."""
    lo_x, lo_y = (x + BORDER_WIDTH + inset) * PIXELS_PER_METER, (y + BORDER_WIDTH + inset) * PIXELS_PER_METER
    hi_x, hi_y = (x + BORDER_WIDTH + 1.0 - inset) * PIXELS_PER_METER, (y + BORDER_WIDTH + 1.0 - inset) * PIXELS_PER_METER
    return [(lo_x, lo_y), (hi_x, lo_y), (hi_x, hi_y), (lo_x, hi_y)]


def frame_primitives(dim: int, racks: Sequence[int], agent_positions, agent_delivery_targets,
                     pickup_point_targets) -> list:
    """One frame as an ordered list of ("polygon", vertices, colour) / ("circle", radius, centre, colour),
    in the reference's drawing order (core.py:486-615)."""
    out: list = []
    full = (dim + 2 * BORDER_WIDTH) * PIXELS_PER_METER
    out.append(("polygon", [(0.0, 0.0), (full, 0.0), (full, full), (0.0, full)], GREY_DARK))       # frame
    lo, hi = BORDER_WIDTH * PIXELS_PER_METER, (dim + BORDER_WIDTH) * PIXELS_PER_METER
    out.append(("polygon", [(lo, lo), (hi, lo), (hi, hi), (lo, hi)], WHITE))                       # floor
    for p, (x, y) in enumerate(pickup_cells(racks)):                                                # core.py:518-542
        out.append(("polygon", _square(x, y, 0.1), BLUE if pickup_point_targets[p] > -1 else GREY_LIGHT))
    carried = set(int(t) for t in np.asarray(agent_delivery_targets).ravel())
    for d, (x, y) in enumerate(delivery_cells(dim)):                                                # core.py:545-592
        out.append(("polygon", _square(x, y, 0.1), BLUE if d in carried else GREY_LIGHT))
        out.append(("polygon", _square(x, y, 0.2), GREY_LIGHT))
    for pos, tgt in zip(np.asarray(agent_positions, dtype=np.float32), agent_delivery_targets):    # core.py:595-613
        centre = tuple(float(v) for v in (pos + np.float32(0.5) + BORDER_WIDTH) * PIXELS_PER_METER)
        out.append(("circle", AGENT_RADIUS * PIXELS_PER_METER, centre, GREY_DARK))
        out.append(("circle", AGENT_RADIUS * 3 / 4 * PIXELS_PER_METER, centre, GREY_LIGHT))
        if tgt > -1:
            out.append(("circle", AGENT_RADIUS / 2 * PIXELS_PER_METER, centre, BLUE))
    return out


def animation_frames(state: dict, frames: int):
    """core.py:448-462: `frames` positions from the previous step's cells towards the current ones
    (prev + (cur - prev) / frames * i, i = 0..frames-1), each drawn with the PREVIOUS targets."""
    prev, cur = state["prev_agent_positions"], state["agent_positions"]
    for i in range(frames):
        yield (prev + (cur - prev) / frames * i, state["prev_agent_delivery_targets"], state["prev_pickup_point_targets"])


class PrimitiveRecorder:
    """Viewer stand-in when gym's pyglet viewer is not importable: keeps the primitives of the frames drawn
    since the last `clear()` (the reference's viewer would rasterise them)."""

    def __init__(self, width: int, height: int):
        self.width, self.height, self.frames = width, height, []

    def show(self, primitives: list) -> None:
        self.frames.append(primitives)

    def clear(self) -> None:
        self.frames = []

    def close(self) -> None:
        self.frames = []


class GymViewer:
    """The reference's viewer (gym.envs.classic_control.rendering.Viewer, core.py:480-485)."""

    def __init__(self, width: int, height: int, rendering):
        self._rendering = rendering
        self._viewer = rendering.Viewer(width, height)

    def show(self, primitives: list) -> None:
        for prim in primitives:
            if prim[0] == "polygon":
                self._viewer.draw_polygon(prim[1], color=prim[2])
            else:
                _, radius, centre, colour = prim
                self._viewer.draw_circle(radius, CIRCLE_RESOLUTION, color=colour).add_attr(
                    self._rendering.Transform(translation=centre))
        self._viewer.render()

    def close(self) -> None:
        self._viewer.close()


def make_viewer(area_dimension: int):
    px = viewport_px(area_dimension)
    try:  # pragma: no cover - gym / pyglet are absent in the build image
        from gym.envs.classic_control import rendering  # type: ignore
        return GymViewer(px, px, rendering)
    except Exception:  # noqa: BLE001
        return PrimitiveRecorder(px, px)


def draw(viewer, dim: int, racks: Sequence[int], state: dict, animate: bool, frames: int,
         steps_per_second: float, sleep=time.sleep) -> int:
    """`Warehouse.render` body (core.py:444-475). Returns the number of frames drawn."""
    if not animate:
        viewer.show(frame_primitives(dim, racks, state["agent_positions"], state["agent_delivery_targets"],
                                     state["pickup_point_targets"]))
        return 1
    budget = 1.0 / (steps_per_second * frames)                                # core.py:464-469: pace the animation
    for pos, tgt, pickups in animation_frames(state, frames):
        t0 = time.time()
        viewer.show(frame_primitives(dim, racks, pos, tgt, pickups))
        spent = time.time() - t0
        if spent < budget:
            sleep(budget - spent)
    return frames
