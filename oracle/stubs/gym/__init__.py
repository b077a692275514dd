"""Minimal stand-in for the `gym` package (old <=0.21 API) — TEST INFRASTRUCTURE ONLY.

`gym` is not installed in this image. The reference (`/root/reference/warehouse/core.py:5,118-148`,
`baseline/solvers.py:7`) only touches `gym.spaces.{Discrete,Box,MultiBinary,Dict}` and the
`gym.Space` name, so this stub provides exactly those, with `contains`/`sample` semantics of
gym 0.21. It lets `oracle/make_golden.py` and the live-reference tests import the reference
UNMODIFIED. Product code never imports this.
"""
from . import spaces
from .spaces import Space

__all__ = ["spaces", "Space"]
