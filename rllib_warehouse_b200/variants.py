"""Variant constructors with the reference's names and signatures (warehouse/variants.py:19-98)."""
from .config import LARGE, MEDIUM, SMALL
from .core import Warehouse

__all__ = [
    "WarehouseSmall", "WarehouseMedium", "WarehouseLarge",
    "WarehouseSmallTrain", "WarehouseMediumTrain", "WarehouseLargeTrain",
]


def _ctor_kwargs(cfg):
    return dict(num_requests=cfg.num_requests, area_dimension=cfg.area_dimension,
                pickup_racks_arrangement=list(cfg.pickup_racks_arrangement),
                episode_duration=cfg.episode_duration, pickup_wait_duration=cfg.pickup_wait_duration)


class _Variant(Warehouse):
    max_num_agents = 0
    _cfg = None

    def __init__(self, num_agents: int, **kw) -> None:
        assert 1 <= num_agents <= self.max_num_agents                         # variants.py:24,39,54
        super().__init__(num_agents=num_agents, **_ctor_kwargs(self._cfg), **kw)


class WarehouseSmall(_Variant):           # variants.py:19-32
    max_num_agents = 4
    _cfg = SMALL


class WarehouseMedium(_Variant):          # variants.py:35-47
    max_num_agents = 9
    _cfg = MEDIUM


class WarehouseLarge(_Variant):           # variants.py:50-62
    max_num_agents = 16
    _cfg = LARGE


class _TrainMixin:
    """variants.py:65-98: num_agents ~ U{1..max} at construction and again on every reset(). The
    redraw happens on the device (wh_reset with random_num_agents), keyed by (seed, env, episode)."""

    def __init__(self, **kw) -> None:
        import numpy as np
        Warehouse.__init__(self, num_agents=int(np.random.randint(1, self.max_num_agents + 1)),
                           random_num_agents=True,
                           max_num_agents=self.max_num_agents, **_ctor_kwargs(self._cfg), **kw)


class WarehouseSmallTrain(_TrainMixin, WarehouseSmall):
    pass


class WarehouseMediumTrain(_TrainMixin, WarehouseMedium):
    pass


class WarehouseLargeTrain(_TrainMixin, WarehouseLarge):
    pass
