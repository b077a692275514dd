"""Geometry/constants of a warehouse variant (reference: warehouse/core.py:78-108, variants.py:19-62)."""
from dataclasses import dataclass

MAX_RACKS = 8


@dataclass(frozen=True)
class WarehouseConfig:
    num_requests: int
    area_dimension: int
    pickup_racks_arrangement: tuple
    episode_duration: int = 200
    pickup_wait_duration: int = 200
    max_num_agents: int = 0          # 0 -> num_requests
    random_num_agents: bool = False  # *Train variants (variants.py:65-98)

    def __post_init__(self):
        object.__setattr__(self, "pickup_racks_arrangement", tuple(int(r) for r in self.pickup_racks_arrangement))
        if self.max_num_agents == 0:
            object.__setattr__(self, "max_num_agents", self.num_requests)

    @property
    def num_pickup_points(self) -> int:      # core.py:96
        return 4 * len(self.pickup_racks_arrangement) ** 2

    @property
    def num_delivery_points(self) -> int:    # core.py:97
        return 4 * (self.area_dimension - 4)

    @property
    def null_position(self) -> int:          # core.py:107
        return self.area_dimension // 2

    def replace(self, **kw):
        d = dict(num_requests=self.num_requests, area_dimension=self.area_dimension,
                 pickup_racks_arrangement=self.pickup_racks_arrangement,
                 episode_duration=self.episode_duration, pickup_wait_duration=self.pickup_wait_duration,
                 max_num_agents=self.max_num_agents, random_num_agents=self.random_num_agents)
        d.update(kw)
        return WarehouseConfig(**d)


# variants.py:25-32 / 40-47 / 55-62
SMALL = WarehouseConfig(4, 12, (4, 8), 200, 200, 4)
MEDIUM = WarehouseConfig(9, 16, (4, 8, 12), 200, 200, 9)
LARGE = WarehouseConfig(16, 20, (4, 8, 12, 16), 200, 200, 16)
VARIANTS = {"small": SMALL, "medium": MEDIUM, "large": LARGE}
