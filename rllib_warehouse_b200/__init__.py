"""rllib_warehouse_b200 — B200-native batched implementation of the ffahleraz/rllib-warehouse
environment step, observation build and greedy solver (hand-written sm_100a CUDA behind a C ABI).

Reference-compatible surface: `Warehouse`, `WarehouseSmall/Medium/Large`, `Warehouse*Train`
(warehouse/__init__.py:1-11), `WarehouseRandomGreedySolver` (baseline/solvers.py).
Batched surface: `BatchedWarehouse`, `BatchedGreedySolver`.
"""
from .config import LARGE, MEDIUM, SMALL, VARIANTS, WarehouseConfig
from .batched import BatchedWarehouse, StepGraph
from .core import Warehouse
from .variants import (WarehouseLarge, WarehouseLargeTrain, WarehouseMedium, WarehouseMediumTrain,
                       WarehouseSmall, WarehouseSmallTrain)
from .solvers import BatchedGreedySolver, WarehouseRandomGreedySolver, WarehouseSolver
from .vector_env import WarehouseVectorEnv
from .host_env import HostWarehouse
from .sampler import RolloutSampler, mlp_policy

__all__ = [
    "Warehouse", "WarehouseSmall", "WarehouseMedium", "WarehouseLarge",
    "WarehouseSmallTrain", "WarehouseMediumTrain", "WarehouseLargeTrain",
    "WarehouseConfig", "SMALL", "MEDIUM", "LARGE", "VARIANTS",
    "BatchedWarehouse", "StepGraph", "HostWarehouse", "WarehouseVectorEnv", "RolloutSampler", "mlp_policy", "BatchedGreedySolver", "WarehouseRandomGreedySolver", "WarehouseSolver",
]
name = "rllib_warehouse_b200"
