"""Import-path shim: `from warehouse import WarehouseSmall, ...` (as the reference's drivers do,
baseline/run.py:6-10, scripts/train.py:11-15, scripts/rollout.py:13-17) resolves to the
B200-native implementation in `rllib_warehouse_b200`."""
from rllib_warehouse_b200.core import Warehouse
from rllib_warehouse_b200.variants import (
    WarehouseLarge, WarehouseLargeTrain, WarehouseMedium, WarehouseMediumTrain, WarehouseSmall,
    WarehouseSmallTrain,
)

__all__ = [
    "Warehouse", "WarehouseSmall", "WarehouseMedium", "WarehouseLarge",
    "WarehouseSmallTrain", "WarehouseMediumTrain", "WarehouseLargeTrain",
]
name = "warehouse"
