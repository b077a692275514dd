"""`from solvers import WarehouseRandomGreedySolver` (reference baseline/run.py:12) resolves here:
the same solver interface, evaluated by the batched warp-argmin CUDA kernel."""
from rllib_warehouse_b200.solvers import WarehouseRandomGreedySolver, WarehouseSolver

__all__ = ["WarehouseSolver", "WarehouseRandomGreedySolver"]
