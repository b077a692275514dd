"""HostWarehouse — the host-buffer layer of the C ABI (`wh_env_*`, include/wh_b200.h) for callers
that live on the CPU (numpy policies, RLlib samplers): actions come from host memory and
rewards / dones go back to host memory on every step, pipelined over env chunks on several
streams, while state and observations stay resident in HBM. Observation tensors can still be
reached zero-copy as torch CUDA tensors (`obs_tensors()`), or copied out per step (`with_obs=True`).
"""
import ctypes as C

import numpy as np
import torch

from . import _native as nv
from .config import WarehouseConfig

__all__ = ["HostWarehouse"]


class _CudaView:
    """Minimal __cuda_array_interface__ holder so torch can wrap a raw device pointer."""

    def __init__(self, ptr, shape, typestr, owner):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2}
        self._owner = owner


class HostWarehouse:
    def __init__(self, config: WarehouseConfig, num_envs: int, device: int = 0, seed: int = 0,
                 env_id0: int = 0, chunks: int = 0, compact: bool = False):
        if not torch.cuda.is_available():
            raise nv.NativeError("HostWarehouse needs a CUDA device: there is no CPU fallback")
        self.lib, self.config = nv.lib(), config
        self.N, self.R, self.device, self.compact = int(num_envs), config.num_requests, int(device), bool(compact)
        self._cfg = nv.make_config(config)
        self._h = C.c_void_p()
        nv.check(self.lib.wh_env_create(C.byref(self._cfg), self.N, self.device, int(env_id0),
                                        int(seed) & (2**64 - 1), int(chunks), C.byref(self._h)), "wh_env_create")
        adt, rdt = (torch.int8, torch.uint8) if compact else (torch.int32, torch.float32)
        # pinned staging owned by this object; step() accepts any array-like and copies into it
        self._actions = torch.zeros((self.N, self.R), dtype=adt).pin_memory()
        self._rewards = torch.zeros((self.N, self.R), dtype=rdt).pin_memory()
        self._dones = torch.zeros(self.N, dtype=torch.uint8).pin_memory()
        self._obs_host = None

    def close(self):
        if self._h:
            self.lib.wh_env_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    def reset(self):
        nv.check(self.lib.wh_env_reset(self._h), "wh_env_reset")

    def step(self, actions, with_obs: bool = False):
        """actions [N,R] integers (-1 = agent absent). Returns (rewards [N,R] float32 numpy view,
        dones [N] bool numpy, obs dict of numpy arrays or None). Finished envs auto-reset."""
        self._actions.copy_(torch.as_tensor(np.asarray(actions)).reshape(self.N, self.R))
        if self.compact:
            rc = self.lib.wh_env_step_host_compact(self._h, self._actions.data_ptr(), self._rewards.data_ptr(),
                                                   self._dones.data_ptr())
            obs = None
            assert not with_obs, "use obs_tensors() with the compact wire format"
        else:
            oh = None
            if with_obs:
                if self._obs_host is None:
                    self._obs_host = {k: torch.zeros(v.shape, dtype=v.dtype).pin_memory()
                                      for k, v in self.obs_tensors().items()}
                oh = nv.Obs(**{k: v.data_ptr() for k, v in self._obs_host.items()})
            rc = self.lib.wh_env_step_host(self._h, self._actions.data_ptr(), self._rewards.data_ptr(),
                                           self._dones.data_ptr(), C.byref(oh) if oh is not None else None)
            obs = {k: v.numpy() for k, v in self._obs_host.items()} if with_obs else None
        nv.check(rc, "wh_env_step_host")
        return self._rewards.numpy().astype(np.float32, copy=False), self._dones.numpy().astype(bool), obs

    def greedy_step(self):
        """The greedy solver runs on the device; only rewards and dones cross PCIe."""
        assert not self.compact
        nv.check(self.lib.wh_env_greedy_step_host(self._h, self._rewards.data_ptr(), self._dones.data_ptr()),
                 "wh_env_greedy_step_host")
        return self._rewards.numpy(), self._dones.numpy().astype(bool)

    def obs_tensors(self):
        """Zero-copy torch views of the resident observation tensors (wh_obs layout)."""
        ob = nv.Obs()
        nv.check(self.lib.wh_env_obs_ptrs(self._h, C.byref(ob)), "wh_env_obs_ptrs")
        N, R = self.N, self.R
        shapes = {"num_agents": ((N, R, 1), "<i4"), "self_position": ((N, R, 2), "<i4"),
                  "self_availability": ((N, R, 1), "|i1"), "self_delivery_target": ((N, R, 2), "<i4"),
                  "other_positions": ((N, R, R - 1, 2), "<i4"), "other_availabilities": ((N, R, R - 1), "|i1"),
                  "other_delivery_targets": ((N, R, R - 1, 2), "<i4"), "requests": ((N, R, R, 4), "<i4")}
        with torch.cuda.device(self.device):
            return {k: torch.as_tensor(_CudaView(getattr(ob, k), shp, ts, self), device=f"cuda:{self.device}")
                    for k, (shp, ts) in shapes.items()}

    def stats(self):
        buf = (C.c_ulonglong * nv.NUM_STATS)()
        nv.check(self.lib.wh_env_stats_host(self._h, buf), "wh_env_stats_host")
        return np.array(list(buf), dtype=np.int64)
