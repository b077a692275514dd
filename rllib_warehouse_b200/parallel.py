"""Multi-GPU plumbing: envs shard trivially (no per-step communication). One process per GPU;
the only collective is the end-of-rollout all-reduce of the int64 episode-statistics vector
(NCCL over NVLink on GPUs; gloo in the CPU tests). Rewards are integer-valued, so integer sums
are exact and independent of reduction order."""
import os

import torch
import torch.distributed as dist

from . import _native as nv


def shard_range(num_envs_total: int, rank: int, world: int):
    """Contiguous global env-id range [lo, hi) of `rank`. The RNG is keyed by the GLOBAL env id
    (wh_config / env_id0), so per-env results do not depend on `world`."""
    lo = num_envs_total * rank // world
    hi = num_envs_total * (rank + 1) // world
    return lo, hi


def bind_to_gpu_numa_node(device_index: int):
    """Pins this process to the CPUs of the NUMA node its GPU hangs off (sysfs: the PCI device's `numa_node`
    and the node's `cpulist`), so that page-locked host buffers allocated afterwards are local to the GPU's
    PCIe root. torchrun does not do this; without it half the ranks of a two-socket box reach their host
    buffers across the socket interconnect. Returns the node id, or None when it cannot be determined."""
    try:
        bus = torch.cuda.get_device_properties(device_index).pci_bus_id
        dom = torch.cuda.get_device_properties(device_index).pci_domain_id
        dev = torch.cuda.get_device_properties(device_index).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node"
        node = int(open(path).read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except Exception:  # noqa: BLE001
        return None


def init_from_env(backend=None, device=None):
    """Reads RANK / WORLD_SIZE / MASTER_* (torchrun contract). Returns (rank, world)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {"device_id": device} if (backend == "nccl" and device is not None) else {}
        dist.init_process_group(backend, **kw)
    return rank, world


def allreduce_stats(stats: torch.Tensor, group=None) -> torch.Tensor:
    """Sum of the per-shard statistics vectors (layout: include/wh_b200.h). Returns a new tensor."""
    assert stats.dtype == torch.int64 and stats.numel() == nv.NUM_STATS
    out = stats.clone()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
    return out


def stats_to_metrics(stats, max_agents: int):
    """The reference's episode metrics (scripts/train.py:18-23): avg_agent_reward_all and
    avg_agent_reward_{n} from the reduced vector."""
    s = [int(v) for v in stats.tolist()]
    out = dict(episodes=s[0], return_sum=s[1], pickups=s[2], deliveries=s[3], expired=s[4])
    num = 0.0
    for n in range(1, max_agents + 1):
        ep, ret = s[8 + 2 * (n - 1)], s[9 + 2 * (n - 1)]
        if ep:
            out[f"avg_agent_reward_{n}"] = ret / n / ep
            num += ret / n
    if s[0]:
        out["avg_agent_reward_all"] = num / s[0]
    return out


class RawNcclStats:
    """The same reduction through the C ABI (`wh_stats_allreduce`) on a raw `ncclComm_t`, i.e. what
    a non-PyTorch host program would do. The communicator is created here with the NCCL that
    PyTorch already loaded (unique id broadcast over the existing torch.distributed group)."""

    def __init__(self, device):
        import ctypes as C
        self.C, self.device = C, torch.device(device)
        self.nccl = C.CDLL("libnccl.so.2")
        rank, world = dist.get_rank(), dist.get_world_size()

        class UniqueId(C.Structure):
            _fields_ = [("internal", C.c_byte * 128)]

        uid = UniqueId()
        if rank == 0:
            rc = self.nccl.ncclGetUniqueId(C.byref(uid))
            assert rc == 0, rc
        t = torch.tensor(list(bytes(uid)), dtype=torch.uint8, device=self.device)
        dist.broadcast(t, src=0)
        C.memmove(C.byref(uid), bytes(t.cpu().tolist()), 128)
        self.comm = C.c_void_p()
        self.nccl.ncclCommInitRank.argtypes = [C.c_void_p, C.c_int, UniqueId, C.c_int]
        with torch.cuda.device(self.device):
            rc = self.nccl.ncclCommInitRank(C.byref(self.comm), world, uid, rank)
        assert rc == 0, f"ncclCommInitRank -> {rc}"

    def allreduce(self, stats: torch.Tensor) -> torch.Tensor:
        out = stats.clone()
        with torch.cuda.device(self.device):
            rc = nv.lib().wh_stats_allreduce(out.data_ptr(), self.comm,
                                             torch.cuda.current_stream(self.device).cuda_stream)
        nv.check(rc, "wh_stats_allreduce")
        return out

    def close(self):
        if self.comm:
            self.nccl.ncclCommDestroy.argtypes = [self.C.c_void_p]
            self.nccl.ncclCommDestroy(self.comm)
            self.comm = None
