"""Multi-GPU check, launched by tests/test_multi_gpu.py (or by hand) as
    python -m torch.distributed.run --nproc-per-node G tests/dist_rollout_check.py
Each rank owns a contiguous shard of global env ids on its own GPU, runs 230 fused greedy steps with
auto-reset, then reduces the episode statistics (a) with torch.distributed/NCCL and (b) through the
C ABI's wh_stats_allreduce on a raw ncclComm_t. Rank 0 compares both with an unsharded run."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from rllib_warehouse_b200 import MEDIUM, BatchedWarehouse  # noqa: E402
from rllib_warehouse_b200.parallel import RawNcclStats, allreduce_stats, shard_range  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n_total, steps = 3001, 230                      # deliberately not divisible by the world size
    lo, hi = shard_range(n_total, rank, world)
    cfg = MEDIUM.replace(random_num_agents=True)
    env = BatchedWarehouse(cfg, hi - lo, device=dev, seed=99, env_id0=lo, auto_reset=True)
    env.reset()
    for _ in range(steps):
        env.greedy_step(want_actions=False)
    via_torch = allreduce_stats(env.stats)
    raw = RawNcclStats(dev)
    via_raw = raw.allreduce(env.stats)
    torch.cuda.synchronize()
    assert torch.equal(via_torch, via_raw), (via_torch[:6], via_raw[:6])
    # shard invariance of the state itself: gather time/num_agents checksums
    chk = torch.stack([env.state["agent_pos"].long().sum(), env.state["pickup_tgt"].long().sum(),
                       env.state["num_agents"].long().sum()])
    dist.all_reduce(chk)
    if rank == 0:
        whole = BatchedWarehouse(cfg, n_total, device=dev, seed=99, env_id0=0, auto_reset=True)
        whole.reset()
        for _ in range(steps):
            whole.greedy_step(want_actions=False)
        torch.cuda.synchronize()
        assert torch.equal(whole.stats, via_raw), (whole.stats[:6], via_raw[:6])
        ref = torch.stack([whole.state["agent_pos"].long().sum(), whole.state["pickup_tgt"].long().sum(),
                           whole.state["num_agents"].long().sum()])
        assert torch.equal(ref, chk), (ref, chk)
        assert int(via_raw[0]) == n_total
        print(f"multi-gpu ok: world={world} episodes={int(via_raw[0])} return_sum={int(via_raw[1])}")
    raw.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
