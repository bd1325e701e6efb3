"""CPU tests of the multi-GPU host logic: contiguous sharding of independent filters, final gather and statistics
all-reduce, run over gloo with world_size 2 and 3 (the same code runs over NCCL on GPUs)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from shermbot_navigation_b200 import shard


def test_shard_range_partitions_exactly():
    for total in (0, 1, 7, 64, 65536, 1_000_003):
        for world in (1, 2, 3, 4, 8):
            spans = [shard.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard.shard_range(10, 3, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard.shard_range(total, rank, world)
    full = torch.arange(total * 5, dtype=torch.float64).reshape(total, 5)
    local = full[lo:hi].clone() * 1.0
    got = shard.gather_states(local, total)
    stats = shard.allreduce_stats(torch.tensor([float(hi - lo), float(rank + 1)], dtype=torch.float64))
    ok = torch.equal(got, full) and stats[0].item() == total and stats[1].item() == world * (world + 1) / 2
    np.save(os.path.join(out_dir, f"ok{rank}.npy"), np.array([ok]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,total", [(2, 1001), (3, 64), (2, 2)])
def test_gather_and_stats_over_gloo(tmp_path, world, total):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, total, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert np.load(tmp_path / f"ok{r}.npy")[0]
