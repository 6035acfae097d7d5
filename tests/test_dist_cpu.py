"""Host-side logic of the multi-GPU mode on CPU: stream sharding and the per-step all-gather of the 48-double
per-stream results, world_size 2 over gloo (the GPU run uses the same code over NCCL)."""
import os
import sys

import numpy as np
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_shard_streams_partitions_every_stream_once():
    import bench
    for world in (1, 2, 4, 8):
        owned = [bench.shard_streams(8 * world, world, r) for r in range(world)]
        flat = sorted(s for o in owned for s in o)
        assert flat == list(range(8 * world)) and all(len(o) == 8 for o in owned)
        assert all(s % world == r for r, o in enumerate(owned) for s in o)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    import bench
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ids = bench.shard_streams(6, world, rank)
    local = torch.tensor([[float(s)] * 48 for s in ids], dtype=torch.float64)
    gathered = bench.gather_systems(local, world)
    t = torch.tensor([1.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)  # max-over-ranks timing reduction used by bench.py
    out[rank] = (gathered.numpy().copy(), float(t.item()))
    dist.barrier()
    dist.destroy_process_group()


def test_gather_of_systems_world_size_2():
    world, port = 2, 29500 + (os.getpid() % 2000)
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    for rank in range(world):
        g, tmax = out[rank]
        assert g.shape == (6, 48) and tmax == 2.0
        # rank-major order: rank 0 owns streams 0,2,4 and rank 1 owns 1,3,5
        assert np.array_equal(g[:, 0], np.array([0, 2, 4, 1, 3, 5], dtype=np.float64))
