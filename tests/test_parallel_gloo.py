"""CPU, world_size 2 over gloo: the one collective of the path (checkpoint replication) and the batch
sharding produce on every rank exactly what a single process would."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from torch import nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class Toy(nn.Module):
    def __init__(self, seed):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.lin = nn.Linear(48, 32)
        with torch.no_grad():
            self.lin.weight.copy_(torch.randn(32, 48, generator=g))
            self.lin.bias.copy_(torch.randn(32, generator=g))
        self.register_buffer("f8", torch.randn(64, 48, generator=g).to(torch.float8_e4m3fn))
        self.register_buffer("scale", torch.rand((), generator=g))
        self.register_buffer("bf", torch.randn(7, generator=g).to(torch.bfloat16))


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from flux_fp8_api_b200 import parallel as PAR

    PAR.init_distributed("gloo")
    model = Toy(seed=rank)  # ranks start from different states
    sent = PAR.broadcast_state(model, src=0, bucket_bytes=1024)  # tiny buckets: exercise the flush logic
    ref = Toy(seed=0)
    same = all(torch.equal(a.view(torch.uint8) if a.dtype.itemsize == 1 else a, b.view(torch.uint8) if b.dtype.itemsize == 1 else b)
               for (_, a), (_, b) in zip(sorted(model.state_dict().items()), sorted(ref.state_dict().items())))
    total = 5
    req = {"img": torch.arange(total * 3, dtype=torch.float32).reshape(total, 3), "y": torch.arange(total)}
    mine = PAR.shard_request(req, rank, world)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine["img"])
    whole = torch.cat(gathered)
    mx = PAR.max_over_ranks(float(rank + 1), "cpu")
    PAR.barrier()
    ret[rank] = (same, sent, torch.equal(whole, req["img"]), mx)
    dist.destroy_process_group()


def test_broadcast_and_shard_world2():
    world = 2
    port = 29500 + (os.getpid() % 2000)
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
        out = dict(ret)
    expected_bytes = sum(t.numel() * t.element_size() for t in Toy(0).state_dict().values())
    for rank in range(world):
        same, sent, covered, mx = out[rank]
        assert same, f"rank {rank}: state differs from rank 0's after broadcast"
        assert sent == expected_bytes
        assert covered, "shards do not reassemble to the full batch"
        assert mx == float(world)
