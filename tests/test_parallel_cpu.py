"""world_size-2 gloo test of the bucketed gradient all-reduce (host-side logic of the N>1 path)."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from speechmix_b200.parallel import GradientAllReducer, shard_batch
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.ReLU(), torch.nn.Linear(32, 8), torch.nn.Linear(8, 4))
    model[2].bias.requires_grad = False          # frozen parameter: no bucket slot
    red = GradientAllReducer(model, world, bucket_mb=0.0002)   # tiny buckets -> several collectives
    assert len(red.buckets) >= 2
    g = torch.Generator().manual_seed(1)
    X = torch.randn(8, 16, generator=g)
    Y = torch.randn(8, 4, generator=g)
    sl = shard_batch(8, rank, world)
    for _ in range(2):                            # two steps: bucket state must reset
        model.zero_grad(set_to_none=True)
        loss = ((model(X[sl]) - Y[sl]) ** 2).mean()
        loss.backward()
        red.finish()
    got = [p.grad.clone() for p in model.parameters() if p.requires_grad]
    # single-process reference on the full batch
    model.zero_grad(set_to_none=True)
    red.remove()
    ((model(X) - Y) ** 2).mean().backward()
    ref = [p.grad.clone() for p in model.parameters() if p.requires_grad]
    ok = all(torch.allclose(a, b, atol=1e-6) for a, b in zip(got, ref))
    ret[rank] = ok
    dist.destroy_process_group()


def test_bucketed_allreduce_matches_full_batch():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, 29611, ret), nprocs=world, join=True)
    assert all(ret[r] for r in range(world))


def test_shard_batch():
    from speechmix_b200.parallel import shard_batch
    assert [shard_batch(64, r, 8) for r in (0, 7)] == [slice(0, 8), slice(56, 64)]
