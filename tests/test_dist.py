"""world_size-2 gloo test of the multi-GPU host logic: utterance sharding + ONE all-reduce of the
flat joiner weight gradient + one of the scalar losses (run on CPU)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from speech2text_b200.distributed import FlatGradBucket, reduce_scalars, shard_bounds


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        lin = torch.nn.Linear(4, 3)  # identical replicas, as under DDP
        bucket = FlatGradBucket(lin.parameters())
        data = torch.arange(8 * 4, dtype=torch.float32).reshape(8, 4) / 10.0  # 8 "utterances"
        mine = data[list(shard_bounds(8, rank, world))]
        bucket.zero()
        loss = lin(mine).pow(2).sum() / mine.shape[0]
        loss.backward()
        bucket.all_reduce(average=True)
        scal = reduce_scalars([loss, loss * 2])
        out.put((rank, bucket.flat.clone(), scal.clone()))
    finally:
        dist.destroy_process_group()


def test_flat_gradient_allreduce_matches_full_batch():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    torch.testing.assert_close(res[0][1], res[1][1])  # every rank holds the same reduced gradient
    torch.testing.assert_close(res[0][2], res[1][2])
    # single-process reference: mean over the two shards' mean losses
    torch.manual_seed(0)
    lin = torch.nn.Linear(4, 3)
    data = torch.arange(8 * 4, dtype=torch.float32).reshape(8, 4) / 10.0
    loss = 0.5 * (lin(data[:4]).pow(2).sum() / 4 + lin(data[4:]).pow(2).sum() / 4)
    loss.backward()
    flat = torch.cat([lin.weight.grad.reshape(-1), lin.bias.grad.reshape(-1)])
    torch.testing.assert_close(res[0][1], flat)
    torch.testing.assert_close(res[0][2][0], loss.detach())
