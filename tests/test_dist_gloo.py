"""world_size-2 gloo test of the multi-rank host logic (key sharding, flat
gradient all-reduce, max-over-ranks timing) on CPU."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dgppo_b200.trainer import distributed as D


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        keys = np.arange(101)
        mine = D.shard_keys(keys)
        lo, hi = D.shard_bounds(101, rank, world)
        assert (mine == keys[lo:hi]).all()
        # every key is owned exactly once
        cnt = torch.zeros(101)
        cnt[lo:hi] = 1
        dist.all_reduce(cnt)
        assert (cnt == 1).all()
        # flat mean all-reduce == mean of the per-rank gradients
        g = [torch.full((3, 5), float(rank + 1)), torch.arange(4, dtype=torch.float32) * (rank + 1)]
        r = D.allreduce_mean_flat(g)
        assert torch.allclose(r[0], torch.full((3, 5), (1 + world) / 2))
        assert torch.allclose(r[1], torch.arange(4, dtype=torch.float32) * (1 + world) / 2)
        assert r[0].shape == (3, 5) and r[1].shape == (4,)
        assert D.max_over_ranks(10.0 + rank) == 10.0 + world - 1
        assert (D.same_shuffle(32, 9) == D.same_shuffle(32, 9)).all()
        out.put((rank, "ok"))
    except Exception as e:          # pragma: no cover
        out.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = dict(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(timeout=60)
    assert res == {0: "ok", 1: "ok"}, res
