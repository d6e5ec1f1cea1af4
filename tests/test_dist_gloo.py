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


def _update_worker(rank, world, port, out):
    """Each rank takes half of a minibatch: after the flat mean all-reduce inside clip_and_step both ranks
    must hold the parameters a single process gets from the whole minibatch."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from dgppo_b200.algo import params as P
        from dgppo_b200.algo import update as U
        import sys
        sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
        from test_update_cpu import make_graph, torch_graph
        rng = np.random.default_rng(7)
        mb, T = 4, 2
        gr, dims = make_graph(rng, mb * T)
        n = dims[0]
        hs = rng.standard_normal((mb, T, n, 64)) * 0.3
        tgt = rng.standard_normal((mb, T, n, 2))
        vh = P.init_value_params(7, 4, 2, 1, seed=6, jitter=0.1)

        def run(rows, distributed):
            sel = {k: v.reshape((mb, T) + v.shape[1:])[rows].reshape((len(rows) * T,) + v.shape[1:]) for k, v in gr.items()}
            tg, gi = torch_graph(sel, dims, torch.float64)
            st = U.NetTrainState(vh, "cpu", 1e-3, dtype=torch.float64)        # the product path's optimiser state
            loss = U.loss_Vh(st.tree(), tg, torch.tensor(hs[rows]), torch.tensor(tgt[rows]), gi, 1)
            if distributed:
                st.step(loss, 2.0)
            else:       # single process: same code path with the collective switched off
                saved = U.D.allreduce_mean_flat
                U.D.allreduce_mean_flat = lambda g: list(g)
                try:
                    st.step(loss, 2.0)
                finally:
                    U.D.allreduce_mean_flat = saved
            return st.flat.detach().clone()
        mine = run(list(range(rank * mb // world, (rank + 1) * mb // world)), True)
        full = run(list(range(mb)), False)
        assert torch.allclose(mine, full, rtol=1e-9, atol=1e-12), float((mine - full).abs().max())
        out.put((rank, "ok"))
    except Exception as e:          # pragma: no cover
        import traceback
        out.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_two_rank_update_equals_single_process():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_update_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = dict(q.get(timeout=240) for _ in ps)
    for p in ps:
        p.join(timeout=60)
    assert res == {0: "ok", 1: "ok"}, res


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
