"""CPU, world_size 2 over gloo: the host-side multi-GPU logic (ray sharding, one-bucket gradient
all-reduce, strided query sharding + final gather) that bench.py and the refinement driver use."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nefes_b200 import parallel as P
    try:
        lo, hi = P.shard_range(6145, rank, world)
        sizes = [torch.zeros(1, dtype=torch.long) for _ in range(world)]
        dist.all_gather(sizes, torch.tensor([hi - lo]))
        assert sum(int(s) for s in sizes) == 6145 and max(sizes) - min(sizes) <= 1
        # two "models" with flat gradients: per-rank gradient = rank+1 -> sum = 3
        a = torch.nn.Parameter(torch.zeros(1000))
        b = torch.nn.Parameter(torch.zeros(37))
        a.grad = torch.full((1000,), float(rank + 1))
        b.grad = torch.arange(37.) * (rank + 1)
        n = P.allreduce_grads([a, b])
        assert n == 1037
        assert torch.equal(a.grad, torch.full((1000,), 3.0)) and torch.equal(b.grad, torch.arange(37.) * 3)
        # refinement: strided query sharding, gather of per-query [12] poses
        ids = P.shard_strided(7, rank, world)
        local = torch.stack([torch.full((12,), float(i)) for i in ids])
        full = P.gather_rows(local, ids, 7)
        assert torch.equal(full[:, 0], torch.arange(7.))
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(30)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_single_process_is_identity():
    from nefes_b200 import parallel as P
    assert P.world() == (0, 1)
    assert P.shard_range(10, 0, 1) == (0, 10)
    p = torch.nn.Parameter(torch.zeros(3))
    p.grad = torch.ones(3)
    assert P.allreduce_grads([p]) == 0
