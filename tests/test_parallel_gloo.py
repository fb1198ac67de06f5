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
        # refine_queries: every rank refines only its strided share of the queries, all ranks end with all poses
        from nefes_b200 import refine as R
        seen = []

        def stub(init_c2w, feat_target, H, W, focal, kw, **opts):
            seen.append(int(feat_target[0, 0]))
            return init_c2w * 2.0 + float(feat_target[0, 0]), []
        real, R.refine_pose = R.refine_pose, stub
        try:
            n_q = 5
            inits = torch.arange(n_q * 12, dtype=torch.float32).reshape(n_q, 3, 4)
            targets = torch.arange(n_q, dtype=torch.float32).reshape(n_q, 1, 1).expand(n_q, 4, 6)
            out = R.refine_queries(inits, targets, 2, 3, 1.0, {})
        finally:
            R.refine_pose = real
        assert seen == list(range(rank, n_q, world))
        assert torch.equal(out, inits * 2.0 + torch.arange(n_q, dtype=torch.float32).reshape(n_q, 1, 1))
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
