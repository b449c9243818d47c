"""world_size-2 gloo test (CPU) of the data-parallel host logic: shard ranges + the single result all-gather."""
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


def _worker(rank, world, port, n_total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pdae_b200.utils.dist import all_gather_images, shard_range, sharded_autoencode
    full = torch.arange(n_total * 6, dtype=torch.float32).reshape(n_total, 1, 2, 3)
    s, e = shard_range(n_total, rank, world)
    got = all_gather_images(full[s:e] * 2.0, n_total)
    ok = torch.equal(got, full * 2.0)

    class FakeGD:  # the hot path itself needs a GPU; here only the sharding/gather wrapper is under test
        def representation_learning_autoencoding(self, a, b, enc, dec, x):
            assert x.shape[0] > 0, "an empty shard must not reach the hot path (n < world)"
            return x + 1.0
    got2 = sharded_autoencode(FakeGD(), None, None, full)
    ok = ok and torch.equal(got2, full + 1.0)
    # gradient exchange: bucketed SUM all-reduce, 1/world returned for the optimizer to fold in
    from pdae_b200.utils.dist import allreduce_grads_
    ps = [torch.nn.Parameter(torch.zeros(n)) for n in (5, 70000, 3, 12)] + [torch.nn.Parameter(torch.zeros(2))]
    for i, p in enumerate(ps[:-1]):
        p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
    scale = allreduce_grads_(ps, bucket_bytes=4096)          # last param has no grad; several buckets
    ok = ok and scale == 1.0 / world and ps[-1].grad is None
    ok = ok and all(torch.equal(p.grad, torch.full_like(p, 3.0 * (i + 1))) for i, p in enumerate(ps[:-1]))
    # overlapped variant: buckets launch from post-accumulate-grad hooks in gradient-ready order, finish() writes the sums back
    from pdae_b200.utils.dist import OverlappedGradAllReduce
    a = [torch.nn.Parameter(torch.ones(n)) for n in (7, 30000)]
    bq = [torch.nn.Parameter(torch.ones(n)) for n in (5, 11)]
    red = OverlappedGradAllReduce([a, bq], bucket_bytes=64 << 10)
    for it in range(2):   # two steps: the per-step state resets
        loss = sum((p * float(rank + 1 + it)).sum() for p in a + bq)
        loss.backward()
        sc = red.finish()
        ok = ok and sc == 1.0 / world
        want = float(sum(r + 1 + it for r in range(world)))
        ok = ok and all(torch.equal(p.grad, torch.full_like(p, want)) for p in a + bq)
        for p in a + bq:
            p.grad = None
    red.remove()
    # gradients handed out as VIEWS of one flat buffer (what the trainers' GradSink returns: autograd adopts them as `.grad`
    # without a copy) go through both exchanges unchanged
    class _FlatGrads(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x, *params):
            ctx.shapes = [p.shape for p in params]
            return x.sum() + sum(p.sum() for p in params) * 0.0

        @staticmethod
        def backward(ctx, g):
            flat = torch.arange(sum(int(torch.Size(s).numel()) for s in ctx.shapes), dtype=torch.float32) * float(rank + 1)
            outs, off = [], 0
            for s in ctx.shapes:
                n = int(torch.Size(s).numel())
                outs.append(flat[off: off + n].view(s))
                off += n
            return (None, *outs)

    vs = [torch.nn.Parameter(torch.zeros(3, 4)), torch.nn.Parameter(torch.zeros(5)), torch.nn.Parameter(torch.zeros(2, 2, 2))]
    total = torch.arange(12 + 5 + 8, dtype=torch.float32) * float(sum(r + 1 for r in range(world)))
    for mode in ("after", "overlapped"):
        red2 = OverlappedGradAllReduce([vs]) if mode == "overlapped" else None
        _FlatGrads.apply(torch.ones(2), *vs).backward()
        ok = ok and vs[0].grad.untyped_storage().data_ptr() == vs[2].grad.untyped_storage().data_ptr()   # adopted, not copied
        sc = red2.finish() if red2 is not None else allreduce_grads_(vs)
        got = torch.cat([p.grad.reshape(-1) for p in vs])
        ok = ok and sc == 1.0 / world and torch.equal(got, total)
        for p in vs:
            p.grad = None
        if red2 is not None:
            red2.remove()
    q.put((rank, ok, (s, e)))
    dist.destroy_process_group()


def test_two_rank_shard_and_gather():
    ctx = mp.get_context("spawn")
    for n_total in (8, 7, 1):  # even, ragged (remainder to the last rank, as in the reference), fewer images than ranks
        q = ctx.Queue()
        port = _free_port()
        procs = [ctx.Process(target=_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
        for p in procs:
            p.start()
        res = sorted(q.get(timeout=120) for _ in procs)
        for p in procs:
            p.join(timeout=60)
            assert p.exitcode == 0
        assert all(ok for _, ok, _ in res), res
        assert res[0][2] == (0, n_total // 2) and res[1][2] == (n_total // 2, n_total)


def test_shard_range_covers_everything():
    from pdae_b200.utils.dist import shard_range
    for n in (1, 5, 64, 257):
        for w in (1, 2, 4, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
