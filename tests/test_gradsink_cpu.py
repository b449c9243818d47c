"""GradSink (pdae_b200/train.py): the (shape, stride, offset) records that drive `pdae_unpack_grads` are derived from the
un-packing views on the host -- checked here against torch's own permute/reshape with a numpy emulation of the kernel's
index arithmetic (no GPU)."""
import numpy as np
import torch

from pdae_b200.engine import Buf, BufView
from pdae_b200.train import GradSink, _UNPACK_ITEM


def _emulate(tab_bytes, arena, total):
    items = np.frombuffer(tab_bytes, dtype=_UNPACK_ITEM)
    a = arena.numpy()
    out = np.zeros(total, np.float32)
    for it in items:
        base = (int(it["src"]) - arena.data_ptr()) // 4
        s, st = it["shape"], it["stride"]
        idx = np.arange(int(np.prod(s)))
        i3 = idx % s[3]; r = idx // s[3]
        i2 = r % s[2]; r //= s[2]
        i1 = r % s[1]; i0 = r // s[1]
        v = a[base + i0 * st[0] + i1 * st[1] + i2 * st[2] + i3 * st[3]]
        d = int(it["dst_off"])
        out[d:d + idx.size] = out[d:d + idx.size] + v if it["add"] else v
    return out


def test_unpack_records_reproduce_the_view_semantics():
    arena_t = torch.randn(20000)
    arena = Buf(arena_t.shape, torch.float32, arena_t)
    kk, Cin, Cout, E, total = 9, 6, 4, 8, 20
    w = torch.nn.Parameter(torch.zeros(Cout, Cin, 3, 3))
    b = torch.nn.Parameter(torch.zeros(Cout))
    lw = torch.nn.Parameter(torch.zeros(6, E))
    lb = torch.nn.Parameter(torch.zeros(6))
    fc = torch.nn.Parameter(torch.zeros(5, 12))
    shared = torch.nn.Parameter(torch.zeros(7))
    sink = GradSink()
    sink.add(w, BufView(arena, 100), kk * Cin * Cout, lambda t: t.view(kk, Cin, Cout).permute(2, 1, 0))
    sink.add(b, BufView(arena, 400), Cout, lambda t: t)
    sink.add(lw, BufView(arena, 500), E * total, lambda t: t.view(E, total)[:, 5:11].t())
    sink.add(lb, BufView(arena, 700), total, lambda t: t[5:11])
    sink.add(fc, BufView(arena, 800), 60, lambda t: t.view(4, 3, 5).permute(2, 1, 0))
    sink.add(shared, BufView(arena, 900), 7, lambda t: t)
    sink.add(shared, BufView(arena, 950), 7, lambda t: t)          # a second contribution to the same parameter
    sink._build()
    assert len(sink._launches) == 2                                  # the second contribution runs in a later launch
    out = np.zeros(sink._total, np.float32)
    for t_dev, b_dev, nb in sink._launches:
        part = _emulate(t_dev.numpy().tobytes(), arena_t, sink._total)
        items = np.frombuffer(t_dev.numpy().tobytes(), dtype=_UNPACK_ITEM)
        for it in items:
            d, n = int(it["dst_off"]), int(np.prod(it["shape"]))
            out[d:d + n] = out[d:d + n] + part[d:d + n] if it["add"] else part[d:d + n]
        assert nb == b_dev.shape[0] and nb == sum((int(np.prod(it["shape"])) + 4095) // 4096 for it in items)
    a = arena_t
    expect = {id(w): a[100:316].view(kk, Cin, Cout).permute(2, 1, 0).reshape(w.shape), id(b): a[400:404],
              id(lw): a[500:660].view(E, total)[:, 5:11].t(), id(lb): a[705:711],
              id(fc): a[800:860].view(4, 3, 5).permute(2, 1, 0).reshape(5, 12), id(shared): a[900:907] + a[950:957]}
    for pid, (p, off, n) in sink._slots.items():
        assert off % 4 == 0
        np.testing.assert_array_equal(out[off:off + n], expect[pid].contiguous().numpy().ravel())
