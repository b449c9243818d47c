"""Host logic of the plan recorder: buffer liveness / arena assignment (no GPU needed; the plan is never run)."""
import torch

from pdae_b200.engine import Plan, _STREAM


def _overlap(a, b):
    a0, a1 = a.data_ptr(), a.data_ptr() + a.numel() * a.element_size()
    b0, b1 = b.data_ptr(), b.data_ptr() + b.numel() * b.element_size()
    return a0 < b1 and b0 < a1


def test_arena_never_aliases_live_buffers_and_io_is_private():
    P = Plan(torch.device("cpu"), "fp32", check_device=False)
    B, N = 2, 64
    x_in = P.new((B, N), name="in")      # plan input: written by the caller before op 0, first *used* late
    x_in.keep = True
    t = P.new((B,), torch.int64, "t")
    t.keep = True
    freqs = P.fixed(torch.zeros(16))
    chain = []
    prev = P.new((B, 32), name="t0")
    P.call("timestep_embedding", t, B, 32, freqs, prev, _STREAM)
    for i in range(6):                    # temporaries of the same size as the input die before the input's first use
        nxt = P.new((B, N), name=f"tmp{i}")
        P.call("copy_cols", prev, nxt, N, 0, B, min(32, N), _STREAM)
        chain.append((prev, nxt))
        prev = nxt
    out = P.new((B, 2 * N), name="out")
    out.keep = True
    P.call("copy_cols", x_in, out, 2 * N, 0, B, N, _STREAM)
    P.call("copy_cols", prev, out, 2 * N, N, B, N, _STREAM)
    P.finalize()
    bufs = [b for b in P.bufs if b.tensor is not None]
    for a in bufs:
        for b in bufs:
            if a is b:
                continue
            if a.keep or b.keep:
                assert not _overlap(a.tensor, b.tensor), (a.name, b.name)   # I/O buffers share storage with nothing
            elif not (a.last < b.first or b.last < a.first):
                assert not _overlap(a.tensor, b.tensor), (a.name, b.name)   # simultaneously live -> disjoint
    # and recycling does happen for the dead temporaries
    blocks = {b.tensor.data_ptr() for b in bufs if not b.keep}
    assert len(blocks) < len([b for b in bufs if not b.keep])


def test_packed_refresh_tracks_versions():
    from pdae_b200.engine import Packed
    w = torch.nn.Parameter(torch.ones(4, 3))
    pk = Packed([w], lambda: w.detach().t().contiguous() * 2)
    assert torch.equal(pk.tensor, torch.full((3, 4), 2.0))
    with torch.no_grad():
        w.mul_(3)
    pk.refresh()
    assert torch.equal(pk.tensor, torch.full((3, 4), 6.0))


def test_host_cores_respects_quota():
    from pdae_b200.utils.host import host_cores
    import os
    assert 1 <= host_cores() <= (os.cpu_count() or 1)


def test_prologue_ops_are_split_out_and_their_outputs_must_be_private():
    """Step-invariant ops are recorded in a prologue: run once per loop, excluded from the
    per-step launch list; a buffer crossing the boundary must be `keep` or the arena could recycle it."""
    import pytest

    def build(keep):
        P = Plan(torch.device("cpu"), "fp32", check_device=False)
        z = P.new((2, 8), name="z")
        z.keep = True
        with P.prologue():
            zc = P.new((2, 8), name="z_cached")
            zc.keep = keep
            P.call("copy_cols", z, zc, 8, 0, 2, 8, _STREAM)
        out = P.new((2, 8), name="out")
        out.keep = True
        P.call("copy_cols", zc, out, 8, 0, 2, 8, _STREAM)
        P.call("copy_cols", zc, out, 8, 0, 2, 4, _STREAM)
        return P

    P = build(True).finalize()
    assert P._pro_idx == [0] and P._main_idx == [1, 2] and P.n_launch == 2
    with pytest.raises(AssertionError):
        build(False).finalize()


def test_split3_weight_packing_reconstructs_fp32_products():
    """"bf16x3" mode: weights [W_hi | W_hi | W_lo] against activations [a_hi | a_lo | a_hi] reproduce a*W to ~2^-16."""
    from pdae_b200.engine import split3_weights
    torch.manual_seed(0)
    Cout, Cin, taps = 8, 16, 9
    w = torch.randn(Cout, Cin, taps)
    p = split3_weights(w)
    assert p.shape == (taps, Cout, 3 * Cin) and p.dtype == torch.bfloat16
    a = torch.randn(5, Cin)
    a_hi = a.to(torch.bfloat16)
    a_lo = (a - a_hi.float()).to(torch.bfloat16)
    a3 = torch.cat([a_hi, a_lo, a_hi], dim=1).double()                 # what gn_apply_split3 writes
    for t in (0, 4, 8):
        got = a3 @ p[t].double().t()
        want = a.double() @ w[:, :, t].double().t()
        plain = a_hi.double() @ w[:, :, t].to(torch.bfloat16).double().t()
        err, err_plain = (got - want).abs().max().item(), (plain - want).abs().max().item()
        assert err < 2e-4 * want.abs().max().item() and err < err_plain / 50, (err, err_plain)
