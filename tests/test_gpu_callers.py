"""Callers either side of the hot path: fused Adam+EMA, uint8 wire formats, MSE/SSIM -- CUDA (through
the C-ABI) vs the fixtures recorded from the reference and vs the CPU oracle."""
import pytest
import torch

from tests import cases
from tests.util import assert_close, load_golden

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _run_adam(cfg, params, grads, with_ema, group_split=None):
    from pdae_b200.optim import FusedAdamEMA
    ps = [torch.nn.Parameter(p.clone().to(DEV)) for p in params]
    emas = [torch.nn.Parameter(p.clone().to(DEV), requires_grad=False) for p in params] if with_ema else None
    groups = [{"params": ps}] if group_split is None else [{"params": ps[:group_split]}, {"params": ps[group_split:]}]
    opt = FusedAdamEMA(groups, lr=cfg["lr"], betas=tuple(cfg["betas"]), eps=cfg["eps"], weight_decay=cfg["weight_decay"],
                       ema_decay=max(cfg["ema_decay"], 0.0), ema_every=1 if with_ema else 0)
    if with_ema:
        for p, e in zip(ps, emas):
            opt._ema[p] = e
    for gs in grads:
        for p, g in zip(ps, gs):
            p.grad = g.clone().to(DEV)
        opt.step()
    torch.cuda.synchronize()
    return ps, emas, opt


@pytest.mark.parametrize("name", ["caller_adam_ema", "caller_adam_wd"])
def test_adam_ema_matches_reference_fixture(name):
    cfg, g = load_golden(name)
    params, grads = cases.adam_case(cfg)
    with_ema = cfg["ema_decay"] >= 0
    ps, emas, opt = _run_adam(cfg, params, grads, with_ema, group_split=2 if with_ema else None)
    for i, p in enumerate(ps):
        # one Adam step moves a weight by ~lr; 1e-6 relative on |p|~0.1 is 1e-3 of a step
        assert_close(p, g[f"p{i}"], rtol=2e-6, atol=1e-8, what=f"{name} p{i}")
        if with_ema:
            assert_close(emas[i], g[f"e{i}"], rtol=2e-6, atol=1e-8, what=f"{name} e{i}")
    st = opt.state[ps[0]]
    assert set(st) == {"step", "exp_avg", "exp_avg_sq"} and int(st["step"]) == cfg["steps"]


def test_adam_ema_large_ragged_tensors_vs_oracle():
    """Sizes that straddle the 65536-element chunking and are not multiples of 4; grad_scale folded in."""
    from oracle import pdae_oracle as O
    from pdae_b200.utils.synth import synth_normal
    shapes = [(200003,), (65536,), (65537,), (3,), (1,), (512, 513)]
    cfg = dict(lr=1e-3, betas=[0.9, 0.999], eps=1e-8, weight_decay=0.0, ema_decay=0.99)
    params = [synth_normal(s, 300 + i) * 0.05 for i, s in enumerate(shapes)]
    grads = [[synth_normal(s, 400 + 10 * k + i) for i, s in enumerate(shapes)] for k in range(3)]
    from pdae_b200.optim import FusedAdamEMA
    ps = [torch.nn.Parameter(p.clone().to(DEV)) for p in params]
    emas = [p.clone().to(DEV) for p in params]
    opt = FusedAdamEMA(ps, lr=cfg["lr"], betas=tuple(cfg["betas"]), eps=cfg["eps"], ema_decay=cfg["ema_decay"])
    for p, e in zip(ps, emas):
        opt._ema[p] = e
    for gs in grads:
        for p, g in zip(ps, gs):
            p.grad = (g * 4.0).to(DEV)        # as if summed over 4 ranks
        opt.step(grad_scale=0.25)
    want_p = [p.clone() for p in params]
    want_e = [p.clone() for p in params]
    O.adam_ema_steps(want_p, grads, cfg["lr"], tuple(cfg["betas"]), cfg["eps"], 0.0, want_e, cfg["ema_decay"])
    for i in range(len(shapes)):
        assert_close(ps[i], want_p[i], rtol=2e-6, atol=1e-8, what=f"p{i}")
        assert_close(emas[i], want_e[i], rtol=2e-6, atol=1e-8, what=f"ema{i}")


def test_adam_skips_params_without_grad_and_attach_ema_pairs_by_name():
    from pdae_b200.optim import FusedAdamEMA
    m = torch.nn.Sequential(torch.nn.Linear(8, 8), torch.nn.Linear(8, 4)).to(DEV)
    m[1].requires_grad_(False)
    import copy
    ema = copy.deepcopy(m).requires_grad_(False)
    opt = FusedAdamEMA(m.parameters(), lr=1e-2, ema_decay=0.5)
    opt.attach_ema(m, ema)
    before = [p.detach().clone() for p in m.parameters()]
    m[0](torch.randn(4, 8, device=DEV)).sum().backward()
    opt.step()
    after = list(m.parameters())
    assert not torch.equal(before[0], after[0]) and torch.equal(before[2], after[2])      # frozen layer untouched
    assert_close(dict(ema.named_parameters())["0.weight"], 0.5 * before[0] + 0.5 * after[0].detach(), rtol=1e-6, atol=1e-7, what="ema")
    assert torch.equal(dict(ema.named_parameters())["1.weight"], before[2])


def test_wire_formats_bit_exact():
    from oracle import pdae_oracle as O
    from pdae_b200.metric import images_to_uint8, uint8_to_images
    from pdae_b200.utils.synth import synth_normal
    cfg, g = load_golden("caller_metrics_io")
    _, b = cases.caller_io_inputs(cfg)
    u8 = images_to_uint8((b * 1.3).to(DEV))
    assert u8.dtype == torch.uint8 and u8.shape == g["u8"].shape and torch.equal(u8.cpu(), g["u8"])
    assert torch.equal(uint8_to_images(g["u8"].to(DEV)).cpu(), g["back"])
    # bigger, out-of-range values, 1 and 3 channels, odd sizes
    for shp, seed in (((5, 3, 64, 64), 1), ((2, 1, 33, 17), 2), ((3, 3, 128, 128), 3)):
        x = synth_normal(shp, 500 + seed)
        assert torch.equal(images_to_uint8(x.to(DEV)).cpu(), O.images_to_uint8_nhwc(x)), shp
        u = O.images_to_uint8_nhwc(x)
        assert torch.equal(uint8_to_images(u.to(DEV)).cpu(), O.uint8_nhwc_to_images(u)), shp
    # all 256 byte values round-trip through normalise -> to_uint8
    allb = torch.arange(256, dtype=torch.uint8).reshape(1, 16, 16, 1).to(DEV)
    assert torch.equal(images_to_uint8(uint8_to_images(allb)), allb)


def test_metrics_match_reference_and_oracle():
    from oracle import pdae_oracle as O
    from pdae_b200.metric import calculate_mse, calculate_ssim
    from pdae_b200.utils.synth import synth_images, synth_normal
    cfg, g = load_golden("caller_metrics_io")
    a, b = cases.caller_io_inputs(cfg)
    assert_close(calculate_mse(a.to(DEV), b.to(DEV)), g["mse"], rtol=1e-5, atol=0, what="mse vs reference")
    assert_close(calculate_ssim(a.to(DEV), b.to(DEV)), g["ssim"], rtol=1e-5, atol=1e-6, what="ssim vs reference")
    for (B, C, S), seed in (((4, 3, 64), 1), ((2, 1, 37), 2), ((2, 3, 128), 3)):
        x = synth_images(B, C, S, 600 + seed)
        y = (x + 0.2 * synth_normal((B, C, S, S), 700 + seed)).clamp(-1, 1)
        assert_close(calculate_mse(x.to(DEV), y.to(DEV)), O.calculate_mse(x, y), rtol=1e-5, atol=0, what=f"mse {B,C,S}")
        assert_close(calculate_ssim(x.to(DEV), y.to(DEV)), O.calculate_ssim(x, y), rtol=1e-5, atol=1e-6, what=f"ssim {B,C,S}")
    x = synth_images(2, 3, 32, 9).to(DEV)
    assert_close(calculate_ssim(x, x), torch.ones(2), rtol=1e-6, atol=1e-6, what="ssim(x,x)=1")
    assert torch.equal(calculate_mse(x, x).cpu(), torch.zeros(2))
