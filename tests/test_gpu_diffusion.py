"""Per-step diffusion arithmetic and the DDIM loops on the GPU vs the reference fixtures."""
import pytest
import torch

from tests import cases
from tests.util import assert_close, load_golden, rel_l2

pytestmark = pytest.mark.gpu


def gd():
    from pdae_b200.diffusion.gaussian_diffusion import GaussianDiffusion
    return GaussianDiffusion(cases.DIFF, torch.device("cuda"))


def test_tables_bit_exact():
    for bt in ("linear", "cosine"):
        from pdae_b200.diffusion.gaussian_diffusion import GaussianDiffusion
        _, g = load_golden("diffusion_tables_" + bt)
        d = GaussianDiffusion({"timesteps": 1000, "betas_type": bt}, torch.device("cuda"))
        for k, v in g.items():
            assert torch.equal(getattr(d, k).cpu(), v), (bt, k)


def test_elementwise_steps():
    from pdae_b200.utils.synth import synth_images, synth_normal
    _, g = load_golden("diffusion_steps")
    d = gd()
    x0, noise, eps = synth_images(4, 3, 8, 21).cuda(), synth_normal((4, 3, 8, 8), 22).cuda(), synth_normal((4, 3, 8, 8), 23).cuda()
    lr = synth_normal((4, 3, 8, 8), 24).clamp(-1, 1).cuda()
    t = g["t"].cuda()
    assert_close(d.q_sample(x0, t, noise), g["q"], rtol=0, atol=0, what="q_sample (bit exact)")
    assert_close(d.noise_p_sample(x0, t, eps, noise=g["p_noise"].cuda()), g["p_sample"], rtol=1e-6, atol=1e-6, what="p_sample")
    assert_close(d.noise_p_sample(x0, t, eps, lr, noise=g["p_noise"].cuda()), g["p_sample_lr"], rtol=1e-6, atol=1e-6, what="p_lr")


def test_ddim_update_matches_oracle_bitwise():
    from oracle import pdae_oracle as O
    from pdae_b200.utils.synth import synth_normal
    d = gd()._ddim("ddim100")
    D = O.DiffusionOracle(cases.DIFF)
    tabs, tmap, S = D._ddim("ddim100")
    assert torch.equal(tmap, d.timestep_map.cpu()) and S == d.timesteps
    x, eps, grad = (synth_normal((6, 3, 8, 8), s) for s in (51, 52, 53))
    t = torch.tensor([0, 1, 50, 99, 100, 37])
    for direction, tt in (("sample", t.clamp(min=1)), ("encode", t.clamp(max=99))):
        for g_ in (grad, None):
            ref = O.ddim_update(tabs, x, tt, eps, g_, direction)
            got = d._update(x.cuda(), tt.cuda(), eps.cuda(), g_.cuda() if g_ is not None else None, direction)
            assert_close(got, ref, rtol=1e-6, atol=1e-6, what=f"ddim {direction} shift={g_ is not None}")


@pytest.mark.parametrize("precision", ["fp32", "bf16", "bf16x3"])
def test_loops(precision):
    from pdae_b200.utils.synth import synth_images, synth_normal
    # chained random-weight steps amplify per-forward differences ~100x per 10 steps: fp32 (1e-6 per forward) gets atol 2e-4,
    # the split-operand tensor-core mode (1e-5 per forward, test_gpu_parity) atol 2e-3
    tol = dict(rtol=1e-3, atol=2e-4) if precision == "fp32" else (dict(rtol=1e-3, atol=2e-3) if precision == "bf16x3" else None)
    d = gd()
    xT, x0 = synth_normal((2, 3, 16, 16), 25).cuda(), synth_images(2, 3, 16, 26).cuda()

    def check(got, want, what):
        if tol:
            assert_close(got, want, what=what, **tol)
        else:
            # stated bf16 tolerance for a 10-step loop on RANDOM weights: the oracle itself amplifies a 1e-5 input
            # perturbation ~50-100x over these 10 steps (tests/test_oracle_golden.py::test_loop_sensitivity), so the
            # per-forward bf16 error (<= 2e-2, test_gpu_parity) may grow to O(0.3); single forwards are the real gate.
            assert rel_l2(got, want) < 0.35, (what, rel_l2(got, want))

    cfg, g = load_golden("loop_unet_ddim10")
    m, _ = cases.model_case({"kind": "unet", "cfg": cfg["cfg"], "size": 16})
    m = m.cuda()
    m.precision = precision
    with torch.no_grad():
        check(d.ddim_sample("ddim10", m, xT), g["sample"], "unet sample")
        check(d.ddim_encode("ddim10", m, x0), g["encode"], "unet encode")
    cfg, g = load_golden("loop_shift_ddim10")
    m, _ = cases.model_case({"kind": "shiftunet", "cfg": cfg["cfg"], "size": 16})
    m = m.cuda()
    m.precision = precision
    z = synth_normal((2, 64), 27).cuda()
    with torch.no_grad():
        check(d.representation_learning_ddim_sample("ddim10", None, m, None, xT, z), g["sample"], "shift sample")
        check(d.representation_learning_ddim_sample("ddim10", None, m, None, xT, z, stop_percent=0.3), g["sample_stop30"], "stop30")
        check(d.representation_learning_ddim_encode("ddim10", None, m, x0, z), g["encode"], "shift encode")
        # generic (black-box callable) path must agree with the fast in-place path
        slow = d._ddim("ddim10")._loop(lambda a, b, c: m(a, b, c), xT, z, "sample", shift=True)
        fast = d.representation_learning_ddim_sample("ddim10", None, m, None, xT, z)
        if precision == "fp32":  # (bf16 on random weights is chaotic over 10 steps; see `check`)
            assert_close(slow, fast, rtol=1e-3, atol=2e-4, what="generic vs fast loop")


def test_ddim_update_fused_into_head_epilogue_matches_separate_kernel():
    """Tensor-core image heads run the DDIM update inside their epilogue (a step = timestep select + network, one graph).  The
    generic loop (black-box callable -> pdae_ddim_step) must agree: same arithmetic, fp32-grade network mode."""
    from pdae_b200.utils.synth import synth_images, synth_normal
    d = gd()
    cfg, g = load_golden("model_shiftunet_b64")      # base 64: the heads take the tensor-core path
    m, inp = cases.model_case(cfg)
    m = m.cuda().eval()
    m.precision = "bf16x3"
    xT, z = inp["x"].cuda(), inp["z"].cuda()
    x0 = synth_images(2, 3, 16, 26).cuda()
    with torch.no_grad():
        for direction, x, stop in (("sample", xT, 0.0), ("sample", xT, 0.3), ("encode", x0, 0.0)):
            dd = d._ddim("ddim10")
            if direction == "sample":
                fast = dd.shift_ddim_sample_loop(m, z, x, stop_percent=stop)
            else:
                fast = dd.shift_ddim_encode_loop(m, z, x)
            slow = dd._loop(lambda a, b, c: m(a, b, c), x, z, direction, shift=True, stop_step=int(stop * dd.timesteps))
            assert_close(fast, slow, rtol=1e-3, atol=2e-3, what=f"fused-epilogue vs generic loop ({direction}, stop {stop})")
        plain = d._ddim("ddim10").ddim_sample_loop(m, xT, z)     # a ShiftUNet used without its shift: epsilon-only update on the grad head
        slow = d._ddim("ddim10")._loop(lambda a, b, c: m(a, b, c)[0], xT, z, "sample", shift=False)
        assert_close(plain, slow, rtol=1e-3, atol=2e-3, what="eps-only update fused into the shift head")
    plan, _ = m.plan_for(2, 16, 16)
    assert "grad" in plan.head_fuse, "fused DDIM head epilogue not available on the tensor-core head"
    assert int(plan.head_fuse["grad"].tensor[0]) == 0, "fusion descriptor must be switched off after the loop"
    e1, g1 = m(xT, g["t"].cuda(), z)                  # plain forward after the loops: untouched by the (disabled) fusion
    from tests.test_gpu_parity import check
    check(e1, g["eps"], "bf16x3", "forward after fused loops (eps)")
    check(g1, g["grad"], "bf16x3", "forward after fused loops (grad)")


@pytest.mark.parametrize("precision", ["fp32", "bf16", "bf16x3"])
def test_autoencoding_and_latent(precision):
    from pdae_b200.utils.synth import synth_images, synth_normal
    d = gd()
    cfg, g = load_golden("loop_autoencode_ddim10")
    dec, _ = cases.model_case({"kind": "shiftunet", "cfg": cfg["cfg"], "size": 64})
    enc, _ = cases.model_case({"kind": "encoder", "size": 64})
    dec, enc = dec.cuda(), enc.cuda()
    dec.precision = enc.precision = precision
    x0 = synth_images(2, 3, 64, 28).cuda()
    with torch.no_grad():
        rec = d.representation_learning_autoencoding("ddim10", "ddim10", enc, dec, x0)
    if precision == "fp32":
        # 20 chained steps on random weights amplify the ~1e-6 per-forward fp32 differences by ~100x per 10 steps
        assert_close(rec, g["recon"], rtol=1e-3, atol=2e-3, what="autoencode")
    elif precision == "bf16x3":
        assert rel_l2(rec, g["recon"]) < 2e-2, rel_l2(rec, g["recon"])   # 1e-5 per forward, amplified over 20 steps
    else:
        assert rel_l2(rec, g["recon"]) < 0.35
    # reconstruction-MSE metric of the reference (metric/utils.py:62-63, images scaled to [0,1]); the 1e-5 bound of
    # BASELINE.json is asserted in fp32 mode; in bf16 mode on random (non-autoencoding) weights we assert 2e-2
    mse = lambda a, b: float((((a.cpu() + 1) / 2 - (b.cpu() + 1) / 2) ** 2).mean())
    print(f"[{precision}] recon-MSE delta vs reference: {abs(mse(rec, x0) - mse(g['recon'], x0)):.3e}")
    assert abs(mse(rec, x0) - mse(g["recon"], x0)) < {"fp32": 1e-5, "bf16x3": 1e-4, "bf16": 2e-2}[precision]

    cfg, g = load_golden("loop_latent_ddim10")
    m, _ = cases.model_case({"kind": "mlp", "cfg": cfg["cfg"]})
    m = m.cuda()
    zT = synth_normal((2, 64), 29).clamp(-1, 1).cuda()
    nb, tmap = d.get_ddim_betas_and_timestep_map("ddim10", d.latent_diffusion_config["alphas_cumprod"].cpu().numpy())
    from pdae_b200.diffusion.ddim import DDIM
    with torch.no_grad():
        out = DDIM(nb, tmap, torch.device("cuda")).latent_ddim_sample_loop(m, zT)
    assert_close(out, g["z"], rtol=1e-3, atol=1e-4, what="latent loop")
