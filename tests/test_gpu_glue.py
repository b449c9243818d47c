"""The glue around the hot path on the GPU vs fixtures recorded from the real reference (SURVEY.md section 8 rows a16,
a20, a22, a24): trajectory interpolation (ddim.py:149-174), x_0_clip_p_sample (gaussian_diffusion.py:130-146), the
DDPM loops (:216-229, :257-270), latent_diffusion_sample with stop_percent=0.3 (:400-415), manipulation_sample
(:435-443), gap measure (:292-318), one-step denoising (:320-334), latent_ddim_sample (ddim.py:178-198).
The reference's random draws (default CPU generator after torch.manual_seed) are replayed through
GaussianDiffusion._randn / _randn_like / _rand_like (tests/cases.py::CpuStream)."""
import numpy as np
import pytest
import torch

from tests import cases
from tests.util import assert_close, load_golden, rel_l2

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda")
# chained random-weight steps amplify the ~1e-6 per-forward fp32 differences (tests/test_oracle_golden.py::test_loop_sensitivity)
LOOP = dict(rtol=1e-3, atol=2e-3)
ONE = dict(rtol=1e-3, atol=1e-4)


def _gd(cfg=None):
    from pdae_b200.diffusion.gaussian_diffusion import GaussianDiffusion
    return GaussianDiffusion(cfg or cases.DIFF, DEV)


def _net(kind, cfg, size, precision="fp32"):
    m, _ = cases.model_case({"kind": kind, "cfg": cfg, "size": size})
    m = m.cuda().eval()
    m.precision = precision
    return m


def _cu(i):
    return {k: v.cuda() for k, v in i.items()}


@pytest.mark.parametrize("precision", ["fp32", "bf16x3"])
def test_trajectory_interpolation(precision):
    i = _cu(cases.glue_inputs())
    cfg, g = load_golden("glue_interpolation")
    dec = _net("shiftunet", cfg["cfg"], 16, precision)
    with torch.no_grad():
        y = _gd().representation_learning_ddim_trajectory_interpolation(cfg["style"], dec, i["z1"], i["z2"], i["xT"], cfg["alpha"])
    assert_close(y, g["y"], what="trajectory interpolation", **LOOP)


def test_x0_clip_p_sample_and_latent_single_step():
    i = _cu(cases.glue_inputs())
    cfg, g = load_golden("glue_x0_clip")
    d = _gd()
    t = g["t"].cuda()
    for name, kw in (("fixed_clip", {}), ("fixed_noclip", dict(clip_x_0=False)), ("learned_clip", dict(learned_range=i["lr"]))):
        cases.CpuStream(cfg["seed"], DEV).install(d)
        x_in = i["x_t"].clone()
        got = d.x_0_clip_p_sample(x_in, t, i["eps"], **kw)
        assert_close(got, g[name], rtol=1e-5, atol=1e-5, what=name)
        assert torch.equal(x_in, i["x_t"]), "x_0_clip_p_sample must not mutate its input"
    # DDIM.latent_ddim_sample (ddim.py:178-198): the unclamped step, vs its closed form with torch ops on the same tables
    dd = d._ddim("ddim10", d.latent_diffusion_config["alphas_cumprod"])
    z, tt = i["z1"], torch.tensor([3, 10], device=DEV)
    fn = lambda a, b: a * 0.5 + b.float()[:, None] * 1e-3
    out = dd.latent_ddim_sample(fn, z, tt)
    eps = fn(z, dd.timestep_map[tt])
    A, Bm, ap = (tab[tt][:, None] for tab in (dd.sqrt_recip_alphas_cumprod, dd.sqrt_recip_alphas_cumprod_m1, dd.alphas_cumprod_prev))
    assert_close(out, (A * z - Bm * eps) * ap.sqrt() + (1 - ap).sqrt() * eps, rtol=1e-6, atol=1e-6, what="latent_ddim_sample")


def test_ddpm_loops():
    i = _cu(cases.glue_inputs())
    cfg, g = load_golden("glue_ddpm")
    d = _gd({"timesteps": cfg["timesteps"], "betas_type": "linear"})
    s = cfg["seeds"]
    # (modules are built BEFORE seeding: nn.Conv2d's default init draws from the same CPU generator)
    unet, unet_ls, dec = _net("unet", cfg["cfg_unet"], 16), _net("unet", cfg["cfg_sigma"], 16), _net("shiftunet", cfg["cfg_shift"], 16)
    with torch.no_grad():
        cases.CpuStream(s[0], DEV).install(d)
        assert_close(d.regular_ddpm_sample(unet, i["xT"]), g["regular"], what="regular ddpm", **LOOP)
        cases.CpuStream(s[1], DEV).install(d)
        assert_close(d.regular_ddpm_sample(unet_ls, i["xT"]), g["learned_sigma"], what="ddpm with learned sigma", **LOOP)
        cases.CpuStream(s[2], DEV).install(d)
        assert_close(d.representation_learning_ddpm_sample(None, dec, i["xT"], i["xT"], i["z1"]), g["representation"],
                     what="representation ddpm", **LOOP)


@pytest.mark.parametrize("precision", ["fp32", "bf16x3"])
def test_latent_diffusion_sample_with_stop_percent(precision):
    i = _cu(cases.glue_inputs())
    cfg, g = load_golden("glue_latent_sample")
    mlp, _ = cases.model_case({"kind": "mlp", "cfg": cfg["cfg_mlp"]})
    mlp = mlp.cuda().eval()
    dec = _net("shiftunet", cfg["cfg_shift"], 16, precision)
    d = cases.CpuStream(cfg["seed"], DEV).install(_gd())     # seed AFTER building the modules (their default init draws too)
    with torch.no_grad():
        y = d.latent_diffusion_sample("ddim10", "ddim10", mlp, dec, i["xT"], i["mean64"], i["std64"])
    assert_close(y, g["y"], what="latent_diffusion_sample (stop_percent 0.3 tail on the epsilon-only plan)", **LOOP)


def test_manipulation_gap_and_one_step_denoise():
    i = _cu(cases.glue_inputs())
    cfg, g = load_golden("glue_manipulation")
    dec = _net("shiftunet", cfg["cfg"], 64)
    enc, _ = cases.model_case({"kind": "encoder", "size": 64})
    enc = enc.cuda().eval()
    enc.precision = "fp32"
    d = _gd()
    with torch.no_grad():
        y = d.manipulation_sample(cfg["style"], i["cw"], enc, dec, i["x0"], i["xT64"], i["mean512"], i["std512"], cfg["class_id"],
                                  cfg["scale"])
    assert_close(y, g["y"], what="manipulation_sample", **LOOP)
    cfg, g = load_golden("glue_gap")
    d8 = cases.CpuStream(cfg["seed"], DEV).install(_gd({"timesteps": cfg["timesteps"], "betas_type": "linear"}))
    with torch.no_grad():
        gp, ga = d8.representation_learning_gap_measure(enc, dec, i["x0"])
    np.testing.assert_allclose(gp, g["gap_pred"].numpy(), rtol=2e-3)
    np.testing.assert_allclose(ga, g["gap_ae"].numpy(), rtol=2e-3)
    cfg, g = load_golden("glue_denoise_one_step")
    d = cases.CpuStream(cfg["seed"], DEV).install(_gd())
    with torch.no_grad():
        p0, a0 = d.representation_learning_denoise_one_step(enc, dec, i["x0"], cfg["timesteps"])
    # predicted x_0 = A x_t - B eps with B up to ~70 at t=700: the absolute tolerance scales with max|ref|
    for got, want, what in ((p0, g["pred"], "pred"), (a0, g["ae"], "ae")):
        assert rel_l2(got, want) < 1e-4, (what, rel_l2(got, want))
        assert_close(got, want, rtol=1e-3, atol=1e-4 * float(want.abs().max()), what="denoise_one_step " + what)
