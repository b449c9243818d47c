"""Single-forward parity vs the CPU oracle at the configs that are benched / named by BASELINE.json:
ffhq128-proxy and ffhq256-proxy ShiftUNet (config/ffhq_representation_learning.yml:4, model/shift_unet.py:253-284), the real
MNIST UNet (config/mnist_regular.yml:14-28) and the ffhq_latent MLPSkipNet (config/ffhq_latent.yml:16-23), every precision mode.
At 256 px a GroupNorm group spans 4 channels x 65 536 pixels -- the sum / sum-of-squares cancellation case."""
import pytest
import torch

from oracle import pdae_oracle as O
from tests import cases
from tests.configs import FFHQ128_PROXY, FFHQ256_PROXY, FFHQ_LATENT, MNIST
from tests.test_gpu_parity import check
from tests.util import assert_close, rel_l2

pytestmark = pytest.mark.gpu


def _shift_case(proxy, size, seed):
    from pdae_b200.model.shift_unet import ShiftUNet
    from pdae_b200.utils.synth import fill_module_, synth_normal
    cfg = dict(proxy, latent_dim=512)
    m = fill_module_(ShiftUNet(**cfg), seed=seed).eval()
    x, z = synth_normal((1, 3, size, size), seed + 1), synth_normal((1, 512), seed + 2)
    t = torch.tensor([431])
    with torch.no_grad():
        ref = O.shiftunet_forward(cases.sd_of(m), cfg, x, t, z)
    return m.cuda(), x.cuda(), t.cuda(), z.cuda(), ref


@pytest.mark.parametrize("name,proxy,size", [("ffhq128-proxy", FFHQ128_PROXY, 128), ("ffhq256-proxy", FFHQ256_PROXY, 256)])
def test_ffhq_proxy_shiftunet_forward_all_modes(name, proxy, size):
    m, x, t, z, (eps_ref, grad_ref) = _shift_case(proxy, size, 71)
    for precision in ("bf16", "bf16x3", "fp32"):
        m.precision = precision
        with torch.no_grad():
            eps, grad = m(x, t, z)
        check(eps, eps_ref, precision, f"{name} B=1 eps")
        check(grad, grad_ref, precision, f"{name} B=1 grad")
        m._plans().clear()   # free this mode's arena before the next one (the 256-px plans are large)
        torch.cuda.empty_cache()


def test_groupnorm_statistics_with_large_mean_at_256px():
    """The tensor-core epilogue accumulates per-channel sum / sum^2 in fp32: a conv output with |mean| >> std over a
    4 x 65 536-element group is the cancellation case.  One ResBlock at 256x256, input offset by +6."""
    from pdae_b200.model import module as pm
    from pdae_b200.utils.synth import fill_module_, synth_normal
    E = 512
    blk = fill_module_(pm.ResBlockShift(channels=128, emb_channels=E, dropout=0.0), seed=73).eval()
    with torch.no_grad():
        blk.in_layers[2].bias.add_(4.0)      # pushes the mean of h (the GroupNorm-2 input) far from zero
    x = synth_normal((1, 128, 256, 256), 74) * 0.5 + 6.0
    emb, embz = synth_normal((1, E), 75), synth_normal((1, E), 76)
    sd = {"blk." + k: v for k, v in cases.sd_of(blk).items()}
    with torch.no_grad():
        ref = O.resblock(sd, "blk", x, emb, embz)
    blk = blk.cuda()
    for precision in ("fp32", "bf16x3", "bf16"):
        blk.precision = precision
        with torch.no_grad():
            y = blk(x.cuda(), emb.cuda(), embz.cuda())
        r = rel_l2(y, ref)
        print(f"[{precision}] ResBlockShift 256x256 offset input: rel-L2 {r:.3e}")
        assert r < {"fp32": 1e-4, "bf16x3": 3e-4, "bf16": 2e-2}[precision], (precision, r)


def test_mnist_unet_real_config():
    from pdae_b200.model.unet import UNet
    from pdae_b200.utils.synth import fill_module_, synth_normal
    cfg = {k: v for k, v in MNIST.items() if k != "model"}
    m = fill_module_(UNet(**cfg), seed=77).eval()
    x = synth_normal((4, 1, 32, 32), 78)
    t = torch.tensor([0, 17, 500, 999])
    with torch.no_grad():
        ref = O.unet_forward(cases.sd_of(m), cfg, x, t, None)
    m = m.cuda()
    for precision in ("fp32", "bf16x3", "bf16"):
        m.precision = precision
        with torch.no_grad():
            y = m(x.cuda(), t.cuda())
        check(y, ref, precision, "MNIST UNet (config/mnist_regular.yml) B=4")


def test_ffhq_latent_mlp_real_config():
    from pdae_b200.model.mlp_skip_net import MLPSkipNet
    from pdae_b200.utils.synth import fill_module_, synth_normal
    cfg = {k: v for k, v in FFHQ_LATENT.items() if k != "model"}
    m = fill_module_(MLPSkipNet(**cfg), seed=79).eval()
    z = synth_normal((8, 512), 80)
    t = torch.tensor([0, 1, 5, 50, 300, 700, 998, 999])
    with torch.no_grad():
        ref = O.mlp_skip_net_forward(cases.sd_of(m), cfg, z, t)
    m = m.cuda()
    with torch.no_grad():
        y = m(z.cuda(), t.cuda())
    assert_close(y, ref, rtol=1e-3, atol=1e-4, what="ffhq_latent MLPSkipNet (2048 x 10 layers)")
