"""Drop-in boundary: constructor surface, state_dict keys/shapes, trainable sets and zero-init sets must equal the
reference's (manifest recorded from the real reference by tests/golden/make_manifest.py)."""
import copy
import json
import os

import pytest
import torch

from pdae_b200.model.mlp_skip_net import MLPSkipNet
from pdae_b200.model.representation_learning import decoder, encoder, latent_denoise_fn
from pdae_b200.model import denoise_fn
from pdae_b200.model.shift_unet import ShiftUNet
from pdae_b200.model.unet import UNet
from pdae_b200 import _native
from tests.util import GOLDEN

MAN = json.load(open(os.path.join(GOLDEN, "state_dict_manifest.json")))


def build(name):
    cfg = MAN[name]["cfg"]
    if name.startswith("unet"):
        return UNet(**cfg)
    if name.startswith("shiftunet"):
        return ShiftUNet(**cfg)
    if name == "encoder_celeba64":
        return encoder.CELEBA64Encoder(**cfg)
    if name == "encoder_ffhq":
        return encoder.FFHQEncoder(**cfg)
    return MLPSkipNet(**cfg)


@pytest.mark.parametrize("name", sorted(MAN))
def test_state_dict_matches_reference(name):
    m = build(name)
    sd = m.state_dict()
    assert list(sd.keys()) == list(MAN[name]["keys"].keys()) or sorted(sd.keys()) == sorted(MAN[name]["keys"].keys())
    for k, shape in MAN[name]["keys"].items():
        assert list(sd[k].shape) == shape, k
    assert sorted(k for k, p in m.named_parameters() if p.requires_grad) == MAN[name]["trainable"]
    zero = sorted(k for k, v in sd.items() if v.is_floating_point() and float(v.abs().max()) == 0.0)
    assert zero == MAN[name]["zero_init"]


def test_lookup_names_and_kwargs_swallowed():
    assert denoise_fn.MNISTDenoiseFn is UNet
    for n in ("CELEBA64", "FFHQ", "CELEBAHQ", "BEDROOM", "HORSE"):
        assert getattr(decoder, n + "Decoder") is ShiftUNet
        assert hasattr(encoder, n + "Encoder")
    for n in ("CELEBA64", "FFHQ", "HORSE", "BEDROOM"):
        assert getattr(latent_denoise_fn, n + "LatentDenoiseFn") is MLPSkipNet
    cfg = dict(MAN["unet_mnist"]["cfg"])
    UNet(**cfg, some_unknown_key=1)  # extra keys such as `model` must be accepted


def test_shiftunet_mode_switches_and_deepcopy():
    m = build("shiftunet_celeba64_proxy")
    m.train()
    m.freeze()
    m.set_train_mode()
    assert m.shift_out.training and m.shift_middle_block.training and not m.input_blocks.training and not m.out.training
    m.set_eval_mode()
    assert not m.shift_out.training
    c = copy.deepcopy(m)
    assert sorted(c.state_dict()) == sorted(m.state_dict())
    assert m.label_emb.weight.requires_grad and not m.time_embed[0].weight.requires_grad
    # optimizer groups of the reference trainer (train_representation_learning.py:58-70) address these attributes
    for attr in ("label_emb", "shift_middle_block", "shift_output_blocks", "shift_out"):
        assert len(list(getattr(m, attr).parameters())) > 0


def test_no_cpu_fallback():
    m = build("encoder_celeba64").eval()
    with torch.no_grad(), pytest.raises(_native.NativeError):
        m(torch.zeros(1, 3, 64, 64))


def test_dropin_aliases():
    import subprocess
    import sys
    code = ("import pdae_b200.dropin as d; d.install(); from model.unet import UNet; from model.shift_unet import ShiftUNet;"
            "import model.representation_learning.decoder as dm; from diffusion.gaussian_diffusion import GaussianDiffusion;"
            "from diffusion.ddim import DDIM; import pdae_b200.model.unet as u; assert UNet is u.UNet and dm.FFHQDecoder is ShiftUNet;"
            "print('ok')")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-c", code], cwd=root, capture_output=True, text=True)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr


def test_fused_adam_surface_and_no_cpu_fallback():
    """FusedAdamEMA keeps torch.optim.Adam's constructor / param_groups / state_dict surface
    (trainer/train_representation_learning.py:54-69) and refuses CPU parameters instead of falling back."""
    import torch
    from pdae_b200 import _native
    from pdae_b200.metric import calculate_mse, images_to_uint8
    from pdae_b200.optim import FusedAdamEMA
    a, b = torch.nn.Parameter(torch.zeros(4)), torch.nn.Parameter(torch.zeros(3))
    opt = FusedAdamEMA([{"params": [a]}, {"params": [b], "lr": 5e-4}], lr=float("1e-4"), betas=eval("(0.9, 0.999)"),
                       eps=float("1e-8"), weight_decay=float("0.0"))
    assert [g["lr"] for g in opt.param_groups] == [1e-4, 5e-4]
    assert set(opt.state_dict()) == {"state", "param_groups"}
    opt.zero_grad()
    opt.step()                       # no grads: nothing to do, no native call needed
    a.grad = torch.ones(4)
    with pytest.raises(_native.NativeError):
        opt.step()
    with pytest.raises(_native.NativeError):
        calculate_mse(torch.zeros(1, 3, 4, 4), torch.zeros(1, 3, 4, 4))
    with pytest.raises(_native.NativeError):
        images_to_uint8(torch.zeros(1, 3, 4, 4))
