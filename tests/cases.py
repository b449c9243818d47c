"""Rebuild, from a golden fixture's config, the pdae_b200 module + the seeded inputs that
tests/golden/make_golden.py fed to the real reference.  Used by the CPU oracle tests and the GPU parity tests."""
import torch

from pdae_b200.model import module as pm
from pdae_b200.model.mlp_skip_net import MLPSkipNet
from pdae_b200.model.representation_learning.encoder import CELEBA64Encoder, FFHQEncoder
from pdae_b200.model.shift_unet import ShiftUNet
from pdae_b200.model.unet import UNet
from pdae_b200.utils.synth import fill_module_, synth_images, synth_normal

DIFF = {"timesteps": 1000, "betas_type": "linear"}


def block_case(cfg):
    """-> (module, inputs dict).  Seeds: weights 3/4, emb 11, embz 12, x 13/14 (make_golden.block_cases)."""
    B, E = 2, 128
    if cfg["kind"] == "resblock":
        kw = {k: cfg[k] for k in ("channels", "out_channels", "up", "down") if k in cfg}
        cls = pm.ResBlockShift if cfg["shift"] else pm.ResBlock
        m = fill_module_(cls(emb_channels=E, dropout=0.0, **kw), seed=3).eval()
        inp = {"x": synth_normal((B, cfg["channels"], 8, 8), 13), "emb": synth_normal((B, E), 11)}
        if cfg["shift"]:
            inp["emb_z"] = synth_normal((B, E), 12)
        return m, inp
    if cfg["kind"] == "attention":
        m = fill_module_(pm.AttentionBlock(cfg["channels"], cfg["heads"], -1, cfg["new_order"]), seed=4).eval()
        return m, {"x": synth_normal((B, cfg["channels"], 8, 8), 14)}
    raise KeyError(cfg["kind"])


def model_case(cfg):
    """-> (module, inputs dict, oracle kind).  Seeds per make_golden.model_cases."""
    kind = cfg["kind"]
    if kind == "unet":
        c = cfg["cfg"]
        m = fill_module_(UNet(**c), seed=5).eval()
        return m, {"x": synth_normal((2, c["input_channel"], cfg["size"], cfg["size"]), 15)}
    if kind == "shiftunet":
        c = cfg["cfg"]
        m = fill_module_(ShiftUNet(**c), seed=6).eval()
        return m, {"x": synth_normal((2, 3, cfg["size"], cfg["size"]), 16), "z": synth_normal((2, c["latent_dim"]), 17)}
    if kind == "encoder":
        cls = CELEBA64Encoder if cfg["size"] == 64 else FFHQEncoder
        m = fill_module_(cls(latent_dim=512), seed=7).eval()
        return m, {"x": synth_images(2, 3, cfg["size"], 18)}
    if kind == "mlp":
        m = fill_module_(MLPSkipNet(**cfg["cfg"]), seed=8).eval()
        return m, {"x": synth_normal((2, 64), 19)}
    raise KeyError(kind)


def sd_of(module):
    return {k: v.detach().float().cpu().clone() for k, v in module.state_dict().items()}


def caller_io_inputs(cfg):
    """Inputs of tests/golden/make_golden.py::caller_cases (metrics + wire formats)."""
    a = synth_images(cfg["n"], 3, cfg["size"], 41)
    b = (a + 0.1 * synth_normal((cfg["n"], 3, cfg["size"], cfg["size"]), 42)).clamp(-1, 1)
    return a, b


def adam_case(cfg):
    """Initial params and per-step grads of the caller_adam_* fixtures."""
    shapes = [tuple(s) for s in cfg["shapes"]]
    params = [synth_normal(s, 50 + i) * 0.1 for i, s in enumerate(shapes)]
    scale = (lambda i: 10.0 ** (i - 2)) if cfg["ema_decay"] >= 0 else (lambda i: 1.0)
    grads = [[synth_normal(s, 100 + 10 * step + i) * scale(i) for i, s in enumerate(shapes)] for step in range(cfg["steps"])]
    return params, grads


# ---- glue fixtures (tests/golden/make_golden.py::glue_cases) ---------------------------------------------------------
class CpuStream:
    """Replays the reference's random draws: torch.manual_seed(seed) on the default CPU generator, then randn / randn_like /
    rand_like in call order (moved to `device` for the GPU tests)."""

    def __init__(self, seed, device=None):
        self.device = device
        torch.manual_seed(seed)

    def _mv(self, t):
        return t if self.device is None else t.to(self.device)

    def randn(self, shape):
        return self._mv(torch.randn(tuple(shape)))

    def randn_like(self, x):
        return self._mv(torch.randn(tuple(x.shape)))

    def rand_like(self, x):
        return self._mv(torch.rand(tuple(x.shape)))

    def install(self, gd):
        gd._randn, gd._randn_like, gd._rand_like = self.randn, self.randn_like, self.rand_like
        return gd


def glue_inputs():
    """Seeded inputs shared by the glue fixtures."""
    return dict(xT=synth_normal((2, 3, 16, 16), 25), z1=synth_normal((2, 64), 27), z2=synth_normal((2, 64), 61),
                x_t=synth_normal((4, 3, 8, 8), 62), eps=synth_normal((4, 3, 8, 8), 63),
                lr=synth_normal((4, 3, 8, 8), 64).clamp(-1, 1),
                mean64=synth_normal((1, 64), 65) * 0.1, std64=synth_normal((1, 64), 66).abs() + 0.5,
                x0=synth_images(2, 3, 64, 28), xT64=synth_normal((2, 3, 64, 64), 67),
                mean512=synth_normal((1, 512), 34) * 0.1, std512=synth_normal((1, 512), 35).abs() + 0.5,
                cw=synth_normal((5, 512), 68))
