"""Training step (diffusion/gaussian_diffusion.py:234-255): loss and EVERY trainable gradient of the hand-written
backward vs (a) the fixture recorded from the real reference and (b) torch.autograd through the CPU oracle."""
import pytest
import torch

from oracle import pdae_oracle as O
from tests import cases
from tests.util import assert_close, load_golden, rel_l2

pytestmark = pytest.mark.gpu


def _loss(gd, enc, dec, x0, t, noise):
    z = enc(x0)
    x_t = gd.q_sample(x0, t, noise)
    eps, grad = dec(x_t, t, z)
    s = x0.shape
    target = eps + gd.extract_coef_at_t(gd.shift_coef, t, s) * grad
    return gd.p_loss(noise, target, weight=gd.extract_coef_at_t(gd.weight, t, s))


def test_representation_learning_step_matches_reference_and_oracle():
    from pdae_b200.diffusion.gaussian_diffusion import GaussianDiffusion
    from pdae_b200.utils.synth import synth_images
    cfg, g = load_golden("train_representation_learning")
    dec, _ = cases.model_case({"kind": "shiftunet", "cfg": cfg["cfg"], "size": 64})
    enc, _ = cases.model_case({"kind": "encoder", "size": 64})
    x0 = synth_images(2, 3, 64, 31)
    # ---- oracle autograd on the CPU (same weights, same t / noise as the reference run) ----
    dsd = {k: v.requires_grad_(k.startswith(("label_emb", "shift_"))) for k, v in cases.sd_of(dec).items()}
    esd = {k: v.requires_grad_(True) for k, v in cases.sd_of(enc).items()}
    D = O.DiffusionOracle(cases.DIFF)
    ref_loss = D.representation_learning_loss(lambda x: O.encoder_forward(esd, "celeba64", x),
                                              lambda x, t, z: O.shiftunet_forward(dsd, cfg["cfg"], x, t, z), x0, g["t"], g["noise"])
    ref_loss.backward()
    # ---- native path ----
    dev = torch.device("cuda")
    dec, enc = dec.cuda().train(), enc.cuda().train()
    dec.freeze()
    dec.set_train_mode()
    dec.precision = enc.precision = "fp32"
    gd = GaussianDiffusion(cases.DIFF, dev)
    loss = _loss(gd, enc, dec, x0.cuda(), g["t"].cuda(), g["noise"].cuda())
    assert_close(loss, g["loss"], rtol=1e-4, atol=1e-7, what="loss vs reference fixture")
    assert_close(loss, ref_loss, rtol=1e-4, atol=1e-7, what="loss vs oracle")
    loss.backward()
    named = dict(dec.named_parameters())
    n_grad = sum(1 for p in list(dec.parameters()) + list(enc.parameters()) if p.grad is not None)
    assert n_grad == cfg["n_params_with_grad"], (n_grad, cfg["n_params_with_grad"])
    assert all(p.grad is None for k, p in named.items() if not k.startswith(("label_emb", "shift_")))
    worst = (0.0, "")
    bad = []
    gmax = max(float(v.grad.abs().max()) for v in list(dsd.values()) + list(esd.values()) if v.grad is not None)
    floor = 2e-4 * gmax   # absolute floor relative to the largest gradient entry of the step
    print(f"largest |grad| entry {gmax:.3e}, absolute floor {floor:.3e}")
    for prefix, mod, sd in (("dec.", dec, dsd), ("enc.", enc, esd)):
        for k, p in mod.named_parameters():
            if sd[k].grad is None:
                continue
            r = rel_l2(p.grad, sd[k].grad)
            err = float((p.grad.cpu() - sd[k].grad).abs().max())
            ref_max = float(sd[k].grad.abs().max())
            # biases feeding a GroupNorm have analytically ~zero gradient (large sums that cancel): absolute floor
            if r >= 2e-3 and err > 2e-3 * ref_max + floor:
                bad.append((prefix + k, tuple(p.shape), round(r, 4), f"err={err:.2e} ref_max={ref_max:.2e}"))
            elif ref_max > 50 * floor:
                worst = max(worst, (r, prefix + k))
    print("worst grad rel-L2:", worst)
    assert not bad, "gradients off: " + "; ".join(map(str, bad[:40]))
    # spot-check against the fixture from the real reference
    for key, ref in (("label_emb.weight", "g_label_emb_weight"), ("shift_out.2.weight", "g_shift_out_2_weight"),
                     ("shift_middle_block.0.in_layers.2.weight", "g_shift_middle_block_0_in_layers_2_weight"),
                     ("shift_output_blocks.0.0.emb_z_layers.1.weight", "g_shift_output_blocks_0_0_emb_z_layers_1_weight")):
        got = named[key].grad.flatten()[:512].cpu()
        assert rel_l2(got, g[ref]) < 2e-3, key
        assert_close(named[key].grad.double().norm().float(), g["n" + ref], rtol=2e-3, atol=0, what=key + " norm")
    assert rel_l2(dict(enc.named_parameters())["encoder.0.weight"].grad.flatten()[:512].cpu(), g["g_enc_encoder_0_weight"]) < 2e-3


def test_second_step_after_optimizer_update_uses_new_weights():
    """Plans cache packed weights: an Adam step must be seen by the next forward/backward."""
    from pdae_b200.diffusion.gaussian_diffusion import GaussianDiffusion
    from pdae_b200.utils.synth import synth_images
    cfg, g = load_golden("train_representation_learning")
    dec, _ = cases.model_case({"kind": "shiftunet", "cfg": cfg["cfg"], "size": 64})
    enc, _ = cases.model_case({"kind": "encoder", "size": 64})
    dec, enc = dec.cuda().train(), enc.cuda().train()
    dec.freeze()
    dec.set_train_mode()
    dec.precision = enc.precision = "fp32"
    gd = GaussianDiffusion(cases.DIFF, torch.device("cuda"))
    params = [p for p in list(dec.parameters()) + list(enc.parameters()) if p.requires_grad]
    opt = torch.optim.Adam(params, lr=1e-3)
    x0, t, noise = synth_images(2, 3, 64, 31).cuda(), g["t"].cuda(), g["noise"].cuda()
    losses = []
    for _ in range(3):
        opt.zero_grad()
        loss = _loss(gd, enc, dec, x0, t, noise)
        loss.backward()
        opt.step()
        losses.append(float(loss.detach()))
    assert losses[2] < losses[0], losses   # same batch, three Adam steps: the loss must go down


@pytest.mark.parametrize("name", ["unet_tiny", "unet_class"])
def test_regular_dpm_training_step(name):
    """regular_train_one_batch (gaussian_diffusion.py:199-211): full-UNet backward incl. skip-connection gradients,
    down/up-sampling blocks, attention, the time-embedding MLP and the class embedding -- vs oracle autograd."""
    from pdae_b200.diffusion.gaussian_diffusion import GaussianDiffusion
    from pdae_b200.utils.synth import synth_images
    cfg, g = load_golden("model_" + name)
    net, _ = cases.model_case(cfg)
    c = cfg["cfg"]
    size = cfg["size"]
    x0 = synth_images(2, c["input_channel"], size, 32)
    t = torch.tensor([7, 805])
    noise = torch.randn(x0.shape, generator=torch.Generator().manual_seed(3))
    cond = g["cond"] if "cond" in g else None
    sd = {k: v.requires_grad_(True) for k, v in cases.sd_of(net).items()}
    D = O.DiffusionOracle(cases.DIFF)
    ref_loss = D.regular_loss(lambda x, tt, cc: O.unet_forward(sd, c, x, tt, cc), x0, t, noise, cond)
    ref_loss.backward()
    net = net.cuda().train()
    net.precision = "fp32"
    gd = GaussianDiffusion(cases.DIFF, torch.device("cuda"))
    x_t = gd.q_sample(x0.cuda(), t.cuda(), noise.cuda())
    loss = gd.p_loss(noise.cuda(), net(x_t, t.cuda(), cond.cuda() if cond is not None else None))
    assert_close(loss, ref_loss, rtol=1e-4, atol=1e-7, what="regular loss")
    if name == "unet_tiny":
        _, gt = load_golden("train_regular")   # the reference's own loss for this net uses other (t, noise): just a sanity range
        assert float(gt["loss"]) > 0
    loss.backward()
    gmax = max(float(v.grad.abs().max()) for v in sd.values() if v.grad is not None)
    floor = 2e-4 * gmax
    bad, worst = [], (0.0, "")
    for k, p in net.named_parameters():
        assert (p.grad is None) == (sd[k].grad is None), k
        if p.grad is None:
            continue
        r = rel_l2(p.grad, sd[k].grad)
        err = float((p.grad.cpu() - sd[k].grad).abs().max())
        ref_max = float(sd[k].grad.abs().max())
        if r >= 2e-3 and err > 2e-3 * ref_max + floor:
            bad.append((k, tuple(p.shape), round(r, 4), f"err={err:.2e} ref_max={ref_max:.2e}"))
        elif ref_max > 50 * floor:
            worst = max(worst, (r, k))
    print(name, "worst grad rel-L2:", worst)
    assert not bad, "gradients off: " + "; ".join(map(str, bad[:30]))


def test_training_with_dropout_matches_oracle_given_the_same_masks():
    """nn.Dropout(p=0.1) inside the trainable ResBlockShift blocks (module.py:259): masks are drawn on the GPU, then the SAME
    masks are injected into the CPU oracle; loss and all gradients must agree."""
    from pdae_b200.diffusion.gaussian_diffusion import GaussianDiffusion
    from pdae_b200.model.shift_unet import ShiftUNet
    from pdae_b200.utils.synth import fill_module_, synth_images
    cfg, g = load_golden("train_representation_learning")
    c = dict(cfg["cfg"], dropout=0.1)
    dec = fill_module_(ShiftUNet(**c), seed=6)
    enc, _ = cases.model_case({"kind": "encoder", "size": 64})
    dsd = {k: v.requires_grad_(k.startswith(("label_emb", "shift_"))) for k, v in cases.sd_of(dec).items()}
    esd = {k: v.requires_grad_(True) for k, v in cases.sd_of(enc).items()}
    x0 = synth_images(2, 3, 64, 31)
    dec, enc = dec.cuda().train(), enc.cuda().train()
    dec.freeze()
    dec.set_train_mode()
    dec.precision = enc.precision = "fp32"
    gd = GaussianDiffusion(cases.DIFF, torch.device("cuda"))
    torch.manual_seed(5)
    loss = _loss(gd, enc, dec, x0.cuda(), g["t"].cuda(), g["noise"].cuda())
    loss.backward()
    trainer = list(dec._train_cache.values())[0]
    names = {id(m): n for n, m in dec.named_modules()}
    masks = {names[id(blk)]: mk.tensor.permute(0, 3, 1, 2).contiguous().cpu() for blk, mk, p in trainer.fwd.dropout_masks}
    assert len(masks) >= 5 and all(0.8 < float(m.mean()) < 0.98 for m in masks.values())
    assert all(n.startswith("shift_") for n in masks)       # the frozen half stays in eval mode: no dropout there
    O.DROPOUT_MASKS = dict(masks, p=0.1)
    try:
        D = O.DiffusionOracle(cases.DIFF)
        ref = D.representation_learning_loss(lambda x: O.encoder_forward(esd, "celeba64", x),
                                             lambda x, t, z: O.shiftunet_forward(dsd, c, x, t, z), x0, g["t"], g["noise"])
        ref.backward()
    finally:
        O.DROPOUT_MASKS = None
    assert_close(loss, ref, rtol=1e-4, atol=1e-7, what="loss with dropout")
    gmax = max(float(v.grad.abs().max()) for v in list(dsd.values()) + list(esd.values()) if v.grad is not None)
    for mod, sd in ((dec, dsd), (enc, esd)):
        for k, p in mod.named_parameters():
            if sd[k].grad is None:
                continue
            err = float((p.grad.cpu() - sd[k].grad).abs().max())
            assert rel_l2(p.grad, sd[k].grad) < 2e-3 or err < 2e-3 * float(sd[k].grad.abs().max()) + 2e-4 * gmax, k


def _latent_setup(cfg, dropout=0.0):
    from pdae_b200.model.mlp_skip_net import MLPSkipNet
    from pdae_b200.utils.synth import fill_module_
    c = dict(cfg["cfg"], dropout=dropout)
    mlp = fill_module_(MLPSkipNet(**c), seed=9)
    sd = {k: v.requires_grad_(True) for k, v in cases.sd_of(mlp).items() if ".cond_layers." not in k}
    return c, mlp, sd


def _latent_loss(gd, mlp, z0, t, noise):
    lc = gd.latent_diffusion_config
    z_t = gd.extract_coef_at_t(lc["sqrt_alphas_cumprod"], t, z0.shape) * z0 + \
        gd.extract_coef_at_t(lc["sqrt_one_minus_alphas_cumprod"], t, z0.shape) * noise
    return gd.p_loss(noise, mlp(z_t, t), loss_type=lc["loss_type"])


def test_latent_training_step_matches_reference_and_oracle():
    """latent_diffusion_train_one_batch (gaussian_diffusion.py:373-398): L1 loss and the gradient of EVERY MLPSkipNet
    parameter from the hand-written backward vs the reference fixture and vs oracle autograd."""
    from pdae_b200.diffusion.gaussian_diffusion import GaussianDiffusion
    cfg, g = load_golden("train_latent")
    c, mlp, sd = _latent_setup(cfg)
    D = O.DiffusionOracle(cases.DIFF)
    ref = D.latent_diffusion_loss(lambda z, t: O.mlp_skip_net_forward(sd, c, z, t), g["z0"], g["t"], g["noise"])
    ref.backward()
    mlp = mlp.cuda().train()
    mlp.precision = "fp32"
    gd = GaussianDiffusion(cases.DIFF, torch.device("cuda"))
    loss = _latent_loss(gd, mlp, g["z0"].cuda(), g["t"].cuda(), g["noise"].cuda())
    assert_close(loss, g["loss"], rtol=1e-5, atol=1e-7, what="latent loss vs reference fixture")
    loss.backward()
    named = dict(mlp.named_parameters())
    assert sum(1 for p in named.values() if p.grad is not None) == cfg["n_params_with_grad"]
    for k, p in named.items():
        assert rel_l2(p.grad, sd[k].grad) < 1e-4, (k, rel_l2(p.grad, sd[k].grad))
    for k in ("time_embed.0.weight", "layers.1.linear_emb.weight", "layers.2.norm.weight", "layers.4.linear.bias"):
        kk = k.replace(".", "_")
        assert_close(named[k].grad.flatten()[:512], g["g_" + kk], rtol=1e-3, atol=1e-7, what=k + " vs reference")
    # second step through the cached plans after an in-place weight update: gradients change, stay finite
    with torch.no_grad():
        for p in mlp.parameters():
            p.add_(p.grad, alpha=-1e-2)
            p.grad = None
    loss2 = _latent_loss(gd, mlp, g["z0"].cuda(), g["t"].cuda(), g["noise"].cuda())
    loss2.backward()
    assert float(loss2) < float(loss) and all(torch.isfinite(p.grad).all() for p in mlp.parameters())


def test_latent_training_with_dropout_and_full_api_call():
    from pdae_b200.diffusion.gaussian_diffusion import GaussianDiffusion
    from pdae_b200.utils.synth import synth_images, synth_normal
    cfg, g = load_golden("train_latent")
    c, mlp, sd = _latent_setup(cfg, dropout=0.1)
    mlp = mlp.cuda().train()
    mlp.precision = "fp32"
    gd = GaussianDiffusion(cases.DIFF, torch.device("cuda"))
    torch.manual_seed(3)
    loss = _latent_loss(gd, mlp, g["z0"].cuda(), g["t"].cuda(), g["noise"].cuda())
    loss.backward()
    trainer = list(mlp._train_cache.values())[0]
    names = {id(m): n for n, m in mlp.named_modules()}
    masks = {names[id(layer)]: mk.tensor.cpu().clone() for layer, mk, p in trainer.fwd.dropout_masks}
    assert len(masks) == c["num_layers"] - 1 and all(0.8 < float(m.mean()) < 0.98 for m in masks.values())
    O.DROPOUT_MASKS = dict(masks, p=0.1)
    try:
        ref = O.DiffusionOracle(cases.DIFF).latent_diffusion_loss(lambda z, t: O.mlp_skip_net_forward(sd, c, z, t), g["z0"], g["t"],
                                                                  g["noise"])
        ref.backward()
    finally:
        O.DROPOUT_MASKS = None
    assert_close(loss, ref, rtol=1e-5, atol=1e-7, what="latent loss with dropout")
    for k, p in mlp.named_parameters():
        assert rel_l2(p.grad, sd[k].grad) < 1e-4, (k, rel_l2(p.grad, sd[k].grad))
    # the reference-facing call: frozen encoder in eval mode, trainable latent net
    enc, _ = cases.model_case({"kind": "encoder", "size": 64})
    enc = enc.cuda().requires_grad_(False).eval()
    enc.precision = "fp32"
    for p in mlp.parameters():
        p.grad = None
    mean, std = (synth_normal((1, 512), 34) * 0.1).cuda(), (synth_normal((1, 512), 35).abs() + 0.5).cuda()
    out = gd.latent_diffusion_train_one_batch(mlp, enc, synth_images(3, 3, 64, 33).cuda(), mean, std)["prediction_loss"]
    out.backward()
    assert torch.isfinite(out) and all(p.grad is not None and torch.isfinite(p.grad).all() for p in mlp.parameters())
