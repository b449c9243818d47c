#!/usr/bin/env python
"""DDIM-100 autoencoding throughput (BASELINE.json metric) for the pdae_b200 hot path.

  python bench.py [--gpus N] [--steps K] [--warmup W]            # our CUDA path, one process per GPU (torchrun for N>1)
  python bench.py --impl reference [...]                          # reference algorithm on the host CPU cores (oracle port)

One "step" = one full autoencoding pass over one synthetic batch: 1 semantic-encoder forward + S DDIM-encode steps + S
DDIM-decode steps of the ShiftUNet (S=100 -> 200 decoder forwards + 200 fused DDIM updates).  `value` = images/s with
the batch resident in HBM; `e2e` = the same pass driven through the public API
(GaussianDiffusion.representation_learning_autoencoding) from pinned host memory, H2D and D2H inside the timed region.
Rank 0 prints ONE JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from tests.configs import CELEBA64_PROXY, DIFFUSION, FFHQ128_PROXY, FFHQ256_PROXY  # noqa: E402

WORKLOADS = {
    # name: (decoder cfg, image size, encoder kind, encoder input size, default per-GPU batch, GFLOP per decoder image-step)
    "celeba64": (CELEBA64_PROXY, 64, "celeba64", 64, 256, 48.11),
    "ffhq128": (FFHQ128_PROXY, 128, "ffhq128", 128, 64, 258.40),
    "ffhq256": (FFHQ256_PROXY, 256, "ffhq128", 128, 8, 967.20),
}
ENC_GFLOP = {"celeba64": 0.134, "ffhq128": 0.616}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="celeba64", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch (0 = workload default)")
    ap.add_argument("--ddim-steps", type=int, default=100)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32", "bf16x3"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-profile", action="store_true")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.index, self.rows, self.stop = index, [], threading.Event()
        self.th = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self.stop.wait(0.2)

    def __enter__(self):
        self.th.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.th.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for j, n in enumerate(names) if any(r[3 + j].lower().startswith("active") for r in self.rows if len(r) > 3 + j)]
        mx = float(self.rows[0][1]) if self.rows[0][1].replace(".", "").isdigit() else None
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": reasons, "samples": len(self.rows)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1427.1), d.get("hbm_gbs", 6575.1), "measured (MEASURED_PEAKS.json, sustained bf16)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


def oracle_models(cfg, enc_kind, latent_dim=512, seed=0):
    """CPU oracle callables with synthetic weights (shapes/keys from the product modules, values from utils.synth)."""
    from oracle import pdae_oracle as O
    from pdae_b200.model.representation_learning.encoder import CELEBA64Encoder, FFHQEncoder
    from pdae_b200.model.shift_unet import ShiftUNet
    from pdae_b200.utils.synth import fill_module_
    c = dict(cfg, latent_dim=latent_dim)
    dsd = {k: v.clone() for k, v in fill_module_(ShiftUNet(**c), seed=seed).state_dict().items()}
    Enc = CELEBA64Encoder if enc_kind == "celeba64" else FFHQEncoder
    esd = {k: v.clone() for k, v in fill_module_(Enc(latent_dim=latent_dim), seed=seed + 1).state_dict().items()}
    return (lambda x: O.encoder_forward(esd, enc_kind, x)), (lambda x, t, z: O.shiftunet_forward(dsd, c, x, t, z)), O


def cpu_sample(cfg, size, enc_kind, enc_size, S, batch, n_steps, warm=1):
    """Time `n_steps` oracle DDIM steps (ShiftUNet forward + update) on the host cores; images/s extrapolated to 2*S steps
    + 1 encoder forward per image (every step is identical work)."""
    from pdae_b200.utils.synth import synth_images, synth_normal
    enc, dec, O = oracle_models(cfg, enc_kind)
    from pdae_b200.utils.host import host_cores
    cores = host_cores()  # affinity and cgroup quota: the box exposes 128 logical CPUs but grants a 16-CPU quota
    torch.set_num_threads(cores)
    D = O.DiffusionOracle(DIFFUSION)
    tabs, tmap, _ = D._ddim(f"ddim{S}")
    x = synth_normal((batch, 3, size, size), 5)
    x0e = synth_images(batch, 3, enc_size, 6)
    with torch.inference_mode():
        t0 = time.perf_counter()
        z = enc(x0e)
        t_enc = time.perf_counter() - t0
        t = torch.full((batch,), S // 2, dtype=torch.long)
        for _ in range(warm):
            eps, grad = dec(x, tmap[t], z)
        t0 = time.perf_counter()
        for _ in range(n_steps):
            eps, grad = dec(x, tmap[t], z)
            x = O.ddim_update(tabs, x, t, eps, grad, "sample")
        t_step = (time.perf_counter() - t0) / n_steps
    ips = batch / (2 * S * t_step + t_enc)
    return ips, cores, t_step, t_enc


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg, size, enc_kind, enc_size, _, _ = WORKLOADS[args.workload]
    S = args.ddim_steps
    b = 8 if size <= 64 else 2
    vals = []
    for i in range(args.warmup + args.steps):
        ips, cores, t_step, t_enc = cpu_sample(cfg, size, enc_kind, enc_size, S, b, n_steps=1, warm=1 if i == 0 else 0)
        if i >= args.warmup:
            vals.append((ips, t_step))
    ips = sum(v[0] for v in vals) / len(vals)
    ms = 1e3 * sum(v[1] for v in vals) / len(vals)
    sample = f"{b} images x 1 ShiftUNet DDIM step per bench step, extrapolated to {2 * S} steps + 1 encoder forward per image"
    line = {"impl": "reference", "metric": "ddim100_autoencoding_images_per_sec", "value": ips, "unit": "images/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.workload}-proxy ShiftUNet+encoder, DDIM-{S} encode + DDIM-{S} decode", "batch": b},
            "cpu_baseline": {"value": ips, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    import pdae_b200
    from pdae_b200.diffusion.gaussian_diffusion import GaussianDiffusion
    from pdae_b200.model.representation_learning.encoder import CELEBA64Encoder, FFHQEncoder
    from pdae_b200.model.shift_unet import ShiftUNet
    from pdae_b200.utils.synth import fill_module_, synth_images

    cfg, size, enc_kind, enc_size, dflt_b, gflop_step = WORKLOADS[args.workload]
    B = args.batch or dflt_b
    S = args.ddim_steps
    pdae_b200.set_default_precision(args.precision)
    dec = fill_module_(ShiftUNet(latent_dim=512, **cfg), seed=0).eval().to(dev)
    Enc = CELEBA64Encoder if enc_kind == "celeba64" else FFHQEncoder
    enc = fill_module_(Enc(latent_dim=512), seed=1).eval().to(dev)
    gd = GaussianDiffusion(DIFFUSION, dev)
    style = f"ddim{S}"

    x_host = synth_images(B, 3, size, 100 + rank).pin_memory()
    out_host = torch.empty_like(x_host).pin_memory()
    x_dev = x_host.to(dev)
    gather = torch.empty(world * B, 3, size, size, device=dev) if world > 1 else None

    def enc_input(x):
        return x if enc_size == size else torch.nn.functional.avg_pool2d(x, size // enc_size)

    def autoencode(x):
        with torch.inference_mode():
            z = enc(enc_input(x))
            x_T = gd.representation_learning_ddim_encode(style, None, dec, x, z)
            rec = gd.representation_learning_ddim_sample(style, None, dec, None, x_T, z)
            if world > 1:
                import torch.distributed as dist
                dist.all_gather_into_tensor(gather, rec)  # the single collective of the sampling path
            return rec

    def timed(fn, k):
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            import torch.distributed as dist
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            dist.barrier()
        return float(ms.item())

    for _ in range(max(args.warmup, 3) if args.warmup >= 0 else 0):
        autoencode(x_dev)
    with ClockSampler(local) as clk:
        ms_total = timed(lambda: autoencode(x_dev), args.steps)
    ms_step = ms_total / args.steps
    value = world * B / (ms_step / 1e3)

    def e2e_once():
        xd = x_host.to(dev, non_blocking=True)
        rec = autoencode(xd)
        out_host.copy_(rec, non_blocking=True)

    e2e_once()
    ms_e2e = timed(e2e_once, args.steps) / args.steps
    e2e_val = world * B / (ms_e2e / 1e3)

    # kernel-level view of ONE decoder step (CUDA events around every launch of the step plan)
    plan, _ = dec.plan_for(B, size, size)
    launches_per_step = plan.n_launch + 1  # + fused DDIM update
    enc_plan = list(enc._plans().values())[0][0]
    gpu_launches = args.steps * (2 * S * launches_per_step + enc_plan.n_launch)
    roof = None
    kinds = {}
    if not args.no_profile:
        prof = plan.profile(reps=3)
        tot = sum(v["ms"] for v in prof.values())
        kinds = {k: {"ms": round(v["ms"], 4), "share": round(v["ms"] / tot, 4), "launches": v["launches"],
                     "tflops": round(v["flops"] / (v["ms"] * 1e9), 2) if v["flops"] and v["ms"] > 0 else None}
                 for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])}
        peak_tf, peak_bw, peak_src = peaks()
        dom = max((k for k in prof if k.startswith("conv")), key=lambda k: prof[k]["ms"])
        n_l = prof[dom]["launches"]
        ach = prof[dom]["flops"] / (prof[dom]["ms"] * 1e9)
        kname = {"conv_tc2": "pdae::conv_tc2_kernel", "conv_tc": "pdae::conv_tc_kernel"}.get(dom, "pdae::conv_simt_kernel")
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "conv_tc2_traffic.json")
        if dom == "conv_tc2" and os.path.exists(tpath):
            tj = json.load(open(tpath))
            if tj.get("workload") == args.workload and tj.get("batch") == B:   # ncu dram bytes, per launch like `achieved`
                traffic = tj["traffic_bytes_per_launch"]
        roof = {"bound": "tensor", "kernel": kname,
                "achieved": round(ach, 2), "peak": peak_tf, "unit": "TFLOP/s", "frac": round(ach / peak_tf, 4), "traffic": traffic, "traffic_unit": "bytes/launch (ncu dram read+write, profiles/conv_tc2_traffic.json)",
                "peak_source": peak_src, "launches_per_decoder_step": n_l,
                "flops_per_launch_avg": prof[dom]["flops"] / n_l, "ms_per_launch_avg": prof[dom]["ms"] / n_l,
                "step_ms_sum_of_kernels": round(tot, 3)}

    if rank != 0:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()
        return

    peak_tf, _, peak_src = peaks()
    flop_img = (2 * S * gflop_step + ENC_GFLOP[enc_kind]) * 1e9
    line = {
        "metric": "ddim100_autoencoding_images_per_sec", "value": round(value, 4), "unit": "images/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": round(ms_step, 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": {"bf16": "bf16", "bf16x3": "bf16x3 (split-operand bf16 MMAs, fp32-grade)", "fp32": "f32"}[args.precision], "data": "synthetic",
        "config": {"workload": f"{args.workload}-proxy ShiftUNet+encoder (proxy decoder config, SURVEY D4), {size}x{size}x3, "
                               f"DDIM-{S} encode + DDIM-{S} decode", "batch_per_gpu": B, "global_batch": world * B,
                   "parallelism": f"dp{world} batch-sharded, one all-gather of results", "precision": args.precision,
                   "l2": "activations per step exceed L2 (inputs larger than L2)"},
        "e2e": {"value": round(e2e_val, 4), "unit": "images/s", "h2d_bytes_per_step": x_host.numel() * 4,
                "d2h_bytes_per_step": out_host.numel() * 4, "ms_per_step": round(ms_e2e, 3)},
        "gpu_launches": gpu_launches,
        "clocks": clk.summary(),
        "model_flops_utilization": {"algorithmic_tflops": round(value * flop_img / 1e12 / world, 2), "peak_tflops": peak_tf,
                                    "frac": round(value * flop_img / 1e12 / world / peak_tf, 4), "peak_source": peak_src},
        "roofline": roof, "kernels_per_decoder_step": kinds,
    }
    if not args.no_cpu_baseline and world == 1:   # the CPU baseline and the parity probe are an N=1, rank-0 leg
        b = 8 if size <= 64 else 2
        ips, cores, t_step, t_enc = cpu_sample(cfg, size, enc_kind, enc_size, S, b, n_steps=3)
        line["cpu_baseline"] = {"value": ips, "unit": "images/s", "cores": cores, "kind": "port",
                                "sample": f"oracle (torch-CPU restatement of the reference), batch {b}: 1 warm-up + 3 timed "
                                          f"ShiftUNet DDIM steps ({t_step:.2f} s/step) + 1 encoder forward, extrapolated to "
                                          f"{2 * S} steps per image"}
        line["cpu_baseline"]["parity"] = parity_probe(gd, enc, dec, cfg, size, enc_kind, enc_size, enc_input, dev)
    print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def parity_probe(gd, enc, dec, cfg, size, enc_kind, enc_size, enc_input, dev):
    """The oracle used as the CHECKER on a bounded sample of the bench workload: a short DDIM autoencoding of a few images
    through the CPU oracle and through this package (fp32 parity mode, the split-operand tensor-core mode "bf16x3" and the
    bf16 mode the bench times), same weights and
    inputs.  Reports the reconstruction MSE of each (BASELINE.json: "recon MSE vs ref") and the relative L2 distance of
    the reconstructions.  Random-init weights make the sampling chain chaotic (tests/test_oracle_golden.py::
    test_loop_sensitivity), so the bf16 figure measures amplified rounding noise, not image quality."""
    from pdae_b200.utils.synth import synth_images
    n, s = (2, 10) if size <= 64 else ((1, 5) if size <= 128 else (1, 3))
    style = f"ddim{s}"
    x0 = synth_images(n, 3, size, 4242)
    enc_o, dec_o, O = oracle_models(cfg, enc_kind)
    D = O.DiffusionOracle(DIFFUSION)
    with torch.inference_mode():
        z = enc_o(enc_input(x0))
        ref = D.representation_learning_ddim_sample(style, dec_o, D.representation_learning_ddim_encode(style, dec_o, x0, z), z)
    out = {"sample": f"{n} image(s), {style} encode + {style} decode, same synthetic weights/inputs as the oracle",
           "recon_mse_reference": float(((ref - x0) ** 2).mean())}
    prev = (enc.precision if hasattr(enc, "precision") else None, dec.precision if hasattr(dec, "precision") else None)
    for prec in ("fp32", "bf16x3", "bf16"):
        enc.precision = dec.precision = prec
        with torch.inference_mode():
            xd = x0.to(dev)
            zz = enc(enc_input(xd))
            rec = gd.representation_learning_ddim_sample(style, None, dec, None,
                                                         gd.representation_learning_ddim_encode(style, None, dec, xd, zz), zz).cpu()
        out[f"recon_mse_{prec}"] = float(((rec - x0) ** 2).mean())
        out[f"rel_l2_vs_reference_{prec}"] = float((rec - ref).norm() / ref.norm())
    enc.precision, dec.precision = prev
    return out


if __name__ == "__main__":
    main()
