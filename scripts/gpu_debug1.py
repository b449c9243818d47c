"""GPU debugging aid (not a test): stage-wise comparison of the latent MLP and the shift DDIM loop vs the CPU oracle."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from oracle import pdae_oracle as O
from tests import cases
from tests.util import load_golden, rel_l2
from pdae_b200 import _native
from pdae_b200.utils.synth import synth_normal, synth_images

print("cpus: os.cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
try:
    print("cgroup cpu.max:", open("/sys/fs/cgroup/cpu.max").read().strip())
except Exception as e:
    print("cgroup:", e)
dev = torch.device("cuda")
L = _native.lib()
st = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
P = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None

# --- mlp row op in isolation
h = torch.randn(2, 128, device=dev); c = torch.randn(2, 128, device=dev); w = torch.randn(128, device=dev); b = torch.randn(128, device=dev)
out = torch.zeros(2, 192, device=dev)
_native.check(L.pdae_mlp_mod_ln_act(P(h), P(c), P(w), P(b), ctypes.c_float(1e-5), 1, P(out), 192, 2, 128, st()))
ref = F.silu(F.layer_norm(h * (1 + c), (128,), w, b, 1e-5))
print("mlp_mod_ln_act err", float((out[:, :128] - ref).abs().max()), "tail untouched", float(out[:, 128:].abs().max()))
src = torch.randn(2, 64, device=dev)
_native.check(L.pdae_copy_cols(P(src), P(out), 192, 128, 2, 64, st()))
print("copy_cols err", float((out[:, 128:] - src).abs().max()))

cfg, g = load_golden("model_mlp_skip")
m, inp = cases.model_case(cfg)
sd = cases.sd_of(m)
m = m.cuda(); m.precision = "fp32"
with torch.no_grad():
    y = m(inp["x"].cuda(), g["t"].cuda())
print("mlp full rel", rel_l2(y, g["y"]))
plan, (x_in, t_in, outb) = list(m._plans().values())[0]
for i, (fn, args) in enumerate(plan.ops):
    print(i, fn, [getattr(a, "name", None) or (type(a).__name__ if not isinstance(a, (int, float)) else a) for a in args][:12])

# --- shift loop step by step
from pdae_b200.diffusion.gaussian_diffusion import GaussianDiffusion
gd = GaussianDiffusion(cases.DIFF, dev)
cfg, g = load_golden("loop_shift_ddim10")
m, _ = cases.model_case({"kind": "shiftunet", "cfg": cfg["cfg"], "size": 16})
sd = cases.sd_of(m)
m = m.cuda(); m.precision = "fp32"
z = synth_normal((2, 64), 27); xT = synth_normal((2, 3, 16, 16), 25)
D = O.DiffusionOracle(cases.DIFF)
tabs, tmap, S = D._ddim("ddim10")
dd = gd._ddim("ddim10")
x_ref = xT.clone(); x_gpu = xT.cuda()
with torch.no_grad():
    for i in reversed(range(1, S + 1)):
        t = torch.full((2,), i, dtype=torch.long)
        e_r, g_r = O.shiftunet_forward(sd, cfg["cfg"], x_ref, tmap[t], z)
        e_g, g_g = m(x_ref.cuda(), tmap[t].cuda(), z.cuda())   # same input as the oracle at every step
        x_new = O.ddim_update(tabs, x_ref, t, e_r, g_r, "sample")
        x_gpu_new = dd._update(x_ref.cuda(), t.cuda(), e_g, g_g, "sample")
        print(f"step {i}: t={int(tmap[i])} eps rel {rel_l2(e_g, e_r):.2e} grad rel {rel_l2(g_g, g_r):.2e} upd max {float((x_gpu_new.cpu()-x_new).abs().max()):.2e} |eps|max {float(e_r.abs().max()):.2f}")
        x_ref = x_new
    for graph in (False, True):
        type(dd).use_cuda_graph = graph
        fast = gd.representation_learning_ddim_sample("ddim10", None, m, None, xT.cuda(), z.cuda())
        print("graph", graph, "fast loop vs golden rel", rel_l2(fast, g["sample"]), "max", float((fast.cpu() - g["sample"]).abs().max()))
    slow = dd._loop(lambda a, b, c: m(a, b, c), xT.cuda(), z.cuda(), "sample", shift=True)
    print("slow loop vs golden rel", rel_l2(slow, g["sample"]))
