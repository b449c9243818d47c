"""Time / profile gn_apply in isolation.  usage: python scripts/apply_bench.py B H W C1 C2 src_bf16(0/1) raw(0/1)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pdae_b200.engine import Plan, RESAMPLE_NONE
B, H, W, C1, C2, sbf, raw = (int(a) for a in sys.argv[1:8])
dev = torch.device("cuda")
plans = []
for i in range(4):
    P = Plan(dev, "bf16")
    dt = torch.bfloat16 if sbf else torch.float32
    s1 = P.fixed(torch.randn(B, H, W, C1, device=dev).to(dt))
    s2 = P.fixed(torch.randn(B, H, W, C2, device=dev).to(dt)) if C2 else None
    ab = P.fixed(torch.randn(B, 2, C1 + C2, device=dev))
    act, r = P.gn_apply(s1, C1, s2, C2, ab, silu=True, resample=RESAMPLE_NONE, B=B, H=H, W=W, act_dtype=torch.bfloat16,
                        raw_dtype=torch.bfloat16 if raw else None)
    act.keep = True
    if r is not None:
        r.keep = True
    P.finalize()
    plans.append(P)
for P in plans:
    P.run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    for P in plans:
        P.run()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
C = C1 + C2
byts = B * H * W * C * ((2 if sbf else 4) + 2 + (2 if raw else 0))
print(f"gn_apply {B}x{H}x{W} C={C1}+{C2} src_bf16={sbf} raw={raw}: {ms*1e3:.1f} us {byts/ms/1e6:.0f} GB/s")
