"""Probe for next round's tap-reuse design: does a SWIZZLE_128B K-major UMMA shared-memory descriptor whose start address is
shifted by s rows (s x 128 B) read rows s.. of a TMA-written tile correctly, and what must the base_offset field be?
Runs a 1x1 conv over W=128-pixel rows (tile = one image row) normally, then with the A box loaded s pixels early and the
descriptor started s rows in: output pixels x < 128 - s must be identical.  usage: python scripts/desc_shift_probe.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from pdae_b200.engine import Plan

dev = torch.device("cuda")
B, H, W, Cin, Cout = 4, 8, 128, 128, 64
torch.manual_seed(0)
x = torch.randn(B, H, W, Cin, device=dev).to(torch.bfloat16)
w = torch.randn(Cout, Cin, 1, 1, device=dev) / Cin ** 0.5
bias = torch.zeros(Cout, device=dev)


def run(shift, boff):
    os.environ["PDAE_TC_DBG_SHIFT"] = str(shift)
    os.environ["PDAE_TC_DBG_BOFF"] = str(boff)
    P = Plan(dev, "bf16")
    out = P.new((B, H, W, Cout), torch.float32)
    out.keep = True
    P.conv(P.fixed(x), w, bias, out, B=B, H=H, W=W, Cin=Cin, Cout=Cout, k=1)
    P.finalize()
    P.run()
    torch.cuda.synchronize()
    return out.tensor.clone()


ref = run(0, 0)
for shift in (1, 2, 3, 4, 7, 8, 9, 16):
    for boff in sorted({0, shift & 7}):
        got = run(shift, boff)
        ok_rows = W - shift
        d = (got[:, :, :ok_rows] - ref[:, :, :ok_rows]).abs().max().item()
        # per-pixel-column check: which x positions are wrong
        bad = ((got[:, :, :ok_rows] - ref[:, :, :ok_rows]).abs().amax(dim=(0, 1, 3)) > 1e-3).nonzero().flatten().tolist()
        print(f"shift={shift:2d} base_offset={boff}: max|diff| over x<{ok_rows} = {d:.3e}  wrong x: {bad[:12]}{'...' if len(bad) > 12 else ''} ({len(bad)})")
os.environ["PDAE_TC_DBG_SHIFT"] = "0"
os.environ["PDAE_TC_DBG_BOFF"] = "0"
