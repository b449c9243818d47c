"""Backward CUDA-core kernels through the C-ABI vs float64 autograd on the device: conv data gradient (generic tile kernel, the
3-channel image-head kernel, the split-K wide-Linear kernel) and the bias-gradient column sum (vector and scalar paths)."""
import ctypes

import pytest
import torch
import torch.nn.functional as F

from pdae_b200 import _native

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


@pytest.mark.parametrize("B,H,W,Cin,Cout,k,stride", [
    (3, 16, 16, 64, 3, 3, 1),        # image head: 3x3 onto 3 channels (conv3x3_dgrad_smalln)
    (2, 9, 7, 32, 4, 3, 1),          # same kernel, odd image, 4 output channels
    (5, 1, 1, 96, 2048, 1, 1),       # wide Linear: split-K kernel, ragged batch block
    (33, 1, 1, 512, 1100, 1, 1),     # wide Linear: two batch blocks, ragged last weight chunk
    (2, 8, 8, 32, 48, 3, 2),         # generic tile kernel (stride 2)
])
def test_conv_dgrad_matches_float64_autograd(B, H, W, Cin, Cout, k, stride):
    g = torch.Generator().manual_seed(5)
    pad = k // 2
    Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    w = (torch.randn(Cout, Cin, k, k, generator=g) / (k * Cin ** 0.5)).to(DEV)
    dy = torch.randn(B, Ho, Wo, Cout, generator=g).to(DEV)
    wt = w.reshape(Cout, Cin, k * k).permute(2, 0, 1).contiguous()             # [k*k][Cout][Cin]
    dx = torch.full((B, H, W, Cin), float("nan"), device=DEV)
    L = _native.lib()
    _native.check(L.pdae_conv2d_dgrad_simt(_p(dy), _p(wt), _p(dx), B, H, W, Cin, Cout, k, stride, pad, 0, _stream()), "dgrad")
    torch.cuda.synchronize()
    x = torch.zeros(B, Cin, H, W, device=DEV, dtype=torch.float64, requires_grad=True)
    F.conv2d(x, w.double(), stride=stride, padding=pad).backward(dy.double().permute(0, 3, 1, 2))
    ref = x.grad.permute(0, 2, 3, 1)
    err = (dx.double() - ref).abs().max().item()
    assert err <= 2e-5 * ref.abs().max().item() + 1e-6, err


@pytest.mark.parametrize("M,N", [(4096, 128), (1000, 64), (37, 512), (70000, 256), (513, 3), (32, 2052)])
def test_colsum_matches_float64(M, N):
    g = torch.Generator().manual_seed(9)
    dy = torch.randn(M, N, generator=g).to(DEV)
    out = torch.zeros(N, device=DEV)
    _native.check(_native.lib().pdae_colsum(_p(dy), ctypes.c_int64(M), N, _p(out), _stream()), "colsum")
    torch.cuda.synchronize()
    ref = dy.double().sum(0)
    assert (out.double() - ref).abs().max().item() <= 1e-5 * (M ** 0.5) * 4 + 1e-6
