"""Shape predicates of the tensor-core kernels are pure host code in the C-ABI: callable without a GPU (no compute)."""
from pdae_b200 import _native


def test_wgrad_tc_supported_shapes():
    L = _native.lib()
    ok = L.pdae_wgrad_tc_supported
    assert ok(64, 64, 128, 64, 3) and ok(32, 32, 256, 128, 3) and ok(8, 8, 1024, 512, 3) and ok(16, 16, 256, 256, 1)
    assert ok(64, 64, 64, 64, 3) and ok(32, 32, 192, 64, 3)          # tap-pair mode: neither channel count a multiple of 128
    assert ok(4, 4, 128, 128, 3)                                        # 4x4 images: four per 64-pixel box
    assert not ok(32, 32, 96, 64, 3) and not ok(32, 32, 64, 3, 3)      # channel counts must be multiples of 64
    assert not ok(32, 32, 128, 128, 5) and not ok(32, 32, 128, 128, 2)  # 1x1 and 3x3 only
    assert ok(6, 6, 128, 128, 3)                                        # odd sizes: 2 x 2 pixels x 16 images per box


def test_conv_tc3_supported_shapes():
    ok = _native.lib().pdae_conv_tc3_supported
    assert ok(64, 64, 64, 64) and ok(16, 16, 768, 256) and ok(32, 8, 128, 128)
    assert not ok(8, 8, 512, 512)         # one 16 x 8 output tile needs H % 16 == 0 (the 8x8 level stays on conv_tc2)
    assert not ok(64, 64, 96, 64) and not ok(64, 64, 64, 96) and not ok(64, 12, 64, 64)
