"""The C-ABI library builds for sm_100a, loads, and exports every symbol include/pdae_b200.h declares (no GPU needed)."""
import ctypes
import os
import re

from pdae_b200 import _native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "pdae_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pdae_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported_and_bound():
    _native.build()
    lib = ctypes.CDLL(_native.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/pdae_b200.h but not exported"
    assert sorted(_native.EXPORTS) == syms, "ctypes bindings out of sync with the header"


def test_abi_version_and_error_string():
    L = _native.lib()
    assert L.pdae_abi_version() == 1
    # argument validation happens before any CUDA call, so it is testable without a device
    rc = L.pdae_conv3x3_smalln(None, 0, None, None, None, 1, 8, 8, 64, 3, None)
    assert rc == -1 and b"null pointer" in L.pdae_last_error()
    rc = L.pdae_gn_stats(ctypes.c_void_p(16), 30, None, 0, 1, 64, ctypes.c_void_p(16), None)
    assert rc == -1 and b"unsupported" in L.pdae_last_error()


def test_sass_is_blackwell_native():
    """tcgen05.mma -> UTCHMMA, TMA -> UTMALDG, tcgen05.ld -> LDTM must be present in the sm_100a cubin."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        import pytest
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", _native.LIB_PATH], capture_output=True, text=True).stdout
    for mnem in ("UTCHMMA", "UTMALDG", "LDTM"):
        assert mnem in sass, mnem
    assert "sm_100a" in sass
