"""The C-ABI launch-plan executor (pdae_plan_*, include/pdae_b200.h) on the host: the generated trampoline table is in step
with pdae_b200/_native.py, recording validates entry names / arity / stream slots, and argument values of every class
(pointer, int, int64, float) reach the entry point in order -- checked through entry points that reject their arguments
BEFORE touching a device, so no GPU is needed."""
import ctypes
import subprocess
import sys
from pathlib import Path

import pytest

from pdae_b200 import _native

ROOT = Path(__file__).resolve().parents[1]


def _plan():
    L = _native.lib()
    h = ctypes.c_void_p()
    _native.check(L.pdae_plan_create(ctypes.byref(h)), "pdae_plan_create")
    return L, h


def test_generated_table_is_up_to_date():
    r = subprocess.run([sys.executable, str(ROOT / "scripts" / "gen_plan_exec.py"), "--check"])
    assert r.returncode == 0, "run scripts/gen_plan_exec.py and rebuild: csrc/plan_exec_table.inc is stale"


def test_recording_validates_entry_arity_and_stream_slot():
    L, h = _plan()
    try:
        fn = L.pdae_colsum
        blob = _native.pack_args(fn, [None, ctypes.c_int64(5), 4, None, None])
        assert L.pdae_plan_add(h, b"pdae_colsum", blob, 5, 4) == 0
        assert L.pdae_plan_size(h) == 1 and L.pdae_plan_op_name(h, 0) == b"pdae_colsum"
        assert L.pdae_plan_add(h, b"pdae_no_such_entry", blob, 5, 4) != 0
        assert b"not a recordable entry point" in L.pdae_last_error()
        assert L.pdae_plan_add(h, b"pdae_colsum", blob, 4, 3) != 0 and b"takes 5 arguments" in L.pdae_last_error()
        assert L.pdae_plan_add(h, b"pdae_colsum", blob, 5, 1) != 0 and b"bad stream slot" in L.pdae_last_error()   # slot 1 is an int64
        assert L.pdae_plan_add(h, b"pdae_conv_tc2_create", blob, 5, -1) != 0          # create functions are not recordable
        assert L.pdae_plan_size(h) == 1
        # the recorded call runs (and fails in the entry point's own argument check: null pointers)
        assert L.pdae_plan_run_step(h, None) != 0 and b"colsum: bad args" in L.pdae_last_error()
    finally:
        L.pdae_plan_destroy(h)


def test_argument_values_of_every_class_arrive_in_order():
    # pdae_adam_ema_step(table*, block_map*, n_blocks, chunk, lr, beta1, beta2, eps, weight_decay, step(int64), grad_scale,
    #                    ema_decay, stream): its checks name chunk and step in the error text
    L, h = _plan()
    try:
        fn = L.pdae_adam_ema_step
        dummy = ctypes.create_string_buffer(64)
        ptr = ctypes.c_void_p(ctypes.addressof(dummy))

        def record(chunk, step):
            args = [ptr, ptr, 3, chunk, ctypes.c_float(1e-3), ctypes.c_float(0.9), ctypes.c_float(0.999), ctypes.c_float(1e-8),
                    ctypes.c_float(0.0), ctypes.c_int64(step), ctypes.c_float(1.0), ctypes.c_float(0.9999), None]
            _native.check(L.pdae_plan_add(h, b"pdae_adam_ema_step", _native.pack_args(fn, args), len(args), 12), "add")

        record(6, 7)                  # int after two pointers and an int
        assert L.pdae_plan_run_step(h, None) != 0
        assert b"chunk=6 must be a positive multiple of 4" in L.pdae_last_error()
    finally:
        L.pdae_plan_destroy(h)
    L, h = _plan()
    try:
        args = [ptr, ptr, 3, 8, ctypes.c_float(1e-3), ctypes.c_float(0.9), ctypes.c_float(0.999), ctypes.c_float(1e-8),
                ctypes.c_float(0.0), ctypes.c_int64(-(2 ** 40)), ctypes.c_float(1.0), ctypes.c_float(0.9999), None]
        _native.check(L.pdae_plan_add(h, b"pdae_adam_ema_step", _native.pack_args(fn, args), len(args), 12), "add")
        assert L.pdae_plan_run_step(h, None) != 0         # int64 after five floats, beyond 32 bits
        assert f"step={-(2 ** 40)} must be >= 1".encode() in L.pdae_last_error()
    finally:
        L.pdae_plan_destroy(h)


def test_pack_args_rejects_unrecordable_types():
    L = _native.lib()
    with pytest.raises(_native.NativeError):
        _native.pack_args(L.pdae_plan_create, [None])
