"""CPU: the C-ABI library loads and exports what include/fortattack.h declares; no compute without a GPU."""
import ctypes
import os
import re

import pytest

import fortattack_b200 as fab
from fortattack_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_are_exported():
    hdr = open(os.path.join(ROOT, "include", "fortattack.h")).read()
    declared = set(re.findall(r"^(?:int|const char \*)\s*\*?(fa_\w+)\(", hdr, re.M))
    assert declared == set(_capi.SYMBOLS), declared ^ set(_capi.SYMBOLS)
    L = _capi.lib()
    for name in declared:
        assert hasattr(L, name)
    assert L.fa_abi_version() == int(re.search(r"#define FA_ABI_VERSION (\d+)", hdr).group(1))


def test_policy_and_rollout_headers_are_exported():
    """Every entry point of include/fortattack_policy.h and include/fortattack_rollout.h is in the library, and
    they reject bad arguments without a GPU."""
    L = _capi.lib()
    for hdr_name, prefix in (("fortattack_policy.h", "mp_"), ("fortattack_rollout.h", "rl_"), ("mape_world.h", "mw_"),
                             ("fortattack_render.h", "fr_"), ("fortattack_train.h", "tg_")):
        hdr = open(os.path.join(ROOT, "include", hdr_name)).read()
        declared = set(re.findall(r"^int\s+(%s\w+)\(" % prefix, hdr, re.M))
        assert declared, hdr_name
        for name in declared:
            assert hasattr(L, name), name
            # prototypes live in ONE place (_capi.lib): an unbound function would get its pointers truncated to C ints
            assert getattr(L, name).argtypes is not None, name + " has no ctypes prototype in _capi.lib()"
    L.mp_forward.restype = ctypes.c_int
    assert L.mp_forward(*([None] * 3), 3, 3, 8, 0, 0, 0, None, 0, *([None] * 7), None, 0, None, None, None, None) == -1
    assert b"NULL" in L.fa_last_error()
    assert L.rl_gae(None, None, None, None, None, None, 4, 2, 8, 0.99, 0.95, None) == -1
    assert L.tg_linear(None, 128, 1000, 128, None, 128, None, 0, 0, None, 128, None, None) == -1
    assert L.tg_packed_bytes(128, 128) == 4 * 128 * 128 + 4 * 128 and L.tg_packed_bytes(8, 6) == 4 * 16 * 64 + 4 * 16
    blob_bytes = int(re.search(r"#define MP_BLOB_F16_BYTES (\d+)", open(os.path.join(ROOT, "include", "fortattack_policy.h")).read()).group(1))
    from importlib import import_module
    pk = import_module("emergent-multiagent-strategies_b200.policy_kernel")
    assert pk.BLOB_F16_BYTES == blob_bytes


def test_config_validation_and_workspace_size():
    L = _capi.lib()
    n = ctypes.c_size_t()
    cfg = _capi.FaConfig(4096, 3, 3, 100, _capi.FA_F32, 0, 0, 0)
    assert L.fa_workspace_bytes(ctypes.byref(cfg), ctypes.byref(n)) == 0
    state = 4096 * 6 * 28 + 4096 * 8
    assert state <= n.value <= 4 * state + 16 * 256          # state + host-step staging + padding
    for bad in (_capi.FaConfig(4096, 6, 3, 100, 0, 0, 0, 0), _capi.FaConfig(0, 3, 3, 100, 0, 0, 0, 0),
                _capi.FaConfig(8, 3, 3, 0, 0, 0, 0, 0), _capi.FaConfig(8, 3, 3, 10, 7, 0, 0, 0)):
        assert L.fa_workspace_bytes(ctypes.byref(bad), ctypes.byref(n)) == -1
        assert L.fa_last_error()
    assert L.fa_step(None, None, None, None, None, None, 0, None) == -1


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(fab.FaError):
        fab.FortAttackBatch(4, 3, 3, device="cpu")
    # fa_create itself refuses without a device
    L = _capi.lib()
    cfg = _capi.FaConfig(4, 3, 3, 100, 0, 0, 0, 0)
    buf = ctypes.create_string_buffer(1 << 16)
    addr = (ctypes.addressof(buf) + 255) // 256 * 256
    h = ctypes.c_void_p()
    assert L.fa_create(ctypes.byref(cfg), addr, ctypes.byref(h)) == -3
    assert b"no CUDA device" in L.fa_last_error()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "emergent-multiagent-strategies_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(d, f)).read()
                assert not any(w in src for w in ("fa_oracle", "mw_oracle", "render_oracle", "oracle/")), os.path.join(d, f)
