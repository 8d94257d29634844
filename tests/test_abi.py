"""CPU: the C-ABI shared library loads without a GPU and exports every symbol that
include/evfly_b200.h declares; the ctypes table in evfly_b200/_lib.py covers all of them."""
import ctypes
import os
import re

import pytest

from evfly_b200 import _build, _lib


@pytest.fixture(scope="module")
def lib():
    _build.build()
    return _lib.load()


def test_library_is_in_tree_and_loads(lib):
    assert os.path.dirname(_lib.LIB_PATH) == os.path.dirname(os.path.abspath(_lib.__file__))
    assert lib.evfly_abi_version() == 1
    assert lib.evfly_launch_count() >= 0


def test_every_declared_symbol_is_exported_and_bound(lib):
    declared = _lib.header_symbols()
    assert len(declared) >= 14
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(raw, name), f"{name} declared in evfly_b200.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes prototype in _lib.SIGNATURES"
    assert sorted(_lib.SIGNATURES) == declared


def test_prototype_arity_matches_header():
    text = re.sub(r"/\*.*?\*/", "", open(_lib.HEADER_PATH).read(), flags=re.S)
    for name, (_, args) in _lib.SIGNATURES.items():
        m = re.search(r"\b%s\s*\(([^;]*?)\)\s*;" % name, text, flags=re.S)
        assert m, name
        params = m.group(1).strip()
        n = 0 if params in ("", "void") else params.count(",") + 1
        assert n == len(args), f"{name}: header has {n} parameters, ctypes table {len(args)}"


def test_argument_errors_are_reported_not_raised(lib):
    # no GPU needed: validation happens before any CUDA call
    rc = lib.evfly_accumulate_counts(None, 10, 480, 640, None, None)
    assert rc == -1 and b"null" in lib.evfly_last_error()
    rc = lib.evfly_voxelize_window(None, 0, 480, 640, 5, 10, 10, None, None, None, 0, None)
    assert rc == -1 and b"empty window" in lib.evfly_last_error()
    with pytest.raises(_lib.EvflyError):
        _lib.check(rc, "evfly_voxelize_window")


def test_no_cpu_fallback_in_product_package():
    # the product package must not import the oracle (it would void every parity claim)
    pkg = os.path.dirname(os.path.abspath(_lib.__file__))
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f


def test_new_entry_points_validate_arguments_before_touching_the_gpu(lib):
    """Argument errors of the rows added late in the round (no GPU needed: validation precedes every CUDA call)."""
    rc = lib.evfly_counts_normalise(None, 4, 480, 640, 260, 346, 0.2, 0.97, -1.0, 1.0, 0.0, None, None, None)
    assert rc == -1 and b"null" in lib.evfly_last_error()
    rc = lib.evfly_counts_normalise(None, 4, 480, 640, 261, 346, 0.2, 0.97, -1.0, 1.0, 0.0, None, None, None)
    assert rc == -1 and b"even" in lib.evfly_last_error()
    rc = lib.evfly_remap_bicubic_f32(None, 0, 1, 480, 640, None, None, 640, 260, 346, 0, 0, None, None)
    assert rc == -1 and b"remap_bicubic_f32" in lib.evfly_last_error()
    rc = lib.evfly_tc_conv3x3_halo_bf16(None, None, None, None, 1, 20, 20, 20, 20, 48, 48, 1, None)
    assert rc == -1
    assert lib.evfly_accumulate_counts_binned_workspace_bytes(10_000_000, 480, 640) > 256 * 1024
    assert lib.evfly_accumulate_counts_binned_workspace_bytes(-1, 480, 640) == 0


def test_host_mirrors_refuse_cpu_tensors():
    """calibration_tools / pipeline run on the device only; a CPU tensor is an error, never a silent CPU path."""
    import numpy as np
    import torch
    from evfly_b200 import calibration_tools as CT
    if torch.cuda.is_available():
        pytest.skip("needs a machine without a GPU")
    with pytest.raises((_lib.EvflyError, AssertionError, RuntimeError)):
        CT.remap_bicubic(torch.zeros(1, 8, 8), torch.zeros(8, 8), torch.zeros(8, 8))
    with pytest.raises((_lib.EvflyError, AssertionError, RuntimeError)):
        CT.remap_img(np.zeros((8, 8), np.float32), (np.zeros((8, 8), np.float32), np.zeros((8, 8), np.float32)), False, False)
