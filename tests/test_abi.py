"""CPU-only checks of the drop-in boundary: the C-ABI library loads without a GPU, exports
every symbol include/pkb200.h declares, and fails loudly (no CPU fallback) without a device."""

import ctypes
import os
import re

import numpy as np
import pytest

import pocketkaldi_b200 as pk
from pocketkaldi_b200 import binding

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    text = open(os.path.join(ROOT, "include", "pkb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pkb_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_bound_and_exported():
    lib = pk.load_library()
    declared = header_functions()
    assert len(declared) >= 40
    assert sorted(binding.EXPORTED_SYMBOLS) == declared
    for name in declared:
        assert hasattr(lib, name), name


def test_num_frames_matches_reference_rule():
    lib = pk.load_library()
    for n, t in ((0, 0), (399, 0), (400, 1), (559, 1), (560, 2), (160000, 998)):
        assert lib.pkb_fbank_num_frames(n) == t


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(pk.PkbError) as e:
        pk.Context(0)
    assert e.value.code == 4  # PKB_ERR_CUDA
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "pocketkaldi_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cc", ".cpp")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert "libpkoracle" not in src and "libpkref" not in src, f
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f


def test_loglik16_expand_is_host_only_and_exact():
    # pkb_loglik16_expand finishes the compact rows on the host (no GPU involved):
    # out = prob_scale * (half(h) + off[frame])
    import ctypes as C
    import numpy as np
    from pocketkaldi_b200 import binding
    lib = binding.load_library()
    rng = np.random.default_rng(0)
    h = rng.uniform(-40, 0, (37, 129)).astype(np.float16)
    h[0, :4] = [0.0, -0.0, -6.1e-5, -65504.0]          # zero, signed zero, subnormal, largest half
    off = rng.uniform(-9, -7, 37).astype(np.float32)
    out = np.empty(h.shape, np.float32)
    rc = lib.pkb_loglik16_expand(h.view(np.uint16).ctypes.data, off.ctypes.data, 37, 129, C.c_float(0.1),
                                 out.ctypes.data)
    assert rc == 0
    want = (h.astype(np.float32) + off[:, None]) * np.float32(0.1)
    assert np.array_equal(out, want)
    assert lib.pkb_loglik16_expand(None, None, 0, 129, C.c_float(1.0), None) == 0   # empty block
