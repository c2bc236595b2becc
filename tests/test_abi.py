"""The C-ABI libraries load and export every symbol their headers declare (no compute calls)."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT


def declared(header, prefix):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(%s[a-z0-9_]+)\s*\(" % prefix, text)))


def test_host_library_exports_header_symbols():
    lib = C.CDLL(os.path.join(ROOT, "minimaloptix_b200", "libmox_host.so"))
    names = declared("mox_host.h", "moxh_")
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), n


def test_oracle_exports_the_same_surface():
    lib = C.CDLL(os.path.join(ROOT, "oracle", "liboracle.so"))
    for n in declared("mox.h", "mox_"):
        if n in ("mox_trace_closest_device", "mox_create_multi", "mox_gather_export", "mox_gather_import", "mox_gather_push",
                 "mox_read_gathered_begin", "mox_read_gathered_end"):
            continue  # device pointers, several devices and peer memory make no sense for the CPU oracle
        assert hasattr(lib, "orc_" + n[4:]), n


def test_gpu_library_exports_header_symbols():
    path = os.path.join(ROOT, "minimaloptix_b200", "libmox.so")
    if not os.path.exists(path):
        pytest.skip("libmox.so not built (run __graft_entry__.build())")
    lib = C.CDLL(path)  # loads without a GPU: cudart is linked statically, no driver call at load
    names = declared("mox.h", "mox_") + declared("mox_debug.h", "mox_")
    assert "mox_launch" in names and "mox_build_accel" in names and len(names) >= 25
    for n in names:
        assert hasattr(lib, n), n
    lib.mox_abi_version.restype = C.c_int
    assert lib.mox_abi_version() == 3


def test_gpu_library_fails_loudly_without_a_gpu():
    """No CPU fallback: on a box without a usable GPU mox_create must fail with MOX_ERR_CUDA."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    path = os.path.join(ROOT, "minimaloptix_b200", "libmox.so")
    if not os.path.exists(path):
        pytest.skip("libmox.so not built")
    import minimaloptix_b200 as mox
    with pytest.raises(mox.MoxError, match="no CUDA device|CUDA"):
        mox.gpu().context(0)
