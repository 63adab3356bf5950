"""The C-ABI library loads on a GPU-less host and exports exactly what include/emcid_b200.h declares."""
import os
import re

import pytest
import torch

from emcid_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "emcid_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(emcid_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    names = declared_functions()
    assert len(names) >= 15
    lib = _lib.lib()
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/emcid_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == names, "ctypes bindings out of sync with the header"


def test_version_and_sizes_without_gpu():
    lib = _lib.lib()
    assert lib.emcid_version() >= 100
    assert lib.emcid_gemm3x_workspace_bytes(128, 256, 32) > 0
    ws = lib.emcid_mom2_workspace_bytes(3072, 768, 0)
    assert ws > 2 * 3072 * 1536 * 4
    assert lib.emcid_solve_workspace_bytes(1, 3072, 768, 1000) > 3072 * 3072 * 8


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_compute_entry_points_fail_loudly_without_gpu():
    lib = _lib.lib()
    rc = lib.emcid_device_check(0)
    assert rc < 0 and _lib.last_error()
    with pytest.raises(_lib.EmcidError):
        _lib.check(rc)
    from emcid_b200.mom2 import Mom2Accumulator
    with pytest.raises(RuntimeError):
        Mom2Accumulator("cpu", 256, 64)
