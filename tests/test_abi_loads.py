"""CPU: the C-ABI shared library builds, loads and exports every symbol that
include/rbpf_b200.h declares; without a GPU it refuses to create a handle."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from thesis_b200 import build, _lib

    build.build()
    return _lib.load()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "rbpf_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rbpf_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    from thesis_b200 import _lib

    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "missing export " + n
        assert n in _lib.SIGNATURES, "no ctypes signature for " + n
    assert sorted(_lib.SIGNATURES) == names


def test_matcher_lattice_constants(lib):
    import oracle as O

    assert lib.rbpf_rot_step() == O.rot_step()
    assert lib.rbpf_rot_count() == O.rot_count() == 115


def test_no_cpu_fallback(lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from thesis_b200.particles import ParticleSet, RbpfError

    with pytest.raises(RbpfError):
        ParticleSet(4, 180)


def test_config_struct_matches_header(lib):
    from thesis_b200 import _lib

    assert C.sizeof(_lib.RbpfConfig) == 56
    assert C.sizeof(_lib.RbpfStats) == 128
