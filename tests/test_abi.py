"""The C-ABI library loads on a machine without a GPU, exports every symbol include/rumi_orb.h declares, and fails
loudly (no CPU fallback) when there is no CUDA device.  No compute calls here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "rumi_orb.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return set(re.findall(r"\b(rumi_[a-z0-9_]+)\s*\(", text))


def test_header_and_binding_agree():
    from rumi_slam_b200 import _lib
    assert header_symbols() == set(_lib.SIGNATURES), header_symbols() ^ set(_lib.SIGNATURES)


def test_library_exports_every_declared_symbol():
    from rumi_slam_b200 import _lib
    L = _lib.lib()                                     # builds with nvcc if missing; raises if a symbol is absent
    for name in header_symbols():
        assert hasattr(L, name), name


def test_no_gpu_means_error_not_fallback():
    from rumi_slam_b200 import _lib, RumiError, ORBextractor, ORBmatcher
    if _lib.lib().rumi_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(RumiError) as e:
        ORBextractor(1000, 1.2, 8, 20, 7)
    assert e.value.code == -4 and "no CPU fallback" in str(e.value)
    with pytest.raises(RumiError):
        ORBmatcher()
    from rumi_slam_b200 import SparsePyrLK, KFDSample
    with pytest.raises(RumiError) as e:
        SparsePyrLK()
    assert e.value.code == -4 and "no CPU fallback" in str(e.value)
    with pytest.raises(RumiError):
        KFDSample()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "rumi_slam_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cc")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "orb_oracle" not in src and "liborb_ref" not in src and "flow_oracle" not in src, f


def test_descriptor_distance_host_inline(oracle):
    from rumi_slam_b200 import ORBmatcher
    rng = np.random.default_rng(0)
    assert ORBmatcher.DescriptorDistance(np.zeros(32, np.uint8), np.full(32, 255, np.uint8)) == 256
    for _ in range(200):
        a, b = rng.integers(0, 256, (2, 32), dtype=np.uint8)
        assert ORBmatcher.DescriptorDistance(a, b) == oracle.descriptor_distance(a, b)
    assert (ORBmatcher.TH_HIGH, ORBmatcher.TH_LOW, ORBmatcher.HISTO_LENGTH) == (100, 50, 30)
