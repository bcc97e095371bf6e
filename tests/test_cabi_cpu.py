"""CPU tests of the drop-in boundary: libuvip_orb.so loads, exports every symbol include/uvip_orb.h declares,
and refuses loudly (never falls back to a CPU path) when no CUDA device exists."""
import ctypes as C
import os

import numpy as np
import pytest


def _has_gpu(pkg):
    return pkg.capi.lib().uvip_device_count() > 0


def test_library_exports_every_declared_symbol(pkg):
    L = pkg.capi.lib()
    syms = pkg.capi.declared_symbols()
    assert len(syms) >= 28
    missing = [s for s in syms if not hasattr(L, s)]
    assert not missing, missing
    assert L.uvip_abi_version() == 1


def test_keypoint_layout_is_cv_keypoint(pkg):
    assert pkg.capi.KP_DTYPE.itemsize == 28
    assert [pkg.capi.KP_DTYPE.fields[n][1] for n in ('x', 'y', 'size', 'angle', 'response', 'octave', 'class_id')] == \
        [0, 4, 8, 12, 16, 20, 24]


def test_host_only_entry_points(pkg, oracle):
    L = pkg.capi.lib()
    for c in (0.9981, 0.998, 0.9979, 0.5, 1.0):
        assert L.uvip_radius_by_viewing_cos(c) == oracle.lib().uo_radius_by_viewing_cos(c)
    assert L.uvip_radius_by_viewing_cos(0.999) == 2.5 and L.uvip_radius_by_viewing_cos(0.9) == 4.0


def test_no_device_means_error_not_fallback(pkg):
    if _has_gpu(pkg):
        pytest.skip('a CUDA device is present')
    with pytest.raises(pkg.capi.UvipError) as e:
        pkg.ORBextractor(1000, 1.2, 8, 1, 20)
    assert e.value.code == pkg.capi.ERR_NO_DEVICE
    with pytest.raises(pkg.capi.UvipError) as e:
        pkg.ORBmatcher(0.75, True)
    assert e.value.code == pkg.capi.ERR_NO_DEVICE
    assert b'no CPU fallback' in pkg.capi.lib().uvip_last_error()


def test_pattern_tables_identical(pkg):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    a = open(os.path.join(root, 'oracle', 'orb_pattern.inc')).read()
    b = open(os.path.join(root, 'u-vip-slam_b200', 'csrc', 'orb_pattern.inc')).read()
    assert a == b
    vals = [int(v) for v in a.split('*/')[1].replace('\n', '').split(',') if v.strip()]
    assert len(vals) == 1024 and vals[:4] == [8, -3, 9, 5] and vals[-4:] == [-1, -6, 0, -11]
    assert max(abs(v) for v in vals) <= 13


def test_product_does_not_import_oracle():
    """the product package must not route through the oracle (or any CPU fallback)"""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for dirpath, _, files in os.walk(os.path.join(root, 'u-vip-slam_b200')):
        for fn in files:
            if fn.endswith(('.py', '.cu', '.cuh', '.h', '.cpp')):
                txt = open(os.path.join(dirpath, fn)).read()
                assert 'uvip_oracle' not in txt and 'from oracle' not in txt and 'import oracle' not in txt, fn
