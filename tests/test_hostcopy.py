"""The host copy engine (swift3drenderer_b200/csrc/hostcopy.cpp): 24 -> 32 bit pixel expansion and the
threaded, slice-by-slice staging -> caller-buffer copy.  Pure host code, no GPU."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("hc") / "hostcopy_host.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-pthread",
                           os.path.join(ROOT, "tests", "native", "hostcopy_host.cpp"), "-o", out])
    L = ctypes.CDLL(out)
    L.t_unpack24.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]
    L.t_copier.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t,
                           ctypes.c_int, ctypes.c_int]
    return L


@pytest.mark.parametrize("n", [0, 1, 3, 15, 16, 17, 18, 19, 64, 1000, 3840 * 7 + 5])
def test_unpack24_matches_definition(lib, n):
    rs = np.random.RandomState(n)
    src = rs.randint(0, 256, 3 * n + 1).astype(np.uint8)  # +1: a guard byte that must not matter
    dst = np.full(n + 4, 0xDEADBEEF, np.uint32)
    lib.t_unpack24(dst.ctypes.data, src.ctypes.data, n)
    s = src[: 3 * n].reshape(-1, 3).astype(np.uint32)
    assert np.array_equal(dst[:n], s[:, 0] | (s[:, 1] << 8) | (s[:, 2] << 16))
    assert (dst[n:] == 0xDEADBEEF).all()  # never writes past the slice


@pytest.mark.parametrize("workers", [1, 3, 8])
@pytest.mark.parametrize("packed", [0, 1])
def test_threaded_copier_fills_every_pixel(lib, workers, packed):
    px = 640 * 360 + 13
    rs = np.random.RandomState(workers)
    want = rs.randint(0, 1 << 24, px).astype(np.uint32)
    if packed:
        src = np.stack([want & 255, (want >> 8) & 255, want >> 16], -1).astype(np.uint8).reshape(-1)
        src = np.concatenate([src, np.zeros(16, np.uint8)])
    else:
        src = want.copy().view(np.uint8)
    dst = np.zeros(px, np.uint32)
    lib.t_copier(workers, packed, src.ctypes.data, dst.ctypes.data, px, 7, 3)
    assert np.array_equal(dst, want)
