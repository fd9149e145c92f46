"""oracle/refso.py — TEST INFRASTRUCTURE.  Drives ``oracle/_ref/render_ref.so`` (the reference's
unmodified render-cpp/render.cpp built by oracle/Makefile) through its only entry point,
``updateAndRender(const PixelData*, const Input*)`` (render-cpp/render.cpp:264-265).

The reference keeps its camera and scene in process-lifetime statics (render.cpp:51-113) and finds
``data.bin`` next to its own file via ``dladdr`` (render.cpp:161-176), so every camera path gets a
fresh private copy of the library beside the scene file.
"""
from __future__ import annotations

import ctypes
import os
import shutil
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "render_ref.so")


class PixelData(ctypes.Structure):  # render-cpp/render.hpp:7-13
    _fields_ = [
        ("buffer", ctypes.POINTER(ctypes.c_uint32)),
        ("width", ctypes.c_uint32),
        ("height", ctypes.c_uint32),
        ("bytesPerPixel", ctypes.c_uint32),
        ("bufferSize", ctypes.c_uint32),
    ]


class Input(ctypes.Structure):  # render-cpp/render.hpp:15-21
    _fields_ = [
        ("up", ctypes.c_float),
        ("down", ctypes.c_float),
        ("left", ctypes.c_float),
        ("right", ctypes.c_float),
        ("mouse", ctypes.c_float * 2),
    ]


assert ctypes.sizeof(PixelData) == 24 and ctypes.sizeof(Input) == 24


def available() -> bool:
    return os.path.exists(REF_SO)


def make_input(rec) -> Input:
    i = Input()
    i.up, i.down, i.left, i.right = float(rec["up"]), float(rec["down"]), float(rec["left"]), float(rec["right"])
    i.mouse[0], i.mouse[1] = float(rec["mouse"][0]), float(rec["mouse"][1])
    return i


class RefRenderer:
    """One private instance of the reference library bound to one ``data.bin``."""

    def __init__(self, data_bin_path: str, so_path: str = REF_SO):
        if not os.path.exists(so_path):
            raise FileNotFoundError(so_path + " (run `make -C oracle ref` in the build container)")
        self._dir = tempfile.mkdtemp(prefix="refso_")
        self._so = os.path.join(self._dir, "render.so")
        shutil.copy(so_path, self._so)
        dst = os.path.join(self._dir, "data.bin")
        try:
            os.symlink(os.path.abspath(data_bin_path), dst)
        except OSError:
            shutil.copy(data_bin_path, dst)
        self._lib = ctypes.CDLL(self._so)
        self._fn = self._lib.updateAndRender
        self._fn.argtypes = [ctypes.POINTER(PixelData), ctypes.POINTER(Input)]
        self._fn.restype = None

    def update_and_render(self, width: int, height: int, rec, out: np.ndarray | None = None) -> np.ndarray:
        if out is None:
            out = np.empty((height, width), np.uint32)
        assert out.dtype == np.uint32 and out.size == width * height and out.flags.c_contiguous
        pd = PixelData(out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)), width, height, 4, 4 * width * height)
        inp = make_input(rec)
        self._fn(ctypes.byref(pd), ctypes.byref(inp))
        return out

    def close(self) -> None:
        shutil.rmtree(self._dir, ignore_errors=True)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
