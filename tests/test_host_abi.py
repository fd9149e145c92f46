"""Host-side logic of the product that needs no GPU: the C-ABI library loads and exports every symbol
include/*.h declares, the host camera equals the oracle's, errors are loud, and the product never
touches oracle/."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from swift3drenderer_b200 import scene as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = set()
    for h in ("render.h", "s3r_b200.h"):
        text = open(os.path.join(ROOT, "include", h)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names |= set(re.findall(r"\b(s3r_[a-z_0-9]+|updateAndRender)\s*\(", text))
    return names


def test_library_exports_every_declared_symbol(renderer_lib):
    lib = ctypes.CDLL(renderer_lib.LIB_PATH)
    decl = declared_symbols()
    assert "updateAndRender" in decl and len(decl) >= 18
    for name in sorted(decl):
        assert hasattr(lib, name), f"{name} declared in include/ but not exported by render.so"
    assert decl == set(renderer_lib.EXPORTS)


def test_only_the_abi_is_exported(renderer_lib):
    out = subprocess.check_output(["nm", "-D", "--defined-only", renderer_lib.LIB_PATH], text=True)
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    assert exported == declared_symbols()


def test_boundary_struct_layouts(renderer_lib):
    R = renderer_lib
    assert ctypes.sizeof(R.PixelData) == 24 and ctypes.sizeof(R.Input) == 24  # render.hpp:7-21
    assert R.PixelData.width.offset == 8 and R.PixelData.bufferSize.offset == 20
    assert R.Input.mouse.offset == 16


def test_host_camera_equals_oracle_camera(renderer_lib, oracle_port):
    for script, n in (("flythrough", 600), ("spin", 200), ("c1_path", 300), ("still", 3)):
        inp = S.input_script(script, n)
        a, b = renderer_lib.camera_path(inp), oracle_port.camera_path(inp)
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), script


def test_camera_translates_with_old_axes_then_rotates(renderer_lib):
    cam = renderer_lib.Camera()
    rec = np.zeros((), S.INPUT_DTYPE)
    rec["up"] = 1
    rec["mouse"] = (50, 0)
    m = cam.update(rec)
    # translation used the initial axes: position = (0, 0, -0.1)  (render.cpp:136-139 before :140-150)
    assert np.allclose(cam.c.position[:], [0, 0, -0.1], atol=1e-7)
    assert m[8:11] @ np.array([0, 0, 1.0]) < 1.0  # z axis turned


def test_factor_matches_reference_constant(renderer_lib):
    lib = renderer_lib.load_library()
    assert float(lib.s3r_factor(36)).hex() == "0x1.bb2fba0000000p+5"  # read back from the reference build
    assert lib.s3r_factor(2160) == pytest.approx(2160 / (2 * np.tan(np.pi / 10)), rel=1e-6)


def test_create_fails_loudly_without_a_gpu(renderer_lib):
    from conftest import HAS_GPU
    if HAS_GPU:
        pytest.skip("a GPU is present")
    with pytest.raises(renderer_lib.RendererError, match="no CPU fallback"):
        renderer_lib.Renderer(0)


def test_missing_library_is_an_error(renderer_lib, tmp_path):
    with pytest.raises(renderer_lib.RendererError):
        renderer_lib.load_library(str(tmp_path / "nope.so"))


def test_product_never_references_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may touch oracle/: the package, the
    headers and the tools (profiling / recording aids) must not."""
    for top in ("swift3drenderer_b200", "tools", "include"):
        for d, _, files in os.walk(os.path.join(ROOT, top)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".sh", "Makefile")):
                    text = open(os.path.join(d, f)).read()
                    code = "\n".join(l for l in text.splitlines() if "import" in l or "#include" in l or "dlopen" in l or "CDLL" in l)
                    assert "oracle" not in code, f"{f} pulls in oracle/: {code}"


def test_host_band_edges_cover_every_tile_row_once(renderer_lib):
    """s3r_render_host pipelines raster launches with device-to-host copies band by band; its band edges (uniform, last band
    tapered) must start at 0, end at tiles_y and grow strictly — an empty band would be an invalid launch."""
    import ctypes
    lib = renderer_lib.load_library()
    out = (ctypes.c_uint32 * 65)()
    for tiles_y in list(range(1, 140)) + [2048]:
        for bands in (1, 2, 3, 4, 7, 8, 12, 16, 24, 62, 64, 1000):
            for taper in (0, 1):
                n = lib.s3r_debug_band_edges(tiles_y, bands, taper, out, 65)
                e = list(out[:n])
                assert n >= 2 and e[0] == 0 and e[-1] == tiles_y and all(b > a for a, b in zip(e, e[1:])), (tiles_y, bands, taper, e)
                if not taper:
                    assert n - 1 == min(bands, tiles_y, 62)
    n = lib.s3r_debug_band_edges(68, 12, 1, out, 65)          # 4K: 68 tile rows, 12 bands, last one split 3 + 1 + 2
    assert list(out[:n]) == [0, 5, 11, 17, 22, 28, 34, 39, 45, 51, 56, 62, 65, 66, 68]
