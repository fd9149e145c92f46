"""Parity at BASELINE.json's FULL sizes, through the C ABI, against the oracle (C port, pinned to the unmodified
reference) and against digests produced by the unmodified reference itself:

  C1  data.bin scene, 1280x720, frames of the 300-frame c1_path script                      vs the port
  C2  data.bin scene WITH THE REFERENCE'S PPM ATLASES, 3840x2160, recorded fly-through     vs row digests made by render_ref.so
  C3  1 M textured icosahedrons (20 M triangles), 3840x2160, pose 0 of the bench's path    vs the port; bands / rows = whole
  C4  clipping-stress field (50 k solids, ~36 % of triangles straddle the near plane), 4K    vs the port
  C5  4096 poses at 512x512: poses 0 / 700 / 4095, single and inside a batch               vs the port

Bit-exact (zero differing pixels) everywhere; north_star's tolerance (+-1 LSB on >= 99.9 %) is met a fortiori."""
import os
import sys
import zlib

import numpy as np
import pytest

from swift3drenderer_b200 import assets, scene as S

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def assert_same(a, b, what):
    if not np.array_equal(a, b):
        ys, xs = np.nonzero(a != b)
        raise AssertionError(f"{what}: {len(ys)} px differ, first (x={xs[0]}, y={ys[0]}) "
                             f"got {a[ys[0], xs[0]]:06x} want {b[ys[0], xs[0]]:06x}")


def test_c1_path_frames_at_1280x720(gpu_renderer, renderer_lib, oracle_port):
    path = assets.ensure_shipped_data_bin()
    gpu_renderer.load_scene_file(path)
    osc = oracle_port.OracleScene(path=path)
    mats = renderer_lib.camera_path(S.input_script("c1_path", 300))
    for f in (0, 59, 60, 119, 120, 200, 299):
        assert_same(gpu_renderer.render(mats[f], 1280, 720)[0], osc.render(mats[f], 1280, 720)["pixels"], f"C1 frame {f}")


def test_c2_with_reference_ppm_atlases_matches_reference_digests(gpu_renderer, renderer_lib):
    """The shipped scene with the reference's own atlases at 4K against per-row CRCs of frames rendered by the unmodified
    render.cpp (tests/golden/make_ppm_digests.py)."""
    fx = np.load(os.path.join(ROOT, "tests", "golden", "shipped_ppm_4k.npz"))
    path = assets.ensure_shipped_data_bin()
    sc = S.read_data_bin(path)
    if zlib.crc32(np.ascontiguousarray(sc.textures, "<u4").tobytes()) != int(fx["textures_crc32"]):
        pytest.skip("scenes/_gen data.bin was generated without the reference's ppm atlases")
    gpu_renderer.load_scene_file(path)
    mats = renderer_lib.camera_path(S.input_script("flythrough", 600))
    for k, f in enumerate(fx["frames"]):
        got = gpu_renderer.render(mats[f], 3840, 2160)[0]
        rows = np.asarray([zlib.crc32(got[y].tobytes()) for y in range(2160)], np.uint32)
        bad = np.nonzero(rows != fx["row_crc32"][k])[0]
        assert len(bad) == 0, f"C2/ppm frame {f}: {len(bad)} rows differ from the reference, first y={bad[0]}"
        assert zlib.crc32(got.tobytes()) == int(fx["crc32"][k]) and int(got.sum(dtype=np.uint64)) == int(fx["sum"][k])


def test_c2_drop_in_with_ppm_atlases_matches_reference_digests(renderer_lib):
    """The same frames through updateAndRender (host buffers, stale-factor rule, internal camera)."""
    fx = np.load(os.path.join(ROOT, "tests", "golden", "shipped_ppm_4k.npz"))
    path = assets.ensure_shipped_data_bin()
    sc = S.read_data_bin(path)
    if zlib.crc32(np.ascontiguousarray(sc.textures, "<u4").tobytes()) != int(fx["textures_crc32"]):
        pytest.skip("scenes/_gen data.bin was generated without the reference's ppm atlases")
    d = renderer_lib.DropIn(path)
    inp = S.input_script("flythrough", 600)
    out = np.zeros((2, 2160, 3840), np.uint32)
    want = {int(f): k for k, f in enumerate(fx["frames"])}
    for f in range(600):
        d.update_and_render(3840, 2160, inp[f], out=out[f & 1])
        if f in want:
            assert zlib.crc32(out[f & 1].tobytes()) == int(fx["crc32"][want[f]]), f"drop-in frame {f} differs from the reference"
    d.close()


@pytest.fixture(scope="module")
def c3_path():
    import bench
    return bench.c3_data_bin(1_000_000)


@pytest.fixture(scope="module")
def c3_oracle_pose0(c3_path, oracle_port):
    """Pose 0 of the bench's drift path by the port (12 s of one host core): rendered once for the module."""
    import bench
    mats = oracle_port.camera_path(bench.drift_inputs(1))
    return oracle_port.OracleScene(path=c3_path).render(mats[0], 3840, 2160)


def test_c3_full_scene_pose0_at_4k(c3_path, c3_oracle_pose0, renderer_lib):
    """20 M triangles at 3840x2160: whole frame vs the port; the band / interleaved-row partitions reassemble to it."""
    import bench
    r = renderer_lib.Renderer(0)
    try:
        r.load_scene_file(c3_path)
        mats = renderer_lib.camera_path(bench.drift_inputs(8))
        W, H = 3840, 2160
        got = r.render(mats[0], W, H)[0]
        o = c3_oracle_pose0
        assert_same(got, o["pixels"], "C3 pose 0")
        st = r.stats()
        assert st["near_rejected"] == o["stats"]["near_rejected"] and st["clipped"] == o["stats"]["clipped"]
        assert st["setups"] == o["stats"]["rasterized"]
        assert st["culled"] == o["stats"]["offscreen"] + o["stats"]["small_or_backfacing"]
        # screen partitions of the multi-GPU configuration, on one GPU: 8 contiguous bands, 8 interleaved phases
        parts = [r.render(mats[0], W, H, y0=H * k // 8, y1=H * (k + 1) // 8)[0] for k in range(8)]
        assert_same(np.concatenate(parts, 0), got, "C3 pose 0, 8 bands")
        import torch
        th = renderer_lib.tile_height()
        frame = np.zeros((H, W), np.uint32)
        for k in range(8):
            nrows, frame_rows, buf_rows = renderer_lib.rows_layout(H, 8, k, th)
            buf = torch.zeros((nrows, W), dtype=torch.int32, device="cuda:0")
            for attempt in range(4):
                r.render_device_rows(mats[0], W, H, 8, k, buf.data_ptr())
                if not r.finish():
                    break
            frame[frame_rows] = buf.cpu().numpy().view(np.uint32)[buf_rows]
        assert_same(frame, got, "C3 pose 0, 8 interleaved phases")
        # a later pose of the path against its own partition (no oracle: 4 s per frame on the CPU is spent once)
        whole7 = r.render(mats[7], W, H)[0]
        parts = [r.render(mats[7], W, H, y0=H * k // 2, y1=H * (k + 1) // 2)[0] for k in range(2)]
        assert_same(np.concatenate(parts, 0), whole7, "C3 pose 7, 2 bands")
    finally:
        r.close()


def test_c3_drop_in_poses(c3_path, c3_oracle_pose0, renderer_lib):
    """updateAndRender on C3's data.bin (the bench's e2e call): pose 0 against the port, pose 1 against the device path."""
    import bench
    d = renderer_lib.DropIn(c3_path)
    inp = bench.drift_inputs(2)
    out = np.zeros((2, 2160, 3840), np.uint32)
    d.update_and_render(3840, 2160, inp[0], out=out[0])
    d.update_and_render(3840, 2160, inp[1], out=out[1])
    d.close()
    assert_same(out[0], c3_oracle_pose0["pixels"], "C3 drop-in pose 0")
    r = renderer_lib.Renderer(0)
    try:
        r.load_scene_file(c3_path)
        assert_same(out[1], r.render(renderer_lib.camera_path(inp)[1], 3840, 2160)[0], "C3 drop-in pose 1")
    finally:
        r.close()


def test_c4_full_clipping_stress_at_4k(renderer_lib, oracle_port):
    sc = S.clip_stress_scene(50_000)
    r = renderer_lib.Renderer(0)
    try:
        r.load_scene(sc)
        osc = oracle_port.OracleScene(sc)
        mats = renderer_lib.camera_path(S.input_script("strafe", 16))
        for f in (0, 15):
            o = osc.render(mats[f], 3840, 2160)
            assert_same(r.render(mats[f], 3840, 2160)[0], o["pixels"], f"C4 frame {f}")
            st = r.stats()
            assert st["clipped"] == o["stats"]["clipped"] and st["spawned"] == o["stats"]["spawned"]
            assert st["clipped"] > 0.3 * sc.n_triangles   # the configuration's point: > 30 % straddle the near plane
    finally:
        r.close()


def test_c5_poses_single_and_batched(gpu_renderer, renderer_lib, oracle_port):
    path = assets.ensure_shipped_data_bin()
    gpu_renderer.load_scene_file(path)
    osc = oracle_port.OracleScene(path=path)
    mats = renderer_lib.camera_path(S.input_script("spin", 4096))
    for f in (0, 700, 4095):
        want = osc.render(mats[f], 512, 512)["pixels"]
        assert_same(gpu_renderer.render(mats[f], 512, 512)[0], want, f"C5 pose {f}")
        lo = max(0, min(f - 3, 4096 - 64))
        batch = gpu_renderer.render(mats[lo:lo + 64], 512, 512)
        assert_same(batch[f - lo], want, f"C5 pose {f} inside a 64-pose batch")
