"""tests/parity_report.py [--out profiles/parity_r01] — the parity measurement of SURVEY.md 8(d), with diff images.

For a fixed subset of frames of every configuration (C2: every 50th frame of the 600-frame fly-through at
3840x2160; C1; reduced C3/C4 fields; a C5 pose batch) the frame rendered by render.so on cuda:0 is compared with
the CPU checker (oracle/, the C restatement pinned bit-identical to the unmodified render.cpp):

  coverage_identical_pct   pixels whose background / not-background state agrees
  within_1_lsb_pct         pixels whose three channels differ by at most 1
  over_1_lsb, max_diff     count of the others, largest channel difference
  edge_pixels / off-edge   the same counts split by an edge mask = pixels within one pixel of a projected
                           triangle edge (|w_i| <= |dx_i| + |dy_i| for a weight of a surviving triangle whose other
                           two weights are within one pixel of non-negative)

north_star's bar is: identical coverage/depth decisions off-edge and >= 99.9 % of pixels within +-1 LSB.
Writes <out>/report.json and PNGs (zlib only, no imaging library): <cfg>_fNNNN_diff.png — full resolution, 8-bit,
64 x the largest channel difference (black = identical) — and <cfg>_fNNNN_gpu.png, the GPU frame box-filtered down
to at most 640 pixels wide.  Needs a GPU; the checker is used as checker only.
(Lives under tests/: the oracle is test infrastructure.)"""
from __future__ import annotations

import argparse
import json
import os
import struct
import sys
import time
import zlib

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from swift3drenderer_b200 import renderer as R, scene as S  # noqa: E402
from oracle import port  # noqa: E402  (checker)

BACKGROUND = 0x001E1E1E  # render.cpp:96


def write_png(path: str, img: np.ndarray) -> None:
    """img: (H, W) uint8 grey or (H, W, 3) uint8 RGB."""
    h, w = img.shape[:2]
    colour_type = 0 if img.ndim == 2 else 2
    raw = np.concatenate([np.zeros((h, 1), np.uint8), img.reshape(h, -1)], axis=1).tobytes()

    def chunk(tag: bytes, data: bytes) -> bytes:
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)

    with open(path, "wb") as fh:
        fh.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, colour_type, 0, 0, 0))
                 + chunk(b"IDAT", zlib.compress(raw, 9)) + chunk(b"IEND", b""))


def channels(px: np.ndarray) -> np.ndarray:
    return np.stack([(px >> 16) & 255, (px >> 8) & 255, px & 255], axis=-1).astype(np.int16)


def thumbnail(px: np.ndarray, max_w: int = 640) -> np.ndarray:
    h, w = px.shape
    k = max(1, -(-w // max_w))
    hh, ww = (h // k) * k, (w // k) * k
    c = channels(px[:hh, :ww]).astype(np.float32).reshape(hh // k, k, ww // k, k, 3).mean(axis=(1, 3))
    return c.astype(np.uint8)


def edge_mask(setups: np.ndarray, W: int, H: int, limit: int = 4096):
    """Pixels within one pixel of an edge of a surviving triangle (from the checker's setup records, evaluated in
    float64 from wstart/dx/dy — a mask, not a parity-relevant quantity).  None when there are too many triangles."""
    if len(setups) > limit:
        return None
    mask = np.zeros((H, W), bool)
    for s in setups:
        x0, x1, y0, y1 = int(s["xmin"]), int(s["xmax"]), int(s["ymin"]), int(s["ymax"])
        if x1 < x0 or y1 < y0:
            continue
        xs = np.arange(x0, x1 + 1, dtype=np.float64) - x0
        ys = np.arange(y0, y1 + 1, dtype=np.float64)[:, None] - y0
        w = [float(s["wstart"][i]) + xs * float(s["dx"][i]) + ys * float(s["dy"][i]) for i in range(3)]
        band = [abs(float(s["dx"][i])) + abs(float(s["dy"][i])) for i in range(3)]
        near_inside = [(w[i] >= -band[i]) for i in range(3)]
        m = np.zeros(w[0].shape, bool)
        for i in range(3):
            j, k = (i + 1) % 3, (i + 2) % 3
            m |= (np.abs(w[i]) <= band[i]) & near_inside[j] & near_inside[k]
        mask[y0:y1 + 1, x0:x1 + 1] |= m
    return mask


def compare(got: np.ndarray, want: np.ndarray, mask):
    d = np.abs(channels(got) - channels(want)).max(axis=-1)
    cov = (got != BACKGROUND) == (want != BACKGROUND)
    n = got.size
    rec = {
        "pixels": int(n),
        "differing_pixels": int((got != want).sum()),
        "coverage_identical_pct": 100.0 * float(cov.sum()) / n,
        "within_1_lsb_pct": 100.0 * float((d <= 1).sum()) / n,
        "over_1_lsb": int((d > 1).sum()),
        "max_diff": int(d.max()),
    }
    if mask is not None:
        rec["edge_pixels"] = int(mask.sum())
        rec["coverage_mismatch_off_edge"] = int((~cov & ~mask).sum())
        rec["coverage_mismatch_on_edge"] = int((~cov & mask).sum())
        rec["over_1_lsb_off_edge"] = int(((d > 1) & ~mask).sum())
        rec["over_1_lsb_on_edge"] = int(((d > 1) & mask).sum())
    return rec, d


def drift(n: int) -> np.ndarray:
    """tests/run_configs.py's C3 path: creep forward, pan a little."""
    inp = np.zeros(n, S.INPUT_DTYPE)
    for f in range(n):
        inp[f]["up"] = 1.0
        inp[f]["mouse"] = (0.5 * f, 0.2 * f)
    return inp


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "parity_r01"))
    ap.add_argument("--configs", default="c1,c2,c3s,c4s,c5")
    ap.add_argument("--thumb-every", type=int, default=2, help="write a GPU thumbnail for every n-th compared frame")
    args = ap.parse_args()
    os.makedirs(args.out, exist_ok=True)
    want_cfg = set(args.configs.split(","))

    shipped = S.shipped_scene(1)
    cases = []
    if "c1" in want_cfg:
        cases.append(("c1", "C1 data.bin scene, 1280x720, 300-frame path", shipped, R.camera_path(S.input_script("c1_path", 300)),
                      1280, 720, [0, 59, 119, 200, 299]))
    if "c2" in want_cfg:
        cases.append(("c2", "C2 data.bin scene, 3840x2160, 600-frame fly-through", shipped,
                      R.camera_path(S.input_script("flythrough", 600)), 3840, 2160, list(range(0, 600, 50)) + [599]))
    if "c3s" in want_cfg:
        cases.append(("c3s", "C3 generator at 1/50 scale (20 000 textured icosahedrons, 400 000 triangles), 3840x2160",
                      S.c3_scene(20_000), R.camera_path(drift(8)), 3840, 2160, [0, 7]))
    if "c4s" in want_cfg:
        cases.append(("c4s", "C4 generator at 1/10 scale (5 000 icosahedrons around the eye, near-plane stress), 3840x2160",
                      S.clip_stress_scene(5_000), R.camera_path(S.input_script("strafe", 16)), 3840, 2160, [0, 15]))
    if "c5" in want_cfg:
        cases.append(("c5", "C5 data.bin scene, 512x512, poses of the 4096-input path", shipped,
                      R.camera_path(S.input_script("spin", 4096)), 512, 512, [0, 700, 2000, 3000, 4095]))

    r = R.Renderer(0)
    report = {"checker": "oracle/render_oracle.c (pinned bit-identical to the unmodified render-cpp/render.cpp, see "
                         "tests/test_oracle_golden.py)", "bar": "coverage identical off-edge; >= 99.9 % within +-1 LSB",
              "configs": []}
    ok = True
    for key, title, sc, mats, W, H, frames in cases:
        r.load_scene(sc)
        orc = port.OracleScene(sc)
        entry = {"config": title, "W": W, "H": H, "triangles": int(len(sc.vertex_indices) // 3), "frames": []}
        for n_done, f in enumerate(frames):
            t0 = time.time()
            got = r.render(mats[f], W, H)[0]
            t1 = time.time()
            ref = orc.render(mats[f], W, H, want_setups=True)
            mask = edge_mask(ref["setups"], W, H)
            rec, d = compare(got, ref["pixels"], mask)
            rec.update({"frame": int(f), "surviving_triangles": int(len(ref["setups"])),
                        "gpu_call_s": round(t1 - t0, 4), "checker_s": round(time.time() - t1, 2)})
            name = f"{key}_f{f:04d}"
            write_png(os.path.join(args.out, name + "_diff.png"), np.minimum(d.astype(np.int32) * 64, 255).astype(np.uint8))
            rec["diff_png"] = name + "_diff.png"
            if n_done % args.thumb_every == 0:
                write_png(os.path.join(args.out, name + "_gpu.png"), thumbnail(got))
                rec["gpu_png"] = name + "_gpu.png"
            entry["frames"].append(rec)
            good = rec["within_1_lsb_pct"] >= 99.9 and rec.get("coverage_mismatch_off_edge", 100.0 - rec["coverage_identical_pct"]) == 0
            ok = ok and good
        entry["total_differing_pixels"] = sum(x["differing_pixels"] for x in entry["frames"])
        entry["worst_within_1_lsb_pct"] = min(x["within_1_lsb_pct"] for x in entry["frames"])
        report["configs"].append(entry)
        print(f"{key}: {len(frames)} frames, {entry['total_differing_pixels']} differing pixels, "
              f"worst within-1-LSB {entry['worst_within_1_lsb_pct']:.4f} %", flush=True)
    report["pass"] = bool(ok)
    with open(os.path.join(args.out, "report.json"), "w") as fh:
        json.dump(report, fh, indent=1)
    r.close()
    print("parity", "PASS" if ok else "FAIL")
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
