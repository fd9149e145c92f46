"""tools/record_flythrough.py [--out fly.y4m] [--width 1920 --height 1080] [--frames 600] [--format y4m|bgr0]

Headless stand-in for the reference shell's display path (main.swift:124-140): renders the recorded fly-through
device-resident and streams every frame into a file through the frame sink (`s3r_sink_*`, include/s3r_b200.h) —
YUV4MPEG2 4:2:0 converted on the GPU, or raw BGR0 — without the synchronous host round trip of `updateAndRender`.
Prints one JSON line with the frame rate including the file writes.  Needs a GPU."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from swift3drenderer_b200 import assets, renderer as R, scene as S  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="/dev/shm/s3r_flythrough.y4m")
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--frames", type=int, default=600)
    ap.add_argument("--format", default="y4m", choices=["y4m", "bgr0"])
    args = ap.parse_args()
    r = R.Renderer(0)
    r.load_scene_file(assets.ensure_shipped_data_bin())
    mats = R.camera_path(S.input_script("flythrough", args.frames))
    W, H = args.width, args.height
    ring = torch.zeros((2, H, W), dtype=torch.int32, device="cuda:0")
    sink = R.Sink(r, args.out, W, H, fps=(60, 1), fmt=1 if args.format == "y4m" else 0)
    t0 = time.perf_counter()
    for f in range(args.frames):
        r.render_device(mats[f], W, H, ring[f & 1].data_ptr())
        sink.submit(ring[f & 1].data_ptr())
    written = sink.close()
    dt = time.perf_counter() - t0
    r.close()
    print(json.dumps({"file": args.out, "format": args.format, "W": W, "H": H, "frames_written": written,
                      "bytes": os.path.getsize(args.out), "frames_per_s": written / dt, "seconds": dt}))


if __name__ == "__main__":
    main()
