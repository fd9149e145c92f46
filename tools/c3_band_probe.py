"""tools/c3_band_probe.py — development aid: renders one 1/8 screen band of the C3 scene on a single GPU
(what each rank does in the 8-GPU configuration; C3_WORLD=1 C3_PHASE=0 renders the whole frame) so that ncu can list its kernels."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
from swift3drenderer_b200 import renderer as R, scene as S

import bench
n_solids = int(os.environ.get("C3_SOLIDS", "1000000"))
path = bench.c3_data_bin(n_solids)   # the bench's own scene file (generated once per box)
r = R.Renderer(0)
r.load_scene_file(path)
mats = R.camera_path(bench.drift_inputs(4))
for opt in os.environ.get("S3R_OPTS", "").split(","):   # e.g. S3R_OPTS=clusters=0,cluster_cull=0
    if "=" in opt:
        r.set_option(opt.split("=")[0], int(opt.split("=")[1]))
W, H = 3840, 2160
world, phase = int(os.environ.get("C3_WORLD", "8")), int(os.environ.get("C3_PHASE", "3"))
rows, _, _ = R.rows_layout(H, world, phase)
out = torch.zeros((max(rows, H // world + 1), W), dtype=torch.int32, device="cuda:0")
band = os.environ.get("C3_MODE", "rows") == "band"   # contiguous band [H * phase / world, H * (phase + 1) / world) instead of interleaved tile rows


def frame(f):
    if band:
        r.render_device(mats[f], W, H, out.data_ptr(), y0=H * phase // world, y1=H * (phase + 1) // world)
    else:
        r.render_device_rows(mats[f], W, H, world, phase, out.data_ptr())


for rep in range(int(os.environ.get("C3_WARM", "3"))):
    for f in range(4):
        frame(f)
    while r.finish():
        pass
r.set_option("timing", 1); r.timing()
for f in range(4):
    frame(f)
r.finish()
kt = r.kernel_timing()
print({k: round(v["ms"] * 1e3 / max(v["launches"], 1), 1) for k, v in kt.items()}, r.timing(), r.stats())
