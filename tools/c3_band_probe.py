"""tools/c3_band_probe.py — development aid: renders one 1/8 screen band of the C3 scene on a single GPU
(what each rank does in the 8-GPU configuration; C3_WORLD=1 C3_PHASE=0 renders the whole frame) so that ncu can list its kernels."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
from swift3drenderer_b200 import renderer as R, scene as S

n_solids = int(os.environ.get("C3_SOLIDS", "1000000"))
path = f"/dev/shm/s3r_c3_{n_solids}.data.bin"
if not os.path.exists(path):
    S.write_data_bin(path, S.c3_scene(n_solids))
r = R.Renderer(0)
r.load_scene_file(path)
inp = np.zeros(4, S.INPUT_DTYPE)
mats = R.camera_path(inp)
W, H = 3840, 2160
world, phase = int(os.environ.get("C3_WORLD", "8")), int(os.environ.get("C3_PHASE", "3"))
rows, _, _ = R.rows_layout(H, world, phase)
out = torch.zeros((rows, W), dtype=torch.int32, device="cuda:0")
for rep in range(3):
    for f in range(4):
        r.render_device_rows(mats[f], W, H, world, phase, out.data_ptr())
    while r.finish():
        pass
r.set_option("timing", 1); r.timing()
for f in range(4):
    r.render_device_rows(mats[f], W, H, world, phase, out.data_ptr())
r.finish()
print(r.timing(), r.stats())
