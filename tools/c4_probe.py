"""tools/c4_probe.py — development aid: renders a few frames of the C4 clipping-stress scene (whole frame, one GPU)
so that ncu can list / capture its kernels."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
from swift3drenderer_b200 import renderer as R, scene as S

n_solids = int(os.environ.get("C4_SOLIDS", "50000"))
path = f"/dev/shm/s3r_c4_{n_solids}.data.bin"
if not os.path.exists(path):
    S.write_data_bin(path, S.clip_stress_scene(n_solids))
r = R.Renderer(0)
r.load_scene_file(path)
if os.environ.get("FLAT_MAX"):
    r.set_option("flat_max", int(os.environ["FLAT_MAX"]))
mats = R.camera_path(S.input_script("strafe", 16))
W, H = 3840, 2160
out = torch.zeros((H, W), dtype=torch.int32, device="cuda:0")
for rep in range(3):
    for f in range(4):
        r.render_device(mats[f], W, H, out.data_ptr())
    while r.finish():
        pass
r.set_option("timing", 1); r.timing()
for f in range(4):
    r.render_device(mats[f], W, H, out.data_ptr())
r.finish()
print(r.timing(), r.stats())
