"""tools/multi_e2e_sweep.py — frames/s of updateAndRender driving several GPUs from one process (S3R_DEVICES) for a few
S3R_MULTI_BANDS settings, on the demo scene (C2) and optionally the C3 field.  Development aid."""
import os, subprocess, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

CHILD = r'''
import os, sys, time
import numpy as np
sys.path.insert(0, %(root)r)
from swift3drenderer_b200 import assets, renderer as R, scene as S
import bench
W, H = 3840, 2160
which = %(which)r
if which == "c2":
    path, inp, n = assets.ensure_shipped_data_bin(), S.input_script("flythrough", 600), 600
else:
    path, inp, n = bench.c3_data_bin(1_000_000), bench.drift_inputs(8), 8
d = R.DropIn(path, devices=%(devices)r, env={"S3R_MULTI_BANDS": %(bands)r})
buf = np.zeros((2, H, W), np.uint32)
for f in range(min(n, 24)):
    d.update_and_render(W, H, inp[f], out=buf[f & 1])
reps = 1 if which == "c2" else 20
t0 = time.perf_counter()
for _ in range(reps):
    d.reset_camera()
    for f in range(n):
        d.update_and_render(W, H, inp[f], out=buf[f & 1])
dt = time.perf_counter() - t0
print(which, "devices", %(devices)r, "bands", %(bands)r, "fps %%.1f" %% (reps * n / dt), flush=True)
d.close()
'''

if __name__ == "__main__":
    devices = sys.argv[1] if len(sys.argv) > 1 else "0,1"
    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
    for which in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["c2"]):
        for bands in (sys.argv[3].split(",") if len(sys.argv) > 3 else ["1", "2", "3", "4", "6", "8"]):
            subprocess.run([sys.executable, "-c", CHILD % {"root": root, "which": which, "devices": devices, "bands": bands}])
