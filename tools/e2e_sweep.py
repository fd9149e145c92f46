"""tools/e2e_sweep.py — development aid: frames/s of the drop-in call at 4K under the environment's
transport settings (S3R_COPY_THREADS, S3R_HOST_BANDS, S3R_PACK24, S3R_NT_STORES, S3R_PIN_HOST)."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from swift3drenderer_b200 import assets, renderer as R, scene as S

W, H, F = 3840, 2160, 300
d = R.DropIn(assets.ensure_shipped_data_bin())
inp = S.input_script("flythrough", 600)
buf = np.zeros((2, H, W), np.uint32)
for f in range(20):
    d.update_and_render(W, H, inp[f], out=buf[f & 1])
d.reset_camera()
t0 = time.perf_counter()
for f in range(F):
    d.update_and_render(W, H, inp[f], out=buf[f & 1])
dt = time.perf_counter() - t0
print({k: v for k, v in os.environ.items() if k.startswith("S3R_")}, f"{F / dt:.0f} fps  {dt / F * 1e3:.3f} ms/frame")
