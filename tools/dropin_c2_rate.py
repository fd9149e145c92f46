"""tools/dropin_c2_rate.py <devices> — frames/s of updateAndRender on the demo scene at 4K through the in-process multi-GPU
drop-in, from a process that loads nothing but numpy and the library.  Development aid."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from swift3drenderer_b200 import assets, renderer as R, scene as S
W, H = 3840, 2160
path, inp = assets.ensure_shipped_data_bin(), S.input_script("flythrough", 600)
d = R.DropIn(path, devices=sys.argv[1])
buf = np.zeros((2, H, W), np.uint32)
for f in range(24):
    d.update_and_render(W, H, inp[f], out=buf[f & 1])
t0 = time.perf_counter()
for f in range(600):
    d.update_and_render(W, H, inp[f], out=buf[f & 1])
print("c2 devices", sys.argv[1], "fps %.1f" % (600 / (time.perf_counter() - t0)), "torch loaded:", "torch" in sys.modules, flush=True)
d.close()
