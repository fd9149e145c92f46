"""tests/golden/make_golden.py — generates the committed golden frames by running the UNMODIFIED
reference (oracle/_ref/render_ref.so = /root/reference/render-cpp/render.cpp + oracle shim) on the
seeded cases in cases.py.  Run in the build container only:  python tests/golden/make_golden.py
Output: tests/golden/<case>.npz  {frames: (k,) int, pixels: (k, H, W) uint32}"""
import os, sys, tempfile
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, HERE)
import numpy as np
from swift3drenderer_b200 import scene as S
from oracle import refso
from cases import CASES

assert refso.available(), "oracle/_ref/render_ref.so missing: make -C oracle ref"
for name, (factory, script, n, keep, (W, H)) in CASES.items():
    sc = factory()
    assert not S.validate(sc), S.validate(sc)
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "data.bin")
        S.write_data_bin(path, sc)
        ref = refso.RefRenderer(path)
        inp = S.input_script(script, n)
        out = []
        for f in range(max(keep) + 1):
            img = ref.update_and_render(W, H, inp[f])
            if f in keep:
                out.append(img.copy())
        ref.close()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), frames=np.asarray(keep), pixels=np.stack(out))
    print(name, len(keep), "frames", os.path.getsize(os.path.join(HERE, name + ".npz")), "bytes")
