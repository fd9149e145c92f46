"""tests/gpu_check.py — quick CUDA-vs-oracle parity sweep (development aid; the tests in tests/ are
the real gate).  Uses oracle/ only as the checker."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from swift3drenderer_b200 import scene as S, renderer as R
from oracle import port


def compare(name, sc, script, frames, sizes, pick):
    osc = port.OracleScene(sc)
    inp = S.input_script(script, frames)
    mats = R.camera_path(inp)
    r = R.Renderer(0)
    r.load_scene(sc)
    for (W, H) in sizes:
        bad = 0
        worst = 0
        t_gpu = 0.0
        for f in pick:
            t0 = time.time()
            g = r.render(mats[f], W, H)[0]
            t_gpu += time.time() - t0
            o = osc.render(mats[f], W, H)
            d = int((g != o["pixels"]).sum())
            if d:
                bad += 1
                worst = max(worst, d)
                if bad <= 3:
                    ys, xs = np.nonzero(g != o["pixels"])
                    print(f"  {name} {W}x{H} frame {f}: {d} px differ, first at (x={xs[0]}, y={ys[0]}) "
                          f"gpu={g[ys[0], xs[0]]:06x} cpu={o['pixels'][ys[0], xs[0]]:06x}; stats gpu={r.stats()} cpu={o['stats']}")
        print(f"{name} {W}x{H}: {len(pick)} frames, {bad} differ (worst {worst} px), gpu {t_gpu / len(pick) * 1e3:.2f} ms/frame (host API)")
    r.close()


if __name__ == "__main__":
    tx, src = S.default_textures()
    print("textures:", src)
    compare("shipped", S.shipped_scene(1, tx), "flythrough", 600, [(640, 360), (1280, 720), (333, 187)], list(range(0, 600, 12)))
    compare("shipped4k", S.shipped_scene(1, tx), "flythrough", 600, [(3840, 2160)], [0, 60, 120, 260])
    compare("ico_tex", S.icosahedron_field(2000, seed=3, extent=60), "spin", 30, [(640, 360)], list(range(0, 30, 3)))
    compare("ico_col", S.icosahedron_field(2000, seed=4, extent=60, textured=False), "spin", 30, [(640, 360)], list(range(0, 30, 3)))
    compare("clip", S.clip_stress_scene(3000), "spin", 30, [(640, 360)], list(range(0, 30, 3)))
    compare("regfloor", S.shipped_scene(2, tx, regular_floor=True), "flythrough", 120, [(480, 270)], list(range(0, 120, 6)))
