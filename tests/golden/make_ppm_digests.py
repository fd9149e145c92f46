"""tests/golden/make_ppm_digests.py — digests of 4K frames of the shipped scene WITH THE REFERENCE'S OWN PPM ATLASES,
rendered by the UNMODIFIED reference (oracle/_ref/render_ref.so) over the recorded fly-through.  A 4K frame is 33 MB,
so the fixture keeps, per frame, the CRC-32 of every pixel row plus the frame's CRC-32 and pixel sum — enough to tell
which rows differ.  Build container only (needs /root/reference):  python tests/golden/make_ppm_digests.py
Output: tests/golden/shipped_ppm_4k.npz {frames, crc32, sum, row_crc32, textures_crc32}"""
import os, sys, zlib
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
import numpy as np
from swift3drenderer_b200 import assets, scene as S
from oracle import refso

W, H, N = 3840, 2160, 600
KEEP = [0, 100, 215, 330, 470, 599]


def row_crcs(img):
    return np.asarray([zlib.crc32(img[y].tobytes()) for y in range(img.shape[0])], np.uint32)


if __name__ == "__main__":
    assert refso.available() and os.path.isdir(S.REFERENCE_PPM_DIR)
    path = assets.ensure_shipped_data_bin()
    sc = S.read_data_bin(path)
    assert np.array_equal(sc.textures.reshape(-1), S.load_ppm_atlases(S.REFERENCE_PPM_DIR).reshape(-1)), "data.bin does not hold the PPM atlases"
    ref = refso.RefRenderer(path)
    inp = S.input_script("flythrough", N)
    out = np.empty((H, W), np.uint32)
    small = np.empty((9, 16), np.uint32)
    crc, tot, rows = [], [], []
    for f in range(max(KEEP) + 1):
        if f in KEEP:
            ref.update_and_render(W, H, inp[f], out)
            crc.append(zlib.crc32(out.tobytes())); tot.append(int(out.sum(dtype=np.uint64))); rows.append(row_crcs(out))
        else:
            ref.update_and_render(16, 9, inp[f], small)
    ref.close()
    np.savez_compressed(os.path.join(HERE, "shipped_ppm_4k.npz"), frames=np.asarray(KEEP), crc32=np.asarray(crc, np.uint32),
                        sum=np.asarray(tot, np.uint64), row_crc32=np.stack(rows),
                        textures_crc32=np.uint32(zlib.crc32(np.ascontiguousarray(sc.textures, "<u4").tobytes())))
    print("wrote", len(KEEP), "frame digests;", os.path.getsize(os.path.join(HERE, "shipped_ppm_4k.npz")), "bytes")
