"""Golden-vector case table shared by make_golden.py (generator) and the tests (consumers).
Every case is reproducible from seeds alone (procedural textures), so it runs where
/root/reference does not exist."""
from swift3drenderer_b200 import scene as S

CASES = {
    # name: (scene factory, input script, script length, frames to keep, (W, H))
    "shipped_160x90": (lambda: S.shipped_scene(1), "flythrough", 600,
                       [0, 30, 75, 110, 150, 185, 230, 280, 330, 380, 450, 520, 599], (160, 90)),
    "shipped_101x67": (lambda: S.shipped_scene(1), "flythrough", 600, [0, 110, 230, 450], (101, 67)),
    "ico_tex_160x90": (lambda: S.icosahedron_field(300, seed=3, extent=30), "spin", 21, [0, 10, 20], (160, 90)),
    "ico_col_160x90": (lambda: S.icosahedron_field(300, seed=4, extent=30, textured=False), "spin", 21, [0, 10, 20], (160, 90)),
    "clip_160x90": (lambda: S.clip_stress_scene(500), "spin", 16, [0, 7, 15], (160, 90)),
    # the reference's quirks in one scene (hazard list of SURVEY.md 8): still camera, then a slow turn with small steps
    "hazards_320x180": (lambda: S.hazard_scene(), "spin", 12, [0, 1, 5, 11], (320, 180)),
    "hazards_1283x721": (lambda: S.hazard_scene(), "spin", 12, [0, 3], (1283, 721)),
    "regfloor_160x90": (lambda: S.shipped_scene(2, regular_floor=True), "flythrough", 600, [40, 100, 150], (160, 90)),
}
