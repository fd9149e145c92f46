"""Scene assets for the renderer: the ``data.bin`` container, deterministic scene builders,
the rip-map atlas builder and scripted ``Input`` sequences.

The on-disk format is the reference's only input contract.  Writer spec:
``data-generator/main.swift:381-416``; reader: ``render-cpp/render.cpp:177-209``.

Little-endian, five sections, each prefixed by ``[count:u64, 0:u64]``:

  S1  V  x float4 (x, y, z, 1)
  S2  I  x u64 vertex indices            (+8 zero bytes when I is odd)
  S3  A  x 48-byte attribute records:    float4 normal (w = 0) @0 | 16-byte payload @16 |
                                          u32 kind @32 (0 colour, 1 texture) | 12 zero bytes
         colour payload : float r, g, b (0..255), 4 pad bytes
         texture payload: u32 index @16, 4 zero bytes, float u, v @24
  S4  I  x u64 attribute indices         (+8 zero bytes when I is odd)
  S5  nTex << 18 (count of u32 texels), then nTex x 512 x 512 u32 ``0x00RRGGBB`` rip-map atlases

This module is a from-scratch tool (numpy); the reference's generator is Swift + AppKit with
an unseeded RNG and cannot run here.  Geometry tables follow
``data-generator/main.swift:74-106`` (triangle), ``:108-188`` (regular floor), ``:190-216``
(simple floor), ``:218-258`` (tetrahedron), ``:260-373`` (icosahedron).
"""
from __future__ import annotations

import dataclasses
import os
from typing import Iterable, List, Optional, Sequence, Tuple

import numpy as np

ATTR_DTYPE = np.dtype(
    [("normal", "<f4", (4,)), ("payload", "<u4", (4,)), ("kind", "<u4"), ("pad", "<u4", (3,))]
)
assert ATTR_DTYPE.itemsize == 48

INPUT_DTYPE = np.dtype(
    [("up", "<f4"), ("down", "<f4"), ("left", "<f4"), ("right", "<f4"), ("mouse", "<f4", (2,))]
)
assert INPUT_DTYPE.itemsize == 24  # render-cpp/render.hpp:15-21

KIND_COLOR = 0
KIND_TEXTURE = 1
ATLAS = 512
TEXELS_PER_ATLAS = ATLAS * ATLAS  # 1 << 18, render-cpp/render.cpp:347

# NSColor.orange/.red/.blue through CIColor x 255 depends on macOS colour management and is
# unpinnable here; these are the nominal sRGB values.
ORANGE = (255.0, 127.5, 0.0)
RED = (255.0, 0.0, 0.0)
BLUE = (0.0, 0.0, 255.0)


@dataclasses.dataclass
class Scene:
    vertices: np.ndarray  # (V, 4) f32, w == 1
    vertex_indices: np.ndarray  # (I,) u64
    attributes: np.ndarray  # (A,) ATTR_DTYPE
    attribute_indices: np.ndarray  # (I,) u64
    textures: np.ndarray  # (nTex, 512, 512) u32

    @property
    def n_triangles(self) -> int:
        return int(self.vertex_indices.shape[0] // 3)

    def counts(self) -> dict:
        return {
            "V": int(self.vertices.shape[0]),
            "I": int(self.vertex_indices.shape[0]),
            "A": int(self.attributes.shape[0]),
            "T": self.n_triangles,
            "textures": int(self.textures.shape[0]),
        }


# ----------------------------------------------------------------------------------------
# data.bin I/O
# ----------------------------------------------------------------------------------------
def _header(count: int) -> bytes:
    return np.array([count, 0], dtype="<u8").tobytes()


def write_data_bin(path: str, scene: Scene) -> int:
    """Serialise ``scene`` in the reference's layout; returns the file size in bytes."""
    v = np.ascontiguousarray(scene.vertices, dtype="<f4")
    vi = np.ascontiguousarray(scene.vertex_indices, dtype="<u8")
    at = np.ascontiguousarray(scene.attributes, dtype=ATTR_DTYPE)
    ai = np.ascontiguousarray(scene.attribute_indices, dtype="<u8")
    tx = np.ascontiguousarray(scene.textures, dtype="<u4")
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, "wb") as f:
        f.write(_header(v.shape[0]))
        f.write(v.tobytes())
        f.write(_header(vi.shape[0]))
        f.write(vi.tobytes())
        if vi.shape[0] % 2:
            f.write(b"\0" * 8)
        f.write(_header(at.shape[0]))
        f.write(at.tobytes())
        f.write(_header(ai.shape[0]))
        f.write(ai.tobytes())
        if ai.shape[0] % 2:
            f.write(b"\0" * 8)
        f.write(_header(tx.size))
        f.write(tx.tobytes())
        return f.tell()


def read_data_bin(path: str) -> Scene:
    with open(path, "rb") as f:
        raw = f.read()
    off = 0

    def header() -> int:
        nonlocal off
        n = int(np.frombuffer(raw, "<u8", 2, off)[0])  # second word ignored (render.cpp:178)
        off += 16
        return n

    def take(dtype, count):
        nonlocal off
        a = np.frombuffer(raw, dtype, count, off).copy()
        off += a.nbytes
        return a

    nv = header()
    v = take("<f4", nv * 4).reshape(nv, 4)
    ni = header()
    vi = take("<u8", ni)
    off += 8 * (ni % 2)
    na = header()
    at = take(ATTR_DTYPE, na)
    ni2 = header()
    ai = take("<u8", ni2)
    off += 8 * (ni2 % 2)
    nt = header()
    tx = take("<u4", nt).reshape(-1, ATLAS, ATLAS)
    return Scene(v, vi, at, ai, tx)


def validate(scene: Scene) -> List[str]:
    """Returns a list of problems that would make the reference misbehave (empty == OK)."""
    problems: List[str] = []
    c = scene.counts()
    if c["I"] % 3:
        problems.append("index count is not a multiple of 3")
    if scene.attribute_indices.shape[0] != c["I"]:
        problems.append("vertex/attribute index streams differ in length")
    if c["I"] and int(scene.vertex_indices.max()) >= c["V"]:
        problems.append("vertex index out of range")
    if c["I"] and int(scene.attribute_indices.max()) >= c["A"]:
        problems.append("attribute index out of range")
    if not np.all(scene.vertices[:, 3] == 1.0):
        problems.append("vertex w != 1")
    kinds = scene.attributes["kind"]
    if np.any(kinds > 1):
        problems.append("attribute kind not in {0,1} (reference calls an empty std::function)")
    tex = kinds == KIND_TEXTURE
    if np.any(tex):
        uv = scene.attributes["payload"][tex][:, 2:4].view("<f4")
        if np.any(uv < 0) or not np.all(np.isfinite(uv)):
            problems.append("negative/non-finite uv: float->uint32 cast is UB (render.cpp:128-129)")
        if int(scene.attributes["payload"][tex][:, 0].max()) >= c["textures"]:
            problems.append("texture index out of range")
    col = kinds == KIND_COLOR
    if np.any(col):
        rgb = scene.attributes["payload"][col][:, 0:3].view("<f4")
        if np.any(rgb < 0) or np.any(rgb > 255):
            problems.append("colour outside 0..255")
    if np.any(scene.attributes["normal"][:, 3] != 0):
        problems.append("normal w != 0")
    if scene.textures.ndim != 3 or scene.textures.shape[1:] != (ATLAS, ATLAS):
        problems.append("textures must be (n, 512, 512)")
    return problems


# ----------------------------------------------------------------------------------------
# Attribute helpers
# ----------------------------------------------------------------------------------------
def color_attr(normal: Sequence[float], rgb: Sequence[float]) -> np.ndarray:
    a = np.zeros((), ATTR_DTYPE)
    a["normal"][:3] = np.asarray(normal, "<f4")
    a["payload"][:3] = np.asarray(rgb, "<f4").view("<u4")
    a["kind"] = KIND_COLOR
    return a


def texture_attr(normal: Sequence[float], index: int, uv: Sequence[float]) -> np.ndarray:
    a = np.zeros((), ATTR_DTYPE)
    a["normal"][:3] = np.asarray(normal, "<f4")
    a["payload"][0] = index
    a["payload"][2:4] = np.asarray(uv, "<f4").view("<u4")
    a["kind"] = KIND_TEXTURE
    return a


class _Builder:
    def __init__(self) -> None:
        self.v: List[np.ndarray] = []
        self.vi: List[int] = []
        self.at: List[np.ndarray] = []
        self.ai: List[int] = []

    @property
    def nv(self) -> int:
        return len(self.v)

    def add_vertices(self, pts: Iterable[Sequence[float]]) -> int:
        base = len(self.v)
        for p in pts:
            self.v.append(np.asarray(p, "<f4"))
        return base

    def add_attrs(self, attrs: Iterable[np.ndarray]) -> int:
        base = len(self.at)
        self.at.extend(attrs)
        return base

    def scene(self, textures: np.ndarray) -> Scene:
        v = np.ones((len(self.v), 4), "<f4")
        if self.v:
            v[:, :3] = np.stack(self.v)
        at = np.array(self.at, dtype=ATTR_DTYPE) if self.at else np.zeros(0, ATTR_DTYPE)
        return Scene(v, np.asarray(self.vi, "<u8"), at, np.asarray(self.ai, "<u8"), textures)


def _f32(x) -> np.ndarray:
    return np.asarray(x, dtype=np.float32)


def _normalize(v: np.ndarray) -> np.ndarray:
    v = _f32(v)
    return _f32(v / np.sqrt(np.sum(v * v, dtype=np.float32), dtype=np.float32))


def face_normal(v: Sequence[np.ndarray], a: int, b: int, c: int) -> np.ndarray:
    """normalize(cross(v[c]-v[a], v[b]-v[a])) — data-generator/main.swift:69-72."""
    return _normalize(np.cross(_f32(v[c]) - _f32(v[a]), _f32(v[b]) - _f32(v[a])))


def random_unit_axes(rng: np.random.RandomState) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Random orthonormal frame — data-generator/main.swift:15-32, seeded here."""

    def sphere_point() -> np.ndarray:
        cz = np.float32(rng.uniform(-1, 1))
        ang = np.float32(rng.uniform(0, 2 * np.pi))
        s = np.sqrt(np.float32(1) - cz * cz)
        return _f32([np.cos(ang) * s, np.sin(ang) * s, cz])

    x = sphere_point()
    while True:
        q = sphere_point()
        if not (np.all(q == x) or np.all(q == -x)):
            break
    y = _normalize(np.cross(x, q))
    z = _f32(np.cross(x, y))
    return x, y, z


# ----------------------------------------------------------------------------------------
# Scene pieces
# ----------------------------------------------------------------------------------------
def add_simple_floor(b: _Builder, a: int = 30, texture: int = 0) -> None:
    i = b.add_vertices(
        [(-a / 2, -0.5, -a - 2.0), (a / 2, -0.5, -a - 2.0), (-a / 2, -0.5, -2.0), (a / 2, -0.5, -2.0)]
    )
    s = np.float32(15) / np.float32(a)
    t1, t2, t3, t4 = (0, 0), (a * s, 0), (0, a * s), (a * s, a * s)
    b.vi += [i, i + 1, i + 2, i + 2, i + 1, i + 3]
    n = (0, 1, 0)
    j = b.add_attrs([texture_attr(n, texture, t) for t in (t1, t2, t3, t3, t2, t4)])
    b.ai += list(range(j, j + 6))


def add_triangle(b: _Builder, r: float = 1.0, p=(0, 0, -10), texture: int = 1) -> None:
    h = np.sqrt(np.float32(3)) / 2
    v = [_f32([-h, -0.5, 0]), _f32([0, 1, 0]), _f32([h, -0.5, 0])]
    v = [np.float32(r) * q + _f32(p) for q in v]
    i = b.add_vertices(v)
    b.vi += [i, i + 1, i + 2]
    n = face_normal(v, 0, 1, 2)
    j = b.add_attrs([texture_attr(n, texture, uv) for uv in ((0, h), (0.5, 0), (1, h))])
    b.ai += [j, j + 1, j + 2]


def add_regular_floor(b: _Builder, a: int = 30, texture: int = 1, scale: float = 1.0) -> None:
    """The 2*a*a-triangle strip floor the reference ships disabled (main.swift:108-188)."""
    i = b.nv
    for z in range(a + 1):
        for x in range(a + 1):
            b.add_vertices([(x - a / 2 + 0.5 * (z % 2), -0.5, -z - 2.0)])
    n = (0, 1, 0)
    s = np.float32(scale)
    for z in range(a):
        a1 = i + z * (a + 1)
        a2 = i + (z + 1) * (a + 1)
        for x in range(a):
            xs = np.fmod(np.float32(x) * s, np.float32(1))
            ys = np.fmod(np.float32(a - z - 1) * s, np.float32(1))
            if z % 2 == 0:
                b.vi += [a1 + x, a2 + x, a1 + 1 + x, a1 + 1 + x, a2 + x, a2 + 1 + x]
                uvs = [(xs, ys + s), (xs + 0.5 * s, ys), (xs + s, ys + s),
                       (xs + s, ys + s), (xs + 0.5 * s, ys), (xs + 1.5 * s, ys)]
            else:
                b.vi += [a1 + x, a2 + x, a2 + 1 + x, a2 + 1 + x, a1 + 1 + x, a1 + x]
                uvs = [(xs + 0.5 * s, ys + s), (xs, ys), (xs + s, ys),
                       (xs + s, ys), (xs + 1.5 * s, ys + s), (xs + 0.5 * s, ys + s)]
            j = b.add_attrs([texture_attr(n, texture, uv) for uv in uvs])
            b.ai += list(range(j, j + 6))


TETRA_FACES = [(0, 2, 1), (0, 3, 2), (0, 1, 3), (1, 2, 3)]
TETRA_COLORS = [
    (ORANGE, ORANGE, ORANGE), (RED, ORANGE, ORANGE), (ORANGE, ORANGE, BLUE), (ORANGE, ORANGE, ORANGE),
]

ICOSA_FACES = [
    (0, 1, 4), (4, 8, 0), (0, 8, 9), (9, 6, 0), (0, 6, 1), (1, 10, 4), (4, 10, 5), (5, 8, 4),
    (5, 2, 8), (8, 2, 9), (9, 2, 7), (7, 6, 9), (7, 11, 6), (6, 11, 1), (1, 11, 10), (3, 5, 10),
    (10, 11, 3), (3, 11, 7), (7, 2, 3), (3, 2, 5),
]
ICOSA_SPECIAL = {3: (RED, ORANGE, ORANGE), 8: (BLUE, ORANGE, RED), 15: (RED, ORANGE, ORANGE)}


def tetra_vertices(axes, r: float, p) -> List[np.ndarray]:
    x, y, z = axes
    k1, k2, k3 = np.sqrt(_f32(8 / 9)), np.sqrt(_f32(2 / 9)), np.sqrt(_f32(2 / 3))
    v = [z, k1 * x - z / 3, -k2 * x + k3 * y - z / 3, -k2 * x - k3 * y - z / 3]
    return [_f32(np.float32(r) * _f32(q) + _f32(p)) for q in v]


def icosa_vertices(axes, r: float, p) -> List[np.ndarray]:
    x, y, z = axes
    phi = (np.sqrt(np.float32(5)) + 1) / 2
    l = np.float32(1) / np.sqrt(phi + 2)
    k = phi * l
    v = [k * x + l * y, k * x - l * y, -k * x + l * y, -k * x - l * y,
         l * x + k * z, -l * x + k * z, l * x - k * z, -l * x - k * z,
         k * y + l * z, k * y - l * z, -k * y + l * z, -k * y - l * z]
    return [_f32(np.float32(r) * _f32(q) + _f32(p)) for q in v]


def add_tetrahedron(b: _Builder, rng, r: float = 2.0, p=(-10, 5, -10)) -> None:
    v = tetra_vertices(random_unit_axes(rng), r, p)
    i = b.add_vertices(v)
    for f, cols in zip(TETRA_FACES, TETRA_COLORS):
        b.vi += [i + f[0], i + f[1], i + f[2]]
        n = face_normal(v, *f)
        j = b.add_attrs([color_attr(n, c) for c in cols])
        b.ai += [j, j + 1, j + 2]


def add_icosahedron(b: _Builder, rng, r: float = 2.0, p=(10, 5, -10)) -> None:
    v = icosa_vertices(random_unit_axes(rng), r, p)
    i = b.add_vertices(v)
    for k, f in enumerate(ICOSA_FACES):
        b.vi += [i + f[0], i + f[1], i + f[2]]
        n = face_normal(v, *f)
        cols = ICOSA_SPECIAL.get(k, (ORANGE, ORANGE, ORANGE))
        j = b.add_attrs([color_attr(n, c) for c in cols])
        b.ai += [j, j + 1, j + 2]


# ----------------------------------------------------------------------------------------
# Rip-map atlases (render-cpp/render.cpp:124-131 addressing; README.md:21)
# ----------------------------------------------------------------------------------------
def build_ripmap_atlas(img: np.ndarray) -> np.ndarray:
    """256x256x3 u8 image -> 512x512 u32 ``0x00RRGGBB`` atlas.

    Level (Lx, Ly), Lx, Ly in {256, 128, ..., 1}, is the box-filtered image at Lx x Ly texels and
    lives at column offset ``511 & ~(2*Lx-1)`` and row offset ``511 & ~(2*Ly-1)``; row and column
    511 stay white.
    """
    img = np.asarray(img)
    assert img.shape == (256, 256, 3)
    atlas = np.full((ATLAS, ATLAS, 3), 255.0, np.float64)
    base = img.astype(np.float64)
    ly = 256
    rows = base
    while ly >= 1:
        yo = 511 & ~(2 * ly - 1)
        lx = 256
        cur = rows
        while lx >= 1:
            xo = 511 & ~(2 * lx - 1)
            atlas[yo:yo + ly, xo:xo + lx] = cur
            if lx > 1:
                cur = 0.5 * (cur[:, 0::2] + cur[:, 1::2])
            lx //= 2
        if ly > 1:
            rows = 0.5 * (rows[0::2] + rows[1::2])
        ly //= 2
    a = np.clip(np.rint(atlas), 0, 255).astype(np.uint32)
    return (a[..., 0] << 16) | (a[..., 1] << 8) | a[..., 2]


def procedural_image(kind: int) -> np.ndarray:
    """Deterministic 256x256 RGB test images with detail at every scale."""
    y, x = np.mgrid[0:256, 0:256]
    if kind % 2 == 0:
        r = 128 + 100 * np.sin(x * 0.11) * np.cos(y * 0.07)
        g = 128 + 90 * np.sin((x + y) * 0.05)
        bl = 80 + 120 * (((x // 16) + (y // 16)) % 2)
    else:
        rs = np.random.RandomState(1234 + kind)
        noise = rs.randint(0, 64, size=(256, 256))
        r = 60 + 150 * (((x // 8) % 2) ^ ((y // 32) % 2)) + noise
        g = 200 - 0.6 * y + 0.3 * noise
        bl = 40 + 0.7 * x
    img = np.stack([r, g, bl], -1)
    return np.clip(img, 0, 255).astype(np.uint8)


def procedural_textures(n: int = 2) -> np.ndarray:
    return np.stack([build_ripmap_atlas(procedural_image(k)) for k in range(n)])


def load_ppm_atlases(directory: str) -> np.ndarray:
    """Reads pre-built 512x512 P6 atlases the way the reference generator does
    (sorted by name, 15-byte header skipped — data-generator/main.swift:398-414)."""
    out = []
    for name in sorted(os.listdir(directory)):
        with open(os.path.join(directory, name), "rb") as f:
            raw = f.read()[15:]
        rgb = np.frombuffer(raw, np.uint8)[: TEXELS_PER_ATLAS * 3].reshape(ATLAS, ATLAS, 3).astype(np.uint32)
        out.append((rgb[..., 0] << 16) | (rgb[..., 1] << 8) | rgb[..., 2])
    return np.stack(out)


REFERENCE_PPM_DIR = "/root/reference/data-generator/ppms"


def default_textures(prefer_reference: bool = True) -> Tuple[np.ndarray, str]:
    if prefer_reference and os.path.isdir(REFERENCE_PPM_DIR):
        return load_ppm_atlases(REFERENCE_PPM_DIR), "reference-ppm"
    return procedural_textures(2), "procedural"


# ----------------------------------------------------------------------------------------
# Whole scenes
# ----------------------------------------------------------------------------------------
def shipped_scene(seed: int = 1, textures: Optional[np.ndarray] = None, regular_floor: bool = False) -> Scene:
    """The composition at data-generator/main.swift:375-379 (39 vertices, 51 triangles)."""
    rng = np.random.RandomState(seed)
    b = _Builder()
    if regular_floor:
        add_regular_floor(b)
    else:
        add_simple_floor(b)
    add_triangle(b)
    for _ in range(2):
        add_tetrahedron(b, rng)
    for _ in range(2):
        add_icosahedron(b, rng)
    return b.scene(procedural_textures(2) if textures is None else textures)


def _batched_axes(rng: np.random.RandomState, n: int):
    cz = rng.uniform(-1, 1, n).astype(np.float32)
    ang = rng.uniform(0, 2 * np.pi, n).astype(np.float32)
    s = np.sqrt(1 - cz * cz)
    x = np.stack([np.cos(ang) * s, np.sin(ang) * s, cz], -1).astype(np.float32)
    cz = rng.uniform(-1, 1, n).astype(np.float32)
    ang = rng.uniform(0, 2 * np.pi, n).astype(np.float32)
    s = np.sqrt(1 - cz * cz)
    q = np.stack([np.cos(ang) * s, np.sin(ang) * s, cz], -1).astype(np.float32)
    y = np.cross(x, q)
    y = (y / np.linalg.norm(y, axis=-1, keepdims=True)).astype(np.float32)
    z = np.cross(x, y).astype(np.float32)
    return x, y, z


def icosahedron_field(
    n: int,
    seed: int = 7,
    extent=100.0,
    r_range: Tuple[float, float] = (1.0, 10.0),
    textured: bool = True,
    shared_vertices: bool = True,
    textures: Optional[np.ndarray] = None,
    center=(0.0, 0.0, 0.0),
    n_textures: int = 2,
) -> Scene:
    """``n`` random icosahedrons (20 n triangles): r ~ U[r_range], centre ~ U[-extent, extent]^3
    (the reference author's disabled randomisation: data-generator/main.swift:13,229-230,279-280).
    Per-corner attribute records (A = 60 n); vertices shared (V = 12 n) or unshared (V = 60 n,
    needed when many triangles straddle the near plane — reference scratch capacity 2V, H13)."""
    rng = np.random.RandomState(seed)
    x, y, z = _batched_axes(rng, n)
    phi = (np.sqrt(np.float32(5)) + 1) / 2
    l = np.float32(1) / np.sqrt(phi + 2)
    k = phi * l
    X, Y, Z = x[:, None, :], y[:, None, :], z[:, None, :]
    cx = np.array([k, k, -k, -k, l, -l, l, -l, 0, 0, 0, 0], np.float32)[None, :, None]
    cy = np.array([l, -l, l, -l, 0, 0, 0, 0, k, k, -k, -k], np.float32)[None, :, None]
    cz = np.array([0, 0, 0, 0, k, k, -k, -k, l, -l, l, -l], np.float32)[None, :, None]
    unit = cx * X + cy * Y + cz * Z  # (n, 12, 3)
    r = rng.uniform(r_range[0], r_range[1], n).astype(np.float32)[:, None, None]
    ext = np.broadcast_to(np.asarray(extent, np.float32), (3,))
    p = ((rng.uniform(-1, 1, (n, 3)) * ext).astype(np.float32) + np.asarray(center, np.float32))[:, None, :]
    pos = (r * unit + p).astype(np.float32)  # (n, 12, 3)

    faces = np.asarray(ICOSA_FACES, np.int64)  # (20, 3)
    fa, fb, fc = pos[:, faces[:, 0]], pos[:, faces[:, 1]], pos[:, faces[:, 2]]  # (n, 20, 3)
    nrm = np.cross(fc - fa, fb - fa)
    nrm = (nrm / np.linalg.norm(nrm, axis=-1, keepdims=True)).astype(np.float32)

    if shared_vertices:
        verts = np.ones((n * 12, 4), "<f4")
        verts[:, :3] = pos.reshape(-1, 3)
        vi = (np.arange(n, dtype=np.uint64)[:, None, None] * 12 + faces[None].astype(np.uint64)).reshape(-1)
    else:
        verts = np.ones((n * 60, 4), "<f4")
        verts[:, :3] = np.stack([fa, fb, fc], 2).reshape(-1, 3)
        vi = np.arange(n * 60, dtype=np.uint64)

    at = np.zeros((n, 20, 3), ATTR_DTYPE)
    at["normal"][..., :3] = nrm[:, :, None, :]
    if textured:
        h = np.sqrt(np.float32(3)) / 2
        uv = np.array([[0, h], [0.5, 0], [1, h]], np.float32)  # addTriangle pattern (main.swift:101-103)
        at["kind"] = KIND_TEXTURE
        at["payload"][..., 0] = (np.arange(n, dtype=np.uint32) % n_textures)[:, None, None]
        at["payload"][..., 2:4] = uv.view("<u4")[None, None]
    else:
        at["kind"] = KIND_COLOR
        col = rng.uniform(0, 255, (n, 20, 3, 3)).astype(np.float32)
        at["payload"][..., :3] = col.view("<u4")
    ai = np.arange(n * 60, dtype=np.uint64)
    tx = procedural_textures(n_textures) if textures is None else textures
    return Scene(verts, vi, at.reshape(-1), ai, tx)


def concat_scenes(a: Scene, b: Scene) -> Scene:
    """Appends scene ``b`` to ``a`` (indices re-based; both must use the same texture set)."""
    assert a.textures.shape == b.textures.shape and np.array_equal(a.textures, b.textures)
    return Scene(
        np.concatenate([a.vertices, b.vertices]),
        np.concatenate([a.vertex_indices, b.vertex_indices + np.uint64(a.vertices.shape[0])]),
        np.concatenate([a.attributes, b.attributes]),
        np.concatenate([a.attribute_indices, b.attribute_indices + np.uint64(a.attributes.shape[0])]),
        a.textures,
    )


def clip_stress_scene(n: int = 50000, seed: int = 11, textures: Optional[np.ndarray] = None) -> Scene:
    """C4: the camera (origin, looking down -z) sits inside a dense field.  70 % of the solids lie in a
    thin slab around the near plane z = -0.1 (about half of their triangles straddle it), 30 % in front of
    the camera so that clipped and unclipped geometry is actually visible.  Unshared vertices keep the
    reference inside its 2x scratch arrays (hazard H13).  Measured straddle fraction: ~36 %."""
    tx = procedural_textures(2) if textures is None else textures
    n_slab = int(n * 0.7)
    slab = icosahedron_field(n_slab, seed=seed, extent=(40.0, 25.0, 1.5), r_range=(2.0, 3.0), textured=True,
                             shared_vertices=False, textures=tx)
    front = icosahedron_field(n - n_slab, seed=seed + 1, extent=(30.0, 17.0, 25.0), r_range=(0.3, 1.5), textured=True,
                              shared_vertices=False, textures=tx, center=(0.0, 0.0, -28.0))
    return concat_scenes(slab, front)


def hazard_scene(textures: Optional[np.ndarray] = None) -> Scene:
    """A handful of triangles aimed at the reference's quirks (SURVEY.md 8 hazard list), seen from the initial camera
    (origin, looking down -z): H2 exact depth ties (the same triangle twice with different colours, and two coplanar
    triangles overlapping in part: the first in list order must win), H3 a fan with shared edges and a shared apex (pixels
    on an edge belong to both neighbours), H4 slivers and dots around the `area < 10` cull, back faces, H5 kind taken from
    corner 0, H9 vertices exactly on and just around the near plane z = -0.1 (on the plane counts as behind), H12 a textured
    triangle with constant uv (rip-map level inf -> 256), a huge triangle running far off screen (long incremental walks),
    triangles partly and wholly outside the frame."""
    tx = procedural_textures(2) if textures is None else textures
    b = _Builder()
    n = [0.0, 0.0, 1.0]

    def tri(pts, attrs):
        v0 = b.add_vertices(pts)
        a0 = b.add_attrs(attrs)
        b.vi += [v0, v0 + 1, v0 + 2]
        b.ai += [a0, a0 + 1, a0 + 2]

    def col(rgb, normal=n):
        return color_attr(normal, rgb)

    # H2: identical triangles, different colours — the first must own every pixel
    quad = [(-3.0, 1.0, -10.0), (-3.0, 3.0, -10.0), (-1.0, 1.0, -10.0)]
    tri(quad, [col((255, 0, 0))] * 3)
    tri(quad, [col((0, 255, 0))] * 3)
    # H2: coplanar, partly overlapping, later one larger
    tri([(0.0, 1.0, -10.0), (0.0, 3.0, -10.0), (2.0, 1.0, -10.0)], [col((0, 0, 255))] * 3)
    tri([(-0.5, 0.5, -10.0), (-0.5, 3.5, -10.0), (3.0, 0.5, -10.0)], [col((255, 255, 0))] * 3)
    # H3: a fan of six triangles around a shared apex, shared edges, smooth normals
    import math
    c = (3.5, -1.5, -8.0)
    ring = [(c[0] + 1.5 * math.cos(k * math.pi / 3), c[1] + 1.5 * math.sin(k * math.pi / 3), c[2] - 0.3 * (k % 2)) for k in range(6)]
    for k in range(6):
        p, q = ring[(k + 1) % 6], ring[k]
        tri([c, p, q], [col((40 * k + 30, 200 - 30 * k, 90), _normalize([0.2 * (k - 2), 0.1, 1.0]))] * 3)
    # H4: slivers and dots around the area cull, and a back face (negative area)
    for k in range(8):
        x, s_ = -4.0 + 0.45 * k, 0.004 * (k + 1)
        tri([(x, -1.0, -6.0), (x, -1.0 + 30 * s_, -6.0), (x + s_, -1.0, -6.0)], [col((250, 250, 250))] * 3)
        tri([(x, -2.0, -6.0), (x, -2.0 + s_ * 4, -6.0), (x + s_ * 4, -2.0, -6.0)], [col((250, 120, 250))] * 3)
    tri([(-3.0, -3.0, -9.0), (-1.0, -3.0, -9.0), (-3.0, -1.5, -9.0)], [col((10, 250, 250))] * 3)   # wound the other way
    # H5: kind from corner 0 — a textured triangle and a coloured one whose other corners carry the other kind
    # (the reference reads the other corners' records through the same union: a colour record's blue is a texture
    # record's u, a texture record's index bits and u are a colour record's red and blue)
    tri([(-1.0, -3.2, -7.0), (-1.0, -1.7, -7.0), (0.5, -3.2, -7.0)],
        [texture_attr(n, 1, (0.0, 0.0)), col((50, 100, 2.0)), texture_attr(n, 0, (2.0, 1.5))])
    tri([(0.7, -3.2, -7.0), (0.7, -1.7, -7.0), (2.2, -3.2, -7.0)],
        [col((200, 100, 50)), texture_attr(n, 1, (200.0, 3.0)), col((100, 200, 50))])
    # H9: vertices exactly on the near plane (z = -0.1: behind), one ulp in front of it, and a straddler through it
    near = np.float32(-0.1)
    tri([(-0.02, 0.0, near), (-0.02, 0.03, near), (0.0, 0.0, -0.3)], [col((255, 128, 0))] * 3)            # two corners on the plane
    tri([(0.01, 0.0, np.nextafter(near, np.float32(-1))), (0.01, 0.03, -0.2), (0.03, 0.0, -0.3)], [col((128, 255, 0))] * 3)
    tri([(-0.05, -0.04, 0.05), (-0.05, -0.01, -0.4), (-0.02, -0.04, -0.4)], [col((0, 128, 255))] * 3)     # one corner behind the eye
    # H12: constant uv -> no texture gradient -> level = ooz / 0 = inf -> clamped to 256
    tri([(2.5, 1.0, -9.0), (2.5, 2.5, -9.0), (4.0, 1.0, -9.0)], [texture_attr(n, 0, (0.25, 0.75))] * 3)
    # long walks: a floor-like textured triangle running far beyond the frame on three sides
    tri([(-400.0, -4.0, -2.0), (0.0, -4.0, -900.0), (400.0, -4.0, -2.0)],
        [texture_attr([0.0, 1.0, 0.0], 0, (0.0, 0.0)), texture_attr([0.0, 1.0, 0.0], 0, (40.0, 90.0)), texture_attr([0.0, 1.0, 0.0], 0, (80.0, 0.0))])
    # partly and wholly outside the frame
    tri([(5.5, 2.0, -9.0), (5.5, 4.5, -9.0), (9.0, 2.0, -9.0)], [col((90, 90, 250))] * 3)
    tri([(30.0, 0.0, -9.0), (30.0, 2.0, -9.0), (33.0, 0.0, -9.0)], [col((250, 90, 90))] * 3)
    return b.scene(tx)


def c3_scene(n: int = 1_000_000, seed: int = 7, textures: Optional[np.ndarray] = None) -> Scene:
    """C3: n textured icosahedrons (20 n triangles, shared vertices, per-corner attributes) filling the
    view frustum around z = -4000 so that ~21 % of the triangles survive the area >= 10 cull at 4K from
    the origin (10-30 %) with a depth complexity of ~7."""
    return icosahedron_field(n, seed=seed, extent=(1500.0, 850.0, 800.0), r_range=(1.0, 10.0), textured=True,
                             shared_vertices=True, textures=textures, center=(0.0, 0.0, -4000.0))


# ----------------------------------------------------------------------------------------
# Scripted Input sequences (replace input.swift; value ranges per input.swift:30-91)
# ----------------------------------------------------------------------------------------
def input_script(name: str, frames: int) -> np.ndarray:
    """Deterministic ``Input`` records.  ``mouse`` is the absolute accumulated position."""
    inp = np.zeros(frames, INPUT_DTYPE)
    mouse = np.zeros(2, np.float64)
    for f in range(frames):
        t = f / max(frames - 1, 1)
        if name == "still":
            pass
        elif name == "c1_path":
            if f < 60:
                inp[f]["down"] = 1
            elif f < 120:
                inp[f]["right"] = 1
            else:
                mouse += (2.0, -1.0)
                inp[f]["up"] = float(f % 2)
        elif name == "flythrough":  # C2 path, planned below (needs camera feedback)
            return _flythrough(frames)
        elif name == "strafe":  # C4: translate only, so the near plane keeps cutting the slab of solids
            if (f // 8) % 2 == 0:
                inp[f]["right"] = 1
            else:
                inp[f]["left"] = 1
        elif name == "spin":
            mouse += (11.0, 3.0 * np.sin(f * 0.1))
            inp[f]["up"] = 1 if (f // 20) % 2 == 0 else 0
            inp[f]["down"] = 0 if (f // 20) % 2 == 0 else 1
        else:
            raise ValueError(name)
        inp[f]["mouse"] = mouse.astype(np.float32)
    return inp


class _PlanCam:
    """float64 planning model of update_camera (render-cpp/render.cpp:134-156); only used to
    *choose* inputs — the emitted Input records are what every renderer consumes."""

    def __init__(self) -> None:
        self.pos = np.zeros(3)
        self.X, self.Y, self.Z = np.eye(3)
        self.m0 = np.zeros(2)

    @staticmethod
    def _act(q, v):
        t = 2 * np.cross(q[:3], v)
        return v + q[3] * t + np.cross(q[:3], t)

    def step(self, up, down, left, right, m) -> None:
        if max(up, down, left, right) > 0:
            self.pos = self.pos + 0.1 * ((right - left) * self.X + (down - up) * self.Z)
        m = np.asarray(m, float)
        if np.any(m != self.m0):
            z = (self.m0[0] - m[0]) * self.X + (self.m0[1] - m[1]) * self.Y + (100 / 0.3) * self.Z
            z /= np.linalg.norm(z)
            h = self.Z + z
            h /= np.linalg.norm(h)
            q = np.append(np.cross(self.Z, h), self.Z @ h)
            self.X = self._act(q, self.X)
            self.X /= np.linalg.norm(self.X)
            self.Y = self._act(q, self.Y)
            self.Y /= np.linalg.norm(self.Y)
            self.Z = z
            self.m0 = m


def _flythrough(frames: int) -> np.ndarray:
    """600-frame recorded fly-through for the 4K config: (i) look down at the floor (magnified
    texels), (ii) dive and graze it (deep rip-map levels, near-clip of floor + textured triangle),
    (iii) fly through the icosahedrons, then the tetrahedrons (near-clip of coloured solids),
    (iv) retreat to an overview."""
    cam = _PlanCam()
    mouse = np.zeros(2)
    inp = np.zeros(frames, INPUT_DTYPE)
    ico = np.array([10.0, 5.0, -10.0])
    tet = np.array([-10.0, 5.0, -10.0])

    def steer(direction, rate):
        d = np.asarray(direction, float)
        n = np.linalg.norm(d)
        if n < 1.0:  # too close to the target for a stable bearing: hold course
            return np.zeros(2)
        d = d / n
        return np.clip(np.array([d @ cam.X, d @ cam.Y]) * 333.0 * 0.5, -rate, rate)

    for f in range(frames):
        g = f * 600 // max(frames, 1) if frames < 600 else f
        up = down = left = right = 0.0
        if g < 40:  # advance over the floor
            up = 2.0
            dm = np.zeros(2)
        elif g < 90:  # look down: magnified texels
            dm = steer((0, -0.6, -1), 3.0)
        elif g < 140:  # back away and up: minified floor overview
            down = 2.0
            dm = steer((0, -0.6, -1), 3.0)
        elif g < 200:  # dive, flattening out just above the floor
            up = 2.0
            dm = steer((0, -min(max((cam.pos[1] + 0.3) * 0.2, 0.0), 0.6), -1), 5.0)
        elif g < 290:  # graze the floor, through the textured triangle (near-clip)
            up = 1.0
            dm = steer((0, -min(max((cam.pos[1] + 0.3) * 0.2, -0.05), 0.6), -1), 5.0)
        elif g < 400:  # turn, then climb into the icosahedrons
            to = ico - cam.pos
            up = 2.0 if (to / max(np.linalg.norm(to), 1e-9)) @ (-cam.Z) > 0.97 or np.linalg.norm(to) < 1.0 else 0.0
            dm = steer(to, 12.0)
        elif g < 540:  # turn to the tetrahedrons and fly through them
            to = tet - cam.pos
            up = 2.0 if (to / max(np.linalg.norm(to), 1e-9)) @ (-cam.Z) > 0.97 or np.linalg.norm(to) < 1.0 else 0.0
            dm = steer(to, 12.0)
        else:  # retreat to an overview
            down = 2.0
            dm = steer(np.array([0.0, 0.0, -12.0]) - cam.pos, 8.0)
        mouse = np.round(mouse + dm, 3)
        cam.step(up, down, left, right, mouse)
        inp[f]["up"], inp[f]["down"], inp[f]["left"], inp[f]["right"] = up, down, left, right
        inp[f]["mouse"] = mouse.astype(np.float32)
    return inp
