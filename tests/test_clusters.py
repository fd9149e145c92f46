"""The load-time spatial pre-partition of the triangle stream (csrc/cluster.hpp), checked on the host: it is a
re-arrangement (every triangle exactly once, positions intact, the reference's order recoverable) and its bounds — which the
front kernel's whole-cluster rejections rest on — really bound."""
import numpy as np
import pytest

from swift3drenderer_b200 import renderer as R, scene as S


def _check(sc):
    c = R.debug_clusters(sc)
    hdr, pos, tri = c["hdr"], c["pos"], c["tri"]
    T = sc.n_triangles
    n = len(hdr) - 1
    assert hdr["tri_off"][0] == 0 and hdr["v_off"][0] == 0 and hdr["tri_off"][n] == T and hdr["v_off"][n] == pos.shape[1]
    n_tris = np.diff(hdr["tri_off"].astype(np.int64)); n_verts = np.diff(hdr["v_off"].astype(np.int64))
    assert (n_tris >= 1).all() and (n_tris <= 32).all() and (n_verts >= 1).all() and (n_verts <= 16).all()
    # every original triangle exactly once: cluster k holds t0 .. t0 + n_tris - 1
    owner = np.repeat(np.arange(n), n_tris)
    local = np.arange(T) - np.repeat(hdr["tri_off"][:n].astype(np.int64), n_tris)
    orig = hdr["t0"][owner].astype(np.int64) + local
    assert np.array_equal(np.sort(orig), np.arange(T))
    # the triangle words address the cluster's private vertex copies, which equal the original positions
    vi = np.asarray(sc.vertex_indices, np.int64).reshape(-1, 3)
    xyz = np.ascontiguousarray(np.asarray(sc.vertices, np.float32)[:, :3])
    for k in range(3):
        lv = ((tri >> (8 * k)) & 255).astype(np.int64)
        assert (lv < n_verts[owner]).all()
        got = np.ascontiguousarray(pos[:, hdr["v_off"][owner].astype(np.int64) + lv].T)
        assert np.array_equal(got.view(np.uint32), np.ascontiguousarray(xyz[vi[orig, k]]).view(np.uint32))
    assert (tri >> 24 == 0).all()
    # bounds: every vertex inside the sphere, every edge no longer than max_edge (both rounded outwards)
    vown = np.repeat(np.arange(n), n_verts)
    finite = np.isfinite(hdr["radius"][:n])
    d = np.linalg.norm(pos.T.astype(np.float64) - hdr["center"][vown].astype(np.float64), axis=1)
    ok = finite[vown]
    assert (d[ok] <= hdr["radius"][vown][ok].astype(np.float64)).all()
    p = xyz[vi[orig]].astype(np.float64)   # (T, 3, 3)
    edges = np.stack([np.linalg.norm(p[:, a] - p[:, (a + 1) % 3], axis=1) for a in range(3)], 1).max(1)
    okt = finite[owner]
    assert (edges[okt] <= hdr["max_edge"][owner][okt].astype(np.float64)).all()
    return hdr, n_tris, n_verts


def test_icosahedron_field_clusters_are_the_solids():
    sc = S.icosahedron_field(3000, seed=5, extent=200)
    hdr, n_tris, n_verts = _check(sc)
    # consecutive solids are (almost always) far apart: one icosahedron per cluster — 12 vertices, 20 triangles, its own radius;
    # the rare neighbour that is close enough donates a triangle or two before the vertex limit closes the cluster
    assert len(n_tris) == 3000 and (n_tris == 20).mean() > 0.99 and (n_verts == 12).mean() > 0.99
    assert np.median(hdr["radius"][:-1]) < 10.5
    # Morton order: neighbours in the list are neighbours in space (median hop far below the field's extent)
    hop = np.linalg.norm(np.diff(hdr["center"][:-1], axis=0), axis=1)
    assert np.median(hop) < 60


@pytest.mark.parametrize("which", ["shipped", "regfloor", "hazards", "clip", "unshared"])
def test_cluster_invariants(which):
    sc = {"shipped": lambda: S.shipped_scene(1), "regfloor": lambda: S.shipped_scene(2, regular_floor=True),
          "hazards": S.hazard_scene, "clip": lambda: S.clip_stress_scene(400),
          "unshared": lambda: S.icosahedron_field(200, seed=2, extent=30, shared_vertices=False)}[which]()
    hdr, n_tris, n_verts = _check(sc)
    if which == "regfloor":   # a connected grid is cut by the spread rule, not into single triangles
        assert n_tris.mean() > 6


def test_non_finite_vertices_get_bounds_that_never_cull():
    sc = S.icosahedron_field(4, seed=1, extent=10)
    sc.vertices[5, 0] = np.inf
    sc.vertices[30, 1] = np.nan
    c = R.debug_clusters(sc)
    assert np.isinf(c["hdr"]["radius"][:-1]).sum() >= 2   # the front kernel's comparisons against these all fail: per-triangle path
