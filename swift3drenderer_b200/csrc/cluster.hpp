// cluster.hpp — spatial pre-partition of the triangle stream, built once at scene load (host).
//
// The reference walks its triangle list front to back every frame (render-cpp/render.cpp:297-317) and rejects most of a
// large scene triangle by triangle.  Here the list is cut into CLUSTERS — runs of consecutive triangles that sit close
// together (at most CL_MAX_VERTS distinct vertices, CL_MAX_TRIS triangles, bounding box no larger than a few of their
// edges) — each with a bounding sphere and its longest edge, so that a whole cluster can be rejected by one
// conservative test (behind the near plane, off screen, outside the rows a GPU owns, or too small to pass the
// `area < 10` cull) before any of its vertices is transformed.  Clusters are stored in Morton order of their centres (what a
// warp of the front kernel touches is close together in space and in memory); every cluster owns private copies of its
// vertices (streamed, coalesced) and its triangles become one 32-bit word each (three cluster-local vertex numbers).
// The reference's processing order survives as the order key: triangle j of a cluster is original triangle t0 + j.
#pragma once
#include <stdint.h>

#include <vector>

namespace s3r {

constexpr uint32_t CL_MAX_VERTS = 16;    // distinct vertices per cluster (an icosahedron has 12)
constexpr uint32_t CL_MAX_TRIS = 32;     // triangles per cluster (an icosahedron has 20)
constexpr float CL_SPREAD = 4.0f;        // a cluster's bounding-box diagonal stays within CL_SPREAD of its longest edge

struct ClusterHeader {                   // 32 bytes, two 16-byte loads
    float cx, cy, cz, radius;            // bounding sphere of the cluster's vertices (radius rounded up)
    float max_edge;                      // longest triangle edge in object space (rounded up)
    uint32_t t0;                         // original index of the cluster's first triangle (the rest follow consecutively)
    uint32_t v_off;                      // first entry in the cluster-vertex arrays
    uint32_t tri_off;                    // first entry in the triangle-word array
};
static_assert(sizeof(ClusterHeader) == 32, "ClusterHeader must be 32 bytes");

struct ClusterSet {
    std::vector<ClusterHeader> hdr;      // n_clusters + 1: the last one is a sentinel that carries the end offsets
    std::vector<float> px, py, pz;       // cluster-private vertex copies, in cluster order
    std::vector<uint32_t> tri;           // v0 | v1 << 8 | v2 << 16 (cluster-local vertex numbers), in cluster order
    uint32_t n_clusters = 0;
};

// positions: planar float[V]; vi0..2: planar uint32[T] (corner k of triangle t).
void build_clusters(const float *px, const float *py, const float *pz, const uint32_t *vi0, const uint32_t *vi1,
                    const uint32_t *vi2, uint64_t T, ClusterSet &out);

}  // namespace s3r
