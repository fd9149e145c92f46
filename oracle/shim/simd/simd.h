/*
 * oracle/shim/simd/simd.h — TEST INFRASTRUCTURE, not product code.
 *
 * Stand-in for Apple's <simd/simd.h> (header-only, macOS 13 / iOS 16 SDK — not present
 * under /root/reference, not installable here), written from scratch so that the
 * reference's render-cpp/render.cpp compiles UNMODIFIED with g++ on Linux.
 *
 * Only the names the reference actually uses are provided (call sites:
 * render-cpp/render.cpp:1,63,126-131,136-155,227-236,286-291,311-313,336-355,363-370
 * and render-cpp/render.hpp:4,20).
 *
 * PARITY NOTE: Apple's simd is third-party arithmetic whose source is not in the
 * reference, so this header *defines* the pinned oracle semantics:
 *   - every operation is plain IEEE-754 binary32, evaluated left to right, no FMA;
 *   - simd_float3 is 16 bytes / 16-aligned, simd_float2 8 bytes / 8-aligned (Apple ABI),
 *     which fixes sizeof(vertex_attribute_t) == 48 and thereby the data.bin layout;
 *   - simd_fast_normalize(v) == v * (1.0f / sqrtf(dot(v, v)))  (exact sqrt and divide; Apple's
 *     is an rsqrt estimate that differs between arm64 and x86-64 — unpinnable);
 *   - simd_quaternion(from, to) follows Apple's documented construction
 *     (half-vector; two half rotations for obtuse angles);
 *   - simd_act(q, v) == v + q.w * t + cross(q.xyz, t), t = 2 * cross(q.xyz, v).
 */
#ifndef ORACLE_SHIM_SIMD_H
#define ORACLE_SHIM_SIMD_H

#include <limits.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

struct alignas(8) simd_float2 {
    float x, y;
    float &operator[](int i) { return (&x)[i]; }
    const float &operator[](int i) const { return (&x)[i]; }
};

struct alignas(16) simd_float3 {
    float x, y, z, pad_;
    float &operator[](int i) { return (&x)[i]; }
    const float &operator[](int i) const { return (&x)[i]; }
};

struct alignas(16) simd_float4 {
    float x, y, z, w;
    float &operator[](int i) { return (&x)[i]; }
    const float &operator[](int i) const { return (&x)[i]; }
};

struct simd_float4x3 {
    simd_float3 columns[4];
};

struct simd_quatf {
    simd_float4 vector; /* xyz imaginary, w real */
};

/* ---- constructors ------------------------------------------------------------------ */
static inline simd_float2 simd_make_float2(float x, float y) { return simd_float2{x, y}; }
static inline simd_float3 simd_make_float3(float x, float y, float z) { return simd_float3{x, y, z, 0.f}; }
static inline simd_float3 simd_make_float3(simd_float2 xy, float z) { return simd_float3{xy.x, xy.y, z, 0.f}; }
static inline simd_float4 simd_make_float4(float x, float y, float z, float w) { return simd_float4{x, y, z, w}; }
static inline simd_float4 simd_make_float4(simd_float3 v, float w) { return simd_float4{v.x, v.y, v.z, w}; }

/* ---- float2 arithmetic ------------------------------------------------------------- */
static inline simd_float2 operator+(simd_float2 a, simd_float2 b) { return simd_float2{a.x + b.x, a.y + b.y}; }
static inline simd_float2 operator-(simd_float2 a, simd_float2 b) { return simd_float2{a.x - b.x, a.y - b.y}; }
static inline simd_float2 operator*(simd_float2 a, simd_float2 b) { return simd_float2{a.x * b.x, a.y * b.y}; }
static inline simd_float2 operator/(simd_float2 a, simd_float2 b) { return simd_float2{a.x / b.x, a.y / b.y}; }
static inline simd_float2 operator*(simd_float2 a, float s) { return simd_float2{a.x * s, a.y * s}; }
static inline simd_float2 operator*(float s, simd_float2 a) { return simd_float2{s * a.x, s * a.y}; }
static inline simd_float2 operator/(simd_float2 a, float s) { return simd_float2{a.x / s, a.y / s}; }
static inline simd_float2 operator/(float s, simd_float2 a) { return simd_float2{s / a.x, s / a.y}; }
static inline simd_float2 operator-(simd_float2 a) { return simd_float2{-a.x, -a.y}; }
static inline simd_float2 &operator+=(simd_float2 &a, simd_float2 b) { a = a + b; return a; }

/* ---- float3 arithmetic ------------------------------------------------------------- */
static inline simd_float3 operator+(simd_float3 a, simd_float3 b) { return simd_float3{a.x + b.x, a.y + b.y, a.z + b.z, 0.f}; }
static inline simd_float3 operator-(simd_float3 a, simd_float3 b) { return simd_float3{a.x - b.x, a.y - b.y, a.z - b.z, 0.f}; }
static inline simd_float3 operator*(simd_float3 a, simd_float3 b) { return simd_float3{a.x * b.x, a.y * b.y, a.z * b.z, 0.f}; }
static inline simd_float3 operator/(simd_float3 a, simd_float3 b) { return simd_float3{a.x / b.x, a.y / b.y, a.z / b.z, 0.f}; }
static inline simd_float3 operator*(simd_float3 a, float s) { return simd_float3{a.x * s, a.y * s, a.z * s, 0.f}; }
static inline simd_float3 operator*(float s, simd_float3 a) { return simd_float3{s * a.x, s * a.y, s * a.z, 0.f}; }
static inline simd_float3 operator/(simd_float3 a, float s) { return simd_float3{a.x / s, a.y / s, a.z / s, 0.f}; }
static inline simd_float3 operator/(float s, simd_float3 a) { return simd_float3{s / a.x, s / a.y, s / a.z, 0.f}; }
static inline simd_float3 operator-(simd_float3 a) { return simd_float3{-a.x, -a.y, -a.z, 0.f}; }
static inline simd_float3 &operator+=(simd_float3 &a, simd_float3 b) { a = a + b; return a; }

/* ---- reductions / geometry --------------------------------------------------------- */
static inline float simd_dot(simd_float3 a, simd_float3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
static inline simd_float3 simd_cross(simd_float3 a, simd_float3 b) {
    return simd_float3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x, 0.f};
}
static inline simd_float3 simd_max(simd_float3 a, simd_float3 b) {
    return simd_float3{fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z), 0.f};
}
static inline simd_float3 simd_min(simd_float3 a, simd_float3 b) {
    return simd_float3{fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z), 0.f};
}
static inline simd_float2 simd_abs(simd_float2 a) { return simd_float2{fabsf(a.x), fabsf(a.y)}; }
static inline simd_float3 simd_fast_normalize(simd_float3 v) { return v * (1.0f / sqrtf(simd_dot(v, v))); }
static inline simd_float3 simd_normalize(simd_float3 v) { return simd_fast_normalize(v); }

/* ---- matrices ---------------------------------------------------------------------- */
static inline simd_float4x3 simd_matrix_from_rows(simd_float4 r0, simd_float4 r1, simd_float4 r2) {
    simd_float4x3 m;
    for (int c = 0; c < 4; c++) { m.columns[c] = simd_float3{r0[c], r1[c], r2[c], 0.f}; }
    return m;
}
static inline simd_float3 simd_mul(simd_float4x3 m, simd_float4 v) {
    return ((m.columns[0] * v.x + m.columns[1] * v.y) + m.columns[2] * v.z) + m.columns[3] * v.w;
}

/* ---- quaternions ------------------------------------------------------------------- */
static inline simd_quatf shim_quat_half_(simd_float3 from, simd_float3 half) {
    simd_float3 c = simd_cross(from, half);
    return simd_quatf{simd_float4{c.x, c.y, c.z, simd_dot(from, half)}};
}
static inline simd_quatf shim_quat_mul_(simd_quatf p, simd_quatf q) {
    simd_float3 pv = simd_float3{p.vector.x, p.vector.y, p.vector.z, 0.f};
    simd_float3 qv = simd_float3{q.vector.x, q.vector.y, q.vector.z, 0.f};
    simd_float3 v = (qv * p.vector.w + pv * q.vector.w) + simd_cross(pv, qv);
    return simd_quatf{simd_float4{v.x, v.y, v.z, p.vector.w * q.vector.w - simd_dot(pv, qv)}};
}
static inline simd_quatf simd_quaternion(simd_float3 from, simd_float3 to) {
    if (simd_dot(from, to) >= 0.f) {
        return shim_quat_half_(from, simd_fast_normalize(from + to));
    }
    /* obtuse: rotate from -> half -> to, half = normalize(from x to) x ... choose any unit vector
       in the plane halfway; Apple composes two half rotations. */
    simd_float3 half = simd_fast_normalize(from + to);
    if (!(simd_dot(half, half) > 0.f)) { /* from == -to: pick an orthogonal axis */
        simd_float3 ax = fabsf(from.x) < fabsf(from.y) ? simd_float3{1, 0, 0, 0} : simd_float3{0, 1, 0, 0};
        half = simd_fast_normalize(simd_cross(from, ax));
    }
    return shim_quat_mul_(shim_quat_half_(half, to), shim_quat_half_(from, half));
}
static inline simd_float3 simd_act(simd_quatf q, simd_float3 v) {
    simd_float3 qv = simd_float3{q.vector.x, q.vector.y, q.vector.z, 0.f};
    simd_float3 t = 2.f * simd_cross(qv, v);
    return (v + q.vector.w * t) + simd_cross(qv, t);
}

/* ---- Apple libc ------------------------------------------------------------------- */
static inline void memset_pattern4(void *dst, const void *pattern4, size_t len) {
    uint32_t p;
    memcpy(&p, pattern4, 4);
    uint32_t *d = (uint32_t *)dst;
    size_t n = len / 4;
    for (size_t i = 0; i < n; i++) { d[i] = p; }
    memcpy((char *)dst + n * 4, &p, len - n * 4);
}

#endif /* ORACLE_SHIM_SIMD_H */
