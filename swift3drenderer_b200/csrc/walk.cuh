// walk.cuh — exact fast-forward of the reference's incremental barycentric walk.
//
// The reference advances each barycentric weight by *repeated binary32 addition*
// (`weight.w += weight.dx` per pixel, `weight.wy += weight.dy` per row:
// render-cpp/render.cpp:374-379), so the value at pixel (x, y) is
//     fl(...fl(fl(wstart + dy) + dy)... + dx) ... + dx)         (y - ymin, then x - xmin adds)
// and NOT wstart + (x - xmin) * dx + (y - ymin) * dy.  Coverage, depth and texel selection all
// depend on those rounded intermediates, so a tiled rasteriser must reproduce them bit for bit
// when it enters a triangle's bounding box in the middle.
//
// walk_jump(s, d, n) returns exactly the result of n sequential adds in O(#binades crossed):
// while s stays inside one binade (same sign and exponent field) every add moves it by a constant
// whole number of ulps, known from d's bits and the binade's exponent alone (round-to-nearest-even
// needs an even mantissa first in the one binade where d ends in exactly half an ulp), so the steps
// inside the binade collapse into one integer multiply-add on the bit pattern.  Steps that cross a
// binade boundary, zero or the sign are taken as true additions.  Fixed points (d absorbed) end the
// walk early.
//
// Host + device: the same inline function is compiled by g++ for the CPU unit test
// (tests/test_walk_jump.py) and by nvcc for the kernels.
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define S3R_HD __host__ __device__ __forceinline__
#else
#define S3R_HD inline
#endif

namespace s3r {

S3R_HD uint32_t f2u(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    uint32_t u;
    memcpy(&u, &f, 4);
    return u;
#endif
}

S3R_HD float u2f(uint32_t u) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f;
    memcpy(&f, &u, 4);
    return f;
#endif
}

// One true binary32 addition that the compiler may not contract or reassociate.
S3R_HD float add_rn(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fadd_rn(a, b);
#else
    volatile float r = a + b;
    return r;
#endif
}

// floor(a / b) for a < 2^23, 0 < b < 2^23 without an integer divide (both convert to binary32
// exactly; the reciprocal estimate is off by at most one, fixed up with integer arithmetic).
S3R_HD uint32_t div_small(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    uint32_t k = (uint32_t)(__uint2float_rz(a) * __frcp_rn(__uint2float_rz(b)));
    if (k * b > a) { --k; }
    if ((k + 1u) * b <= a) { ++k; }
    return k;
#else
    return a / b;
#endif
}

// n sequential adds of d onto s, bit-exact.
//
// One loop iteration = one true addition (which also performs every binade, zero or sign crossing) followed by
// a jump over all the steps that stay inside the binade the addition landed in.  Inside a binade with ulp u the
// current value is M*u with an integer M, and the exact sum is (M + t)*u with t = |d| / u, the same for every M —
// so each add moves M by q = nearest integer to t, whatever M is.  t comes from d's own bits (|d| = Md * 2^x, so
// t = Md / 2^sh with sh = the exponent distance), hence no second and third probing addition.  Only the binade in
// which t ends in exactly one half is special: round-to-nearest-even sends an odd M to the even neighbour first and
// from an even M every step is the even one of the two candidates, q = k + (k & 1) with k = floor(t); an odd M there
// just takes one more true addition.  The landing value stays strictly inside the binade (mantissa field >= 1
// downwards: an exact sum just below 2^e is rounded on the finer grid underneath, so the floor itself may only be
// reached by a true addition; <= 0x7FFFFF upwards), which also keeps every skipped intermediate inside it.
S3R_HD float walk_jump(float s, float d, uint32_t n) {
    const uint32_t bd = f2u(d);
    const uint32_t ed = (bd >> 23) & 0xFFu;
    const uint32_t md = (bd & 0x7FFFFFu) | (ed ? 0x800000u : 0u);   // |d| = md * 2^(max(ed, 1) - 150)
    const int32_t edp = (int32_t)(ed ? ed : 1u);
    float cur = s;
    while (n > 0) {
        const float nxt = add_rn(cur, d);
        --n;
        uint32_t bn = f2u(nxt);
        if (bn == f2u(cur)) { return nxt; }       // absorbed: the state repeats, so does every further add
        cur = nxt;
        const uint32_t e = (bn >> 23) & 0xFFu;
        const int32_t sh = (int32_t)(e ? e : 1u) - edp;   // ulp(cur) = 2^sh * ulp(d)
        // sh <= 0: |d| spans the binade, true steps only.  sh >= 25: |d| is below half an ulp, the next addition is
        // absorbed (or leaves the binade through its floor).  e == 255: inf / NaN, the next addition repeats it.
        if (n == 0 || sh <= 0 || sh >= 25 || e == 255u) { continue; }
        const uint32_t k0 = md >> sh, rem = md & ((1u << sh) - 1u), half = 1u << (sh - 1);
        uint32_t q = k0 + (rem > half ? 1u : 0u);
        if (rem == half) {
            if (bn & 1u) { continue; }            // tie binade, odd mantissa: one more true addition settles it
            q = k0 + (k0 & 1u);
        }
        if (q == 0) { continue; }
        const uint32_t mant = bn & 0x7FFFFFu;
        const bool up = ((bn ^ bd) >> 31) == 0u;  // same sign: the magnitude grows
        const uint32_t room = up ? div_small(0x7FFFFFu - mant, q) : (mant ? div_small(mant - 1u, q) : 0u);
        const uint32_t k = room < n ? room : n;
        bn = up ? bn + k * q : bn - k * q;
        n -= k;
        cur = u2f(bn);
    }
    return cur;
}

} // namespace s3r
