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
// whole number of ulps once one rounding has happened inside that binade (round-to-nearest-even
// settles the tie case after a single step), so the remaining steps inside the binade collapse
// into one integer multiply-add on the bit pattern.  Steps that cross a binade boundary, zero or
// the sign are taken as true additions.  Fixed points (|d| below half an ulp) end the walk early.
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
S3R_HD float walk_jump(float s, float d, uint32_t n) {
    if (n == 0) { return s; }
    float prev = s;
    float cur = add_rn(s, d);
    --n;
    while (n > 0) {
        const float nxt = add_rn(cur, d);
        --n;
        const uint32_t bp = f2u(prev), bc = f2u(cur), bn = f2u(nxt);
        if ((((bp ^ bc) | (bc ^ bn)) >> 23) == 0) { // prev, cur, nxt: same sign, same binade
            const int32_t q = (int32_t)(bn - bc);    // settled increment in ulps (of the magnitude)
            if (q == 0) { return nxt; }              // fixed point: every further add is absorbed
            const uint32_t mant = bn & 0x7FFFFFu;
            // Stay inside the binade.  Downwards the landing mantissa must remain >= 1: an exact sum
            // just below 2^e is rounded on the finer grid of the binade underneath, so the bottom
            // value itself may only be reached by a true addition.
            const uint32_t room = q > 0 ? div_small(0x7FFFFFu - mant, (uint32_t)q)
                                        : (mant ? div_small(mant - 1u, (uint32_t)(-q)) : 0u);
            const uint32_t k = room < n ? room : n;
            const uint32_t landed = bn + k * (uint32_t)q;
            n -= k;
            prev = u2f(landed - (uint32_t)q);
            cur = u2f(landed);
        } else {
            prev = cur;
            cur = nxt;
        }
    }
    return cur;
}

} // namespace s3r
