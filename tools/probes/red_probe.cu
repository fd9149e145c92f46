// tools/probes/red_probe.cu — how fast is fire-and-forget red.global.max.u64 on scattered / short-run addresses?
// (The general path's direct walk publishes one such reduction per covered pixel: DESIGN.md section 5.)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probes/red_probe tools/probes/red_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t mix(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }
__device__ __forceinline__ void red_max(unsigned long long *p, unsigned long long v) { asm volatile("red.global.max.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory"); }

// mode 0: every lane walks a run of `run` consecutive keys starting at a random pixel (like a (triangle, row) item)
// mode 1: the same, but the key is loaded first and the reduction is only issued when it would win
// mode 2: 32-bit reductions
// mode 3: run along x, but only `hit` in 256 of the pixels issue anything (issue-slot baseline)
__global__ void probe(unsigned long long *keys, uint32_t n_px, uint32_t W, uint32_t items, uint32_t run, int mode, uint32_t seed) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < items; i += gridDim.x * blockDim.x) {
        const uint32_t h = mix(i ^ seed);
        // neighbouring lanes = neighbouring rows of the same small box (as in the walk): row r of box b
        const uint32_t box = i >> 3, row = i & 7u, hb = mix(box * 2654435761u ^ seed);
        const uint32_t x0 = hb % (W - 32u), y0 = (hb >> 12) % (n_px / W - 16u);
        unsigned long long *p = keys + (size_t)(y0 + row) * W + x0;
        const unsigned long long v = ((unsigned long long)(h | 1u) << 32) | i;
        for (uint32_t x = 0; x < run; x++) {
            if (mode == 0) { red_max(p + x, v + x); }
            else if (mode == 1) { if (v + x > p[x]) { red_max(p + x, v + x); } }
            else if (mode == 2) { atomicMax(reinterpret_cast<unsigned int *>(p + x), (unsigned int)(v >> 32)); }
            else if (((h >> 8) + x) % 256u < 8u) { red_max(p + x, v + x); }
        }
    }
}

int main() {
    const uint32_t W = 3840, H = 2160, n_px = W * H;
    unsigned long long *keys;
    cudaMalloc(&keys, (size_t)n_px * 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const uint32_t items = 27u << 20, run = 6;   // ~ the C3 frame: 27 M row items, 6.5 pixels each, 1/3 covered -> here every pixel reduces
    for (int mode = 0; mode < 4; mode++) {
        for (int rep = 0; rep < 3; rep++) {
            cudaMemset(keys, mode == 1 ? 0xC0 : 0, (size_t)n_px * 8);   // mode 1: ~3/4 of the candidates lose against the preset keys
            cudaEventRecord(e0);
            probe<<<148 * 8, 256>>>(keys, n_px, W, items, run, mode, 17u + rep);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (rep == 2) { printf("mode %d: %u items x %u px: %.1f us, %.1f G pixel-ops/s\n", mode, items, run, ms * 1e3, (double)items * run / ms / 1e6); }
        }
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
