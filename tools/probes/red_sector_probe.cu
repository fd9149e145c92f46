// tools/probes/red_sector_probe.cu — does the L2 reduction path work per element or per 32-byte sector?
// Every variant issues the same number of red.global.max.u64 on random small boxes of a 4K key buffer; what changes is
// how many lanes of one instruction fall into the same 32-byte sector (1, 2, 4) or the same 128-byte line (16).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probes/red_sector_probe tools/probes/red_sector_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t mix(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }
__device__ __forceinline__ void red_max(unsigned long long *p, unsigned long long v) { asm volatile("red.global.max.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory"); }

// A warp works on one random box per iteration.  `gx` lanes sit side by side along x (gx = 1, 2, 4, 16), 32 / gx lanes
// along y; every lane then issues `run` reductions, stepping gx pixels along x each time.  `skew` shifts the box start off
// the sector boundary.  Total reductions are the same for every gx.
__global__ void probe(unsigned long long *keys, uint32_t W, uint32_t H, uint32_t boxes, uint32_t run, uint32_t gx, uint32_t skew, uint32_t seed) {
    const uint32_t lane = threadIdx.x & 31u, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    const uint32_t lx = lane % gx, ly = lane / gx;
    for (uint32_t b = warp; b < boxes; b += n_warps) {
        const uint32_t hb = mix(b * 2654435761u ^ seed);
        const uint32_t x0 = ((hb % (W - 256u)) & ~3u) + skew, y0 = (hb >> 12) % (H - 40u);
        unsigned long long *p = keys + (size_t)(y0 + ly) * W + x0 + lx;
        const unsigned long long v = ((unsigned long long)(hb | 1u) << 32) | lane;
        for (uint32_t s = 0; s < run; s++) { red_max(p + s * gx, v + s); }
    }
}

int main() {
    const uint32_t W = 3840, H = 2160;
    unsigned long long *keys;
    cudaMalloc(&keys, (size_t)W * H * 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const uint32_t boxes = 1u << 20, run = 6;
    const uint32_t gxs[] = {1, 2, 4, 4, 8, 16, 32}, skews[] = {0, 0, 0, 2, 0, 0, 0};
    for (int k = 0; k < 7; k++) {
        for (int rep = 0; rep < 3; rep++) {
            cudaMemset(keys, 0, (size_t)W * H * 8);
            cudaEventRecord(e0);
            probe<<<148 * 8, 256>>>(keys, W, H, boxes, run, gxs[k], skews[k], 17u + rep);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (rep == 2) { printf("lanes side by side %2u (skew %u): %u boxes x 32 lanes x %u: %.1f us, %.1f G reductions/s\n", gxs[k], skews[k], boxes, run, ms * 1e3, (double)boxes * 32 * run / ms / 1e6); }
        }
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
