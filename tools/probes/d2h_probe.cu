// tools/probes/d2h_probe.cu — device-to-host bandwidth of 1..n GPUs writing disjoint parts of ONE host frame at the same time
// (what the in-process multi-GPU drop-in does): cudaHostAlloc'd memory against malloc + cudaHostRegister, whole slices
// against interleaved 61 KB rows (2-D copies).
//   nvcc -O3 -o tools/probes/d2h_probe tools/probes/d2h_probe.cu -lpthread
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#include <sys/mman.h>
#include <cuda_runtime.h>

static double run(int n_gpus, char *host, size_t frame_bytes, bool interleaved, std::vector<char *> &dev, std::vector<cudaStream_t> &st) {
    const int reps = 40;
    const size_t tile = 3840 * 4 * 16, n_tiles = frame_bytes / tile;
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (int g = 0; g < n_gpus; g++) {
        th.emplace_back([&, g] {
            cudaSetDevice(g);
            for (int r = 0; r < reps; r++) {
                if (interleaved) {
                    const size_t owned = (n_tiles - g + n_gpus - 1) / n_gpus;
                    cudaMemcpy2DAsync(host + tile * g, tile * n_gpus, dev[g], tile, tile, owned, cudaMemcpyDeviceToHost, st[g]);
                } else {
                    const size_t part = frame_bytes / n_gpus;
                    cudaMemcpyAsync(host + part * g, dev[g], part, cudaMemcpyDeviceToHost, st[g]);
                }
                cudaStreamSynchronize(st[g]);
            }
        });
    }
    for (auto &t : th) { t.join(); }
    const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return (double)frame_bytes * reps / s / 1e9;
}

// the drop-in's pattern: every frame all GPUs start together (barrier), each sends its tile rows in 6 pieces, frames alternate
// between two host buffers
#include <atomic>
static double run_lockstep(int n_gpus, char *host_a, char *host_b, size_t frame_bytes, std::vector<char *> &dev, std::vector<cudaStream_t> &st) {
    const int reps = 60;
    const size_t tile = 3840 * 4 * 16, n_tiles = frame_bytes / tile;
    std::atomic<int> arrived{0};
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (int g = 0; g < n_gpus; g++) {
        th.emplace_back([&, g] {
            cudaSetDevice(g);
            for (int r = 0; r < reps; r++) {
                arrived.fetch_add(1);
                while (arrived.load() < (r + 1) * n_gpus) { }
                char *host = (r & 1) ? host_b : host_a;
                const size_t owned = (n_tiles - g + n_gpus - 1) / n_gpus;
                const size_t edges[7] = {0, owned / 4, owned / 2, owned * 3 / 4, owned * 7 / 8, owned * 15 / 16, owned};
                for (int b = 0; b < 6; b++) {
                    if (edges[b + 1] > edges[b]) {
                        cudaMemcpy2DAsync(host + tile * g + tile * n_gpus * edges[b], tile * n_gpus, dev[g] + tile * edges[b], tile, tile, edges[b + 1] - edges[b], cudaMemcpyDeviceToHost, st[g]);
                    }
                }
                cudaStreamSynchronize(st[g]);
            }
        });
    }
    for (auto &t : th) { t.join(); }
    const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return (double)frame_bytes * reps / s / 1e9;
}

int main() {
    int count = 0;
    cudaGetDeviceCount(&count);
    const size_t frame_bytes = (size_t)3840 * 2160 * 4;
    std::vector<char *> dev(count);
    std::vector<cudaStream_t> st(count);
    for (int g = 0; g < count; g++) { cudaSetDevice(g); cudaMalloc(&dev[g], frame_bytes); cudaMemset(dev[g], g + 1, frame_bytes); cudaStreamCreate(&st[g]); }
    cudaSetDevice(0);
    char *pinned = nullptr;
    cudaHostAlloc(&pinned, frame_bytes, cudaHostAllocPortable);
    char *paged = (char *)aligned_alloc(4096, frame_bytes);
    memset(paged, 1, frame_bytes);
    cudaHostRegister(paged, frame_bytes, cudaHostRegisterPortable);
    // a large calloc (what numpy.zeros does): pages not touched before registration; and the same with a huge-page hint
    char *fresh = (char *)calloc(frame_bytes + 4096, 1);
    cudaHostRegister(fresh, frame_bytes, cudaHostRegisterPortable);
    char *hinted = (char *)aligned_alloc(2 << 20, (frame_bytes + (2 << 20) - 1) & ~(size_t)((2 << 20) - 1));
    madvise(hinted, frame_bytes, MADV_HUGEPAGE);
    memset(hinted, 1, frame_bytes);
    cudaHostRegister(hinted, frame_bytes, cudaHostRegisterPortable);
    char *odd = (char *)malloc(frame_bytes + 4096) + 48;   // not page aligned, touched
    memset(odd, 1, frame_bytes);
    cudaHostRegister(odd, frame_bytes, cudaHostRegisterPortable);
    printf("%d GPU(s), interleaved: untouched calloc %.1f GB/s, huge-page hint %.1f GB/s, unaligned malloc %.1f GB/s\n", count,
           run(count, fresh, frame_bytes, true, dev, st), run(count, hinted, frame_bytes, true, dev, st), run(count, odd, frame_bytes, true, dev, st));
    {
        char *second = (char *)calloc(frame_bytes + 4096, 1);
        cudaHostRegister(second, frame_bytes, cudaHostRegisterPortable);
        printf("%d GPU(s), lockstep frames, 6 pieces per GPU, two buffers: %.1f GB/s\n", count, run_lockstep(count, fresh, second, frame_bytes, dev, st));
    }
    for (int n = 1; n <= count; n *= 2) {
        for (int inter = 0; inter < 2; inter++) {
            run(n, pinned, frame_bytes, inter, dev, st);
            printf("%d GPU(s), %s: cudaHostAlloc %.1f GB/s, registered malloc %.1f GB/s\n", n, inter ? "interleaved 16-row tiles (2-D copy)" : "contiguous slices",
                   run(n, pinned, frame_bytes, inter, dev, st), run(n, paged, frame_bytes, inter, dev, st));
        }
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
