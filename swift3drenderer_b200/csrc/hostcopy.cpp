// hostcopy.cpp — see hostcopy.hpp.  Pure host C++ (compiled by g++, no CUDA in here).
#include "hostcopy.hpp"

#include <immintrin.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>

namespace s3r {

__attribute__((target("ssse3"))) static void unpack24_ssse3(uint32_t *dst, const uint8_t *src, size_t pixels) {
    const __m128i shuf = _mm_setr_epi8(0, 1, 2, -128, 3, 4, 5, -128, 6, 7, 8, -128, 9, 10, 11, -128);
    size_t i = 0;
    // non-temporal stores measured slower than regular stores on the B200 host (Xeon, 16 cores): off unless asked for
    static const bool use_nt = getenv("S3R_NT_STORES") && atoi(getenv("S3R_NT_STORES")) != 0;
    if (use_nt && (reinterpret_cast<uintptr_t>(dst) & 15u) == 0) {
        // aligned destination: non-temporal stores (the frame is far larger than the caches and is
        // written once, so skipping the read-for-ownership saves a third of the memory traffic)
        for (; i + 16 + 2 <= pixels; i += 16) {
            const uint8_t *s = src + 3 * i;
            const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i *>(s));
            const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i *>(s + 12));
            const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i *>(s + 24));
            const __m128i d = _mm_loadu_si128(reinterpret_cast<const __m128i *>(s + 36));
            _mm_stream_si128(reinterpret_cast<__m128i *>(dst + i), _mm_shuffle_epi8(a, shuf));
            _mm_stream_si128(reinterpret_cast<__m128i *>(dst + i + 4), _mm_shuffle_epi8(b, shuf));
            _mm_stream_si128(reinterpret_cast<__m128i *>(dst + i + 8), _mm_shuffle_epi8(c, shuf));
            _mm_stream_si128(reinterpret_cast<__m128i *>(dst + i + 12), _mm_shuffle_epi8(d, shuf));
        }
        _mm_sfence();
    }
    // 16 pixels (48 source bytes) per iteration; stop early enough that the 16-byte loads stay inside src
    for (; i + 16 + 2 <= pixels; i += 16) {
        const uint8_t *s = src + 3 * i;
        const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i *>(s));
        const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i *>(s + 12));
        const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i *>(s + 24));
        const __m128i d = _mm_loadu_si128(reinterpret_cast<const __m128i *>(s + 36));
        _mm_storeu_si128(reinterpret_cast<__m128i *>(dst + i), _mm_shuffle_epi8(a, shuf));
        _mm_storeu_si128(reinterpret_cast<__m128i *>(dst + i + 4), _mm_shuffle_epi8(b, shuf));
        _mm_storeu_si128(reinterpret_cast<__m128i *>(dst + i + 8), _mm_shuffle_epi8(c, shuf));
        _mm_storeu_si128(reinterpret_cast<__m128i *>(dst + i + 12), _mm_shuffle_epi8(d, shuf));
    }
    for (; i < pixels; i++) {
        dst[i] = (uint32_t)src[3 * i] | ((uint32_t)src[3 * i + 1] << 8) | ((uint32_t)src[3 * i + 2] << 16);
    }
}

// plain copy with non-temporal stores when both ends are 16-byte aligned (same reasoning as above)
static void copy32(uint8_t *dst, const uint8_t *src, size_t bytes) {
    size_t i = 0;
    static const bool use_nt = getenv("S3R_NT_STORES") && atoi(getenv("S3R_NT_STORES")) != 0;
    if (use_nt && ((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src)) & 15u) == 0) {
        for (; i + 64 <= bytes; i += 64) {
            const __m128i a = _mm_load_si128(reinterpret_cast<const __m128i *>(src + i));
            const __m128i b = _mm_load_si128(reinterpret_cast<const __m128i *>(src + i + 16));
            const __m128i c = _mm_load_si128(reinterpret_cast<const __m128i *>(src + i + 32));
            const __m128i d = _mm_load_si128(reinterpret_cast<const __m128i *>(src + i + 48));
            _mm_stream_si128(reinterpret_cast<__m128i *>(dst + i), a);
            _mm_stream_si128(reinterpret_cast<__m128i *>(dst + i + 16), b);
            _mm_stream_si128(reinterpret_cast<__m128i *>(dst + i + 32), c);
            _mm_stream_si128(reinterpret_cast<__m128i *>(dst + i + 48), d);
        }
        _mm_sfence();
    }
    memcpy(dst + i, src + i, bytes - i);
}

void unpack24(uint32_t *dst, const uint8_t *src, size_t pixels) {
    static const bool has_ssse3 = __builtin_cpu_supports("ssse3");
    if (has_ssse3) { unpack24_ssse3(dst, src, pixels); return; }
    for (size_t i = 0; i < pixels; i++) {
        dst[i] = (uint32_t)src[3 * i] | ((uint32_t)src[3 * i + 1] << 8) | ((uint32_t)src[3 * i + 2] << 16);
    }
}

struct HostCopier::Impl {
    int n;
    std::vector<std::thread> threads;
    std::mutex m;
    std::condition_variable cv;
    uint64_t generation = 0;
    bool stop = false;
    const std::vector<CopySlice> *slices = nullptr;
    bool packed24 = false;
    std::atomic<int> ready{0}, done{0};

    void work(int w) {
        uint64_t seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> g(m);
                cv.wait(g, [&] { return generation != seen; });
                seen = generation;
                if (stop) { return; }
            }
            const std::vector<CopySlice> &sl = *slices;
            const bool packed = packed24;
            for (size_t i = 0; i < sl.size(); i++) {
                while (ready.load(std::memory_order_acquire) <= (int)i) { _mm_pause(); }
                // this worker's 1/n share of the slice, split on 64-pixel boundaries
                const size_t per = (((sl[i].pixels + (size_t)n - 1) / (size_t)n) + 63) & ~(size_t)63;
                const size_t a = std::min(sl[i].pixels, per * (size_t)w), b = std::min(sl[i].pixels, a + per);
                if (b > a) {
                    if (packed) { unpack24(reinterpret_cast<uint32_t *>(sl[i].dst) + a, sl[i].src + 3 * a, b - a); }
                    else { copy32(sl[i].dst + 4 * a, sl[i].src + 4 * a, 4 * (b - a)); }
                }
            }
            done.fetch_add(1, std::memory_order_release);
        }
    }
};

HostCopier::HostCopier(int workers) : impl_(new Impl) {
    impl_->n = std::max(1, workers);
    for (int w = 0; w < impl_->n; w++) { impl_->threads.emplace_back([this, w] { impl_->work(w); }); }
}

HostCopier::~HostCopier() {
    { std::lock_guard<std::mutex> g(impl_->m); impl_->stop = true; impl_->generation++; }
    impl_->cv.notify_all();
    for (auto &t : impl_->threads) { t.join(); }
    delete impl_;
}

int HostCopier::workers() const { return impl_->n; }

void HostCopier::begin(const std::vector<CopySlice> *slices, bool packed24) {
    impl_->slices = slices;
    impl_->packed24 = packed24;
    impl_->ready.store(0, std::memory_order_relaxed);
    impl_->done.store(0, std::memory_order_relaxed);
    { std::lock_guard<std::mutex> g(impl_->m); impl_->generation++; }
    impl_->cv.notify_all();
}

void HostCopier::publish(int n_ready) { impl_->ready.store(n_ready, std::memory_order_release); }

void HostCopier::wait() {
    while (impl_->done.load(std::memory_order_acquire) < impl_->n) { _mm_pause(); }
}

}  // namespace s3r
