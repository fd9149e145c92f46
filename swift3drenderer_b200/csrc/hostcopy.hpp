// hostcopy.hpp — worker threads that move finished slices of a frame from the library's pinned
// staging buffer into the caller's (pageable) frame buffer while later slices still cross PCIe.
// Default, always-safe way to fill a caller buffer near link speed without registering memory the
// library does not own.  Two modes: plain copy (uint32 pixels) and 24 -> 32 bit expansion: the
// renderer's pixels are 0x00RRGGBB (render-cpp/render.cpp:8), so the zero byte need not cross the link.
#pragma once
#include <stddef.h>
#include <stdint.h>

#include <vector>

namespace s3r {

struct CopySlice {
    const uint8_t *src;   // staging (pinned)
    uint8_t *dst;         // caller memory
    size_t pixels;        // pixels in this slice
};

class HostCopier {
  public:
    explicit HostCopier(int workers);
    ~HostCopier();
    int workers() const;
    // begin(): wake the workers on a slice list (packed24: src holds 3 bytes per pixel);
    // publish(n): slices [0, n) have landed in staging; wait(): every worker has finished every slice.
    void begin(const std::vector<CopySlice> *slices, bool packed24);
    void publish(int n_ready);
    void wait();

  private:
    struct Impl;
    Impl *impl_;
};

// dst[i] = src[3i] | src[3i+1] << 8 | src[3i+2] << 16  (exposed for the unit test)
void unpack24(uint32_t *dst, const uint8_t *src, size_t pixels);

}  // namespace s3r
