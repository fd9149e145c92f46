// tests/native/hostcopy_host.cpp — host build of the product's copy engine for the CPU unit test.
#include "../../swift3drenderer_b200/csrc/hostcopy.cpp"

extern "C" {
void t_unpack24(uint32_t *dst, const uint8_t *src, size_t pixels) { s3r::unpack24(dst, src, pixels); }

// Runs the threaded copier over `n_slices` equal slices, publishing them one at a time.
void t_copier(int workers, int packed, const uint8_t *src, uint8_t *dst, size_t pixels, int n_slices, int rounds) {
    s3r::HostCopier c(workers);
    const size_t bpp = packed ? 3 : 4;
    for (int r = 0; r < rounds; r++) {
        std::vector<s3r::CopySlice> sl;
        for (int i = 0; i < n_slices; i++) {
            const size_t a = pixels * i / n_slices, b = pixels * (i + 1) / n_slices;
            sl.push_back(s3r::CopySlice{src + a * bpp, dst + a * 4, b - a});
        }
        c.begin(&sl, packed != 0);
        for (int i = 0; i < n_slices; i++) { c.publish(i + 1); }
        c.wait();
    }
}
}
