// tools/headless.cpp — a headless stand-in for the reference's main loop (main.swift:36-165), in C++.
//
// It binds the plugin exactly the way the Swift shell does — dlopen(path, RTLD_NOW) + dlsym("updateAndRender")
// (main.swift:95-99) — and reproduces its calling pattern: ONE pageable double buffer obtained from realloc(),
// two frames back to back, alternated on every call (main.swift:117-118); resized with realloc() when the
// window size changes (main.swift:156-165); one synchronous updateAndRender per timer tick (main.swift:121);
// optional 60 Hz pacing with the per-second "load %" printout (main.swift:39,142-153).  Works with any library
// that exports the entry point: the reference's own render.so (oracle/_ref) or the B200 one — which is the point:
// the same caller, the same Input records, the same frame checksums.
//
//   g++ -O2 -std=c++17 tools/headless.cpp -o headless -ldl
//   ./headless --lib /path/render.so --inputs inputs.bin --frames 300 --size 1280x720 ...
//     ... [--resize 120:640x360 --resize 200:1280x720] [--pace] [--every 25] [--dump 299:frame.ppm]
//
// inputs.bin: raw 24-byte Input records {float up, down, left, right; float mouse[2]} (render-cpp/render.hpp:15-21),
// e.g. swift3drenderer_b200.scene.input_script(...).tobytes().  Without --inputs a built-in path is used.
// Output: one JSON line — frames, wall seconds, frames/s, and position-weighted checksums of every --every-th frame.
#include <dlfcn.h>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

struct PixelData { uint32_t *buffer; uint32_t width, height, bytesPerPixel, bufferSize; };   // render-cpp/render.hpp:7-13
struct Input { float up, down, left, right; alignas(8) float mouse[2]; };                     // render-cpp/render.hpp:15-21
static_assert(sizeof(PixelData) == 24 && sizeof(Input) == 24, "ABI structs are 24 bytes");
typedef void (*update_fn)(const PixelData *, const Input *);

struct Resize { int frame; uint32_t w, h; };

// position-sensitive frame checksum, easy to restate with numpy: sum of p[i] * (i + 1) modulo 2^64
static uint64_t checksum(const uint32_t *p, size_t n) {
    uint64_t h = 0;
    for (size_t i = 0; i < n; i++) { h += (uint64_t)p[i] * (uint64_t)(i + 1); }
    return h;
}

static bool parse_size(const char *s, uint32_t &w, uint32_t &h) { return sscanf(s, "%ux%u", &w, &h) == 2 && w && h; }

int main(int argc, char **argv) {
    std::string lib, inputs_path, dump_path;
    int frames = 300, every = 25, dump_frame = -1;
    uint32_t W = 960, H = 540;   // main.swift:66: the window opens at 960 x 540
    bool pace = false;
    std::vector<Resize> resizes;
    for (int i = 1; i < argc; i++) {
        const std::string a = argv[i];
        auto next = [&]() -> const char * { if (i + 1 >= argc) { fprintf(stderr, "missing value for %s\n", a.c_str()); exit(2); } return argv[++i]; };
        if (a == "--lib") { lib = next(); }
        else if (a == "--inputs") { inputs_path = next(); }
        else if (a == "--frames") { frames = atoi(next()); }
        else if (a == "--every") { every = atoi(next()); }
        else if (a == "--pace") { pace = true; }
        else if (a == "--size") { if (!parse_size(next(), W, H)) { fprintf(stderr, "bad --size\n"); return 2; } }
        else if (a == "--resize") {
            Resize r; const char *v = next(); const char *colon = strchr(v, ':');
            if (!colon || !parse_size(colon + 1, r.w, r.h)) { fprintf(stderr, "bad --resize (frame:WxH)\n"); return 2; }
            r.frame = atoi(v); resizes.push_back(r);
        } else if (a == "--dump") {
            const char *v = next(); const char *colon = strchr(v, ':');
            if (!colon) { fprintf(stderr, "bad --dump (frame:path)\n"); return 2; }
            dump_frame = atoi(v); dump_path = colon + 1;
        } else { fprintf(stderr, "unknown argument %s\n", a.c_str()); return 2; }
    }
    if (lib.empty()) { fprintf(stderr, "usage: headless --lib render.so [--inputs file] [--frames n] [--size WxH] ...\n"); return 2; }

    void *handle = dlopen(lib.c_str(), RTLD_NOW);                                   // main.swift:96
    if (!handle) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 1; }
    update_fn update = reinterpret_cast<update_fn>(dlsym(handle, "updateAndRender"));   // main.swift:97
    if (!update) { fprintf(stderr, "dlsym(updateAndRender): %s\n", dlerror()); return 1; }

    std::vector<Input> inputs((size_t)frames);
    memset(inputs.data(), 0, inputs.size() * sizeof(Input));
    if (!inputs_path.empty()) {
        FILE *fh = fopen(inputs_path.c_str(), "rb");
        if (!fh) { perror("inputs"); return 1; }
        const size_t got = fread(inputs.data(), sizeof(Input), inputs.size(), fh);
        fclose(fh);
        if (got < inputs.size()) { fprintf(stderr, "inputs: %zu records, %d frames requested\n", got, frames); return 1; }
    } else {   // creep forward, pan slowly
        for (int f = 0; f < frames; f++) { inputs[f].up = 1.f; inputs[f].mouse[0] = 0.5f * f; inputs[f].mouse[1] = 0.2f * f; }
    }

    uint32_t *memory = nullptr;
    PixelData pd{nullptr, 0, 0, 4, 0};
    auto resize = [&](uint32_t w, uint32_t h) {                                       // main.swift:156-165
        pd.width = w; pd.height = h; pd.bufferSize = pd.bytesPerPixel * w * h;
        memory = static_cast<uint32_t *>(realloc(memory, 2 * (size_t)pd.bufferSize));
        if (!memory) { fprintf(stderr, "realloc failed\n"); exit(1); }
    };
    resize(W, H);

    int buffer_index = 0;
    const double frame_target = 1.0 / 60.0;                                            // main.swift:39
    double busy = 0, busy_window = 0; int loops_window = 0;
    std::string sums;
    const auto t_start = std::chrono::steady_clock::now();
    auto window_start = t_start;
    for (int f = 0; f < frames; f++) {
        for (const Resize &r : resizes) { if (r.frame == f) { resize(r.w, r.h); } }
        pd.buffer = memory + (size_t)buffer_index * pd.width * pd.height;             // main.swift:117
        buffer_index = (buffer_index + 1) % 2;                                         // main.swift:118
        const auto mark = std::chrono::steady_clock::now();
        update(&pd, &inputs[f]);                                                       // main.swift:121
        const auto done = std::chrono::steady_clock::now();
        const double dt = std::chrono::duration<double>(done - mark).count();
        busy += dt; busy_window += dt; loops_window++;
        if (every > 0 && (f % every == 0 || f == frames - 1)) {
            char buf[96];
            snprintf(buf, sizeof(buf), "%s[%d, %u, %u, \"%016llx\"]", sums.empty() ? "" : ", ", f, pd.width, pd.height,
                     (unsigned long long)checksum(pd.buffer, (size_t)pd.width * pd.height));
            sums += buf;
        }
        if (f == dump_frame) {   // 0x00RRGGBB -> binary PPM
            FILE *fh = fopen(dump_path.c_str(), "wb");
            if (fh) {
                fprintf(fh, "P6\n%u %u\n255\n", pd.width, pd.height);
                for (size_t i = 0; i < (size_t)pd.width * pd.height; i++) {
                    const uint32_t p = pd.buffer[i];
                    const unsigned char rgb[3] = {(unsigned char)(p >> 16), (unsigned char)(p >> 8), (unsigned char)p};
                    fwrite(rgb, 1, 3, fh);
                }
                fclose(fh);
            }
        }
        if (pace) {
            std::this_thread::sleep_until(t_start + std::chrono::duration_cast<std::chrono::steady_clock::duration>(
                                                        std::chrono::duration<double>((f + 1) * frame_target)));
            if (std::chrono::duration<double>(std::chrono::steady_clock::now() - window_start).count() >= 1.0) {   // main.swift:142-153
                fprintf(stderr, "# loops: %d\n%.2f%%\n", loops_window, 100.0 * busy_window / (frame_target * loops_window));
                busy_window = 0; loops_window = 0; window_start = std::chrono::steady_clock::now();
            }
        }
    }
    const double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count();
    printf("{\"lib\": \"%s\", \"frames\": %d, \"wall_s\": %.6f, \"busy_s\": %.6f, \"fps_busy\": %.3f, \"paced\": %s, \"checksums\": [%s]}\n",
           lib.c_str(), frames, wall, busy, frames / busy, pace ? "true" : "false", sums.c_str());
    free(memory);
    return 0;   // like the reference's shell, nothing is dlclose()d: the plugin owns its state for the process lifetime
}
