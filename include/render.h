/*
 * include/render.h — the drop-in C ABI of the renderer plugin.
 *
 * Replaces, symbol for symbol, the reference's render-cpp/render.hpp:7-21 (the two boundary
 * structs) and the single exported entry point render-cpp/render.cpp:264-265
 *     __attribute__((visibility("default"))) void updateAndRender(const PixelData*, const Input*);
 * which the reference's main loop binds with dlopen + dlsym("updateAndRender") (main.swift:95-99)
 * and calls once per frame (main.swift:121).
 *
 * Contract kept: synchronous — the caller's pixel buffer (pageable host memory,
 * W*H uint32 0x00RRGGBB, row 0 on top, pitch = width) is complete on return; any call may carry
 * a new width/height; camera state lives inside the library and advances by one Input per call;
 * data.bin is found next to the shared object (then Resources/data.bin, then
 * ../data-generator/data.bin) and the process exits with status 666 & 0xFF when it is missing
 * (render-cpp/render.cpp:161-176).
 */
#ifndef S3R_RENDER_H
#define S3R_RENDER_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {          /* render-cpp/render.hpp:7-13, 24 bytes */
    uint32_t *buffer;
    uint32_t width;
    uint32_t height;
    uint32_t bytesPerPixel;
    uint32_t bufferSize;  /* bytes = 4 * width * height (main.swift:163) */
} PixelData;

typedef struct {          /* render-cpp/render.hpp:15-21, 24 bytes, mouse (simd_float2) 8-aligned @16 */
    float up;
    float down;
    float left;
    float right;
    float mouse[2];       /* absolute accumulated position (input.swift:75-91) */
} Input;

__attribute__((visibility("default"))) void updateAndRender(const PixelData *pixel_data, const Input *input);

#ifdef __cplusplus
}
#endif
#endif
