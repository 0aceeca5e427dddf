// extern "C" shim around the UNMODIFIED reference library -- TEST / BENCH INFRASTRUCTURE ONLY.
// Compiled by oracle/Makefile together with /root/reference/dbde_util.cpp (read where it
// lies; never copied into this repo) into oracle/_ref/libdbde_ref.so.  The shim only
// forwards to the reference's C++ entry points (dbde_util.h:21-37) and adds a
// std::thread fan-out over contiguous frame ranges for the CPU baseline (the reference
// itself is single-threaded but re-entrant).
#include <stdint.h>
#include <stddef.h>
#include <stdio.h>
#include <string.h>
#include <thread>
#include <vector>
#include <chrono>
#include "dbde_util.h"

#define SHIM_API extern "C" __attribute__((visibility("default")))

SHIM_API uint32_t ref_pack_8x8(uint8_t *image, int stride, uint8_t *target) { return dbde_pack_8x8(image, stride, target); }
SHIM_API uint32_t ref_pack_8x8_partial(uint8_t *image, int stride, int rm, int dm, uint8_t *target) {
    return dbde_pack_8x8_partial(image, stride, rm, dm, target);
}
SHIM_API size_t ref_pack_image(uint8_t *image, int W, int H, uint8_t *target) { return dbde_pack_image(image, W, H, target); }
SHIM_API size_t ref_pack_frame(uint64_t index, uint8_t *image, int W, int H, uint8_t *target) {
    return dbde_pack_frame(index, image, W, H, target);
}
SHIM_API size_t ref_pack_frame_header(uint32_t u64s, uint64_t index, uint64_t elapsed_ns, uint8_t *target) {
    frame_header fh; fh.u64s = u64s; fh.index = index; fh.elapsed_ns = elapsed_ns;
    return dbde_pack_frame_header(fh, target);
}
SHIM_API size_t ref_pack_video_header(uint32_t u64s, uint64_t height, uint64_t width, double hz, uint8_t *target) {
    video_header vh; vh.u64s = u64s; vh.height = height; vh.width = width; vh.frame_hz = hz;
    return dbde_pack_video_header(vh, target);
}
SHIM_API void ref_unpack_8x8(uint8_t depth, uint8_t minval, uint8_t *packed, size_t stride, uint8_t *image) {
    dbde_unpack_8x8(depth, minval, packed, stride, image);
}
SHIM_API size_t ref_unpack_image(uint8_t *packed, int W, int H, uint8_t *image) { return dbde_unpack_image(packed, W, H, image); }
// returns bytes consumed (20 when the image block is rejected); out = {u64s, index, elapsed_ns}
SHIM_API size_t ref_unpack_frame(uint8_t *packed, int W, int H, uint8_t *image, uint64_t out[3]) {
    uint8_t *p = packed;
    frame_header fh = dbde_unpack_frame(&p, W, H, image);
    out[0] = fh.u64s; out[1] = fh.index; out[2] = fh.elapsed_ns;
    return (size_t)(p - packed);
}
SHIM_API size_t ref_unpack_video_header(uint8_t *packed, uint64_t out_u[3], double *hz) {
    uint8_t *p = packed;
    video_header vh = dbde_unpack_video_header(&p);
    out_u[0] = vh.u64s; out_u[1] = vh.height; out_u[2] = vh.width; *hz = vh.frame_hz;
    return (size_t)(p - packed);
}
SHIM_API size_t ref_pack_frames(uint8_t *frames, int W, int H, uint64_t first_index, int n, uint8_t *target, uint64_t *sizes) {
    size_t off = 0;
    for (int i = 0; i < n; i++) {
        size_t s = dbde_pack_frame(first_index + i, frames + (size_t)i * W * H, W, H, target + off);
        if (sizes) sizes[i] = s;
        off += s;
    }
    return off;
}

// ---- CPU baseline: T threads, contiguous frame ranges, fixed-stride output slots ----
// slot_bytes >= 32 + 66*wh.  Returns seconds (steady_clock) for `reps` passes.
SHIM_API double ref_encode_mt(uint8_t *frames, int W, int H, int n, int threads, int reps,
                              uint8_t *slots, size_t slot_bytes, uint64_t *sizes) {
    auto t0 = std::chrono::steady_clock::now();
    for (int rep = 0; rep < reps; rep++) {
        std::vector<std::thread> th;
        for (int t = 0; t < threads; t++)
            th.emplace_back([=]() {
                int a = (int)((long long)n * t / threads), b = (int)((long long)n * (t + 1) / threads);
                for (int i = a; i < b; i++)
                    sizes[i] = dbde_pack_frame((uint64_t)i, frames + (size_t)i * W * H, W, H, slots + (size_t)i * slot_bytes);
            });
        for (auto &x : th) x.join();
    }
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}
SHIM_API double ref_decode_mt(uint8_t *slots, size_t slot_bytes, int W, int H, int n, int threads, int reps,
                              uint8_t *frames, int *bad) {
    auto t0 = std::chrono::steady_clock::now();
    std::vector<int> nbad(threads, 0);
    for (int rep = 0; rep < reps; rep++) {
        std::vector<std::thread> th;
        for (int t = 0; t < threads; t++)
            th.emplace_back([&, t]() {
                int a = (int)((long long)n * t / threads), b = (int)((long long)n * (t + 1) / threads);
                for (int i = a; i < b; i++) {
                    uint8_t *p = slots + (size_t)i * slot_bytes;
                    frame_header fh = dbde_unpack_frame(&p, W, H, frames + (size_t)i * W * H);
                    if (fh.u64s != 2) nbad[t]++;
                }
            });
        for (auto &x : th) x.join();
    }
    double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    int tot = 0; for (int v : nbad) tot += v;
    if (bad) *bad = tot;
    return s;
}
