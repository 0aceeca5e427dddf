// Latency of the reference's single-frame entry points (dbde_pack_frame / dbde_unpack_frame) called the
// way an unmodified program calls them: malloc'd buffers, one frame per call, synchronous.
// Built twice by scratch/dropin_latency.sh: against libdbde_b200.so and against the reference object.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <chrono>
#include "dbde_util.h"
#ifdef WITH_B200
#include "dbde_b200.h"
#endif

static uint64_t sm64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
int main(int argc, char **argv) {
    int W = argc > 1 ? atoi(argv[1]) : 2048, H = argc > 2 ? atoi(argv[2]) : 2048, reps = argc > 3 ? atoi(argv[3]) : 50;
    int mode = argc > 4 ? atoi(argv[4]) : 0;      // 0: microscopy-like (depth ~3), 1: noise (depth 8)
    size_t px = (size_t)W * H, wh = (size_t)((W + 7) / 8) * ((H + 7) / 8);
    uint8_t *img = (uint8_t *)malloc(px), *out = (uint8_t *)malloc(px), *rec = (uint8_t *)malloc(32 + 66 * wh);
    for (size_t i = 0; i < px; i++) { uint64_t h = sm64(i); img[i] = mode ? (uint8_t)h : (uint8_t)(12 + __builtin_popcountll(h & 0xFF)); }
    size_t n = 0;
#ifdef WITH_B200
    if (getenv("REGISTER")) {       // what a maintainer adds to get DMA copies: two calls per long-lived buffer
        dbde_b200_host_register(img, px); dbde_b200_host_register(out, px); dbde_b200_host_register(rec, 32 + 66 * wh);
    }
#endif
    for (int i = 0; i < 3; i++) n = dbde_pack_frame(i, img, W, H, rec);          // warm-up (context creation on the GPU build)
    for (int i = 0; i < 3; i++) { uint8_t *p = rec; dbde_unpack_frame(&p, W, H, out); }   // warm-up of the decode side
    auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < reps; i++) n = dbde_pack_frame(i, img, W, H, rec);
    auto t1 = std::chrono::steady_clock::now();
    for (int i = 0; i < reps; i++) { uint8_t *p = rec; frame_header fh = dbde_unpack_frame(&p, W, H, out); if (fh.u64s != 2) return 2; }
    auto t2 = std::chrono::steady_clock::now();
    if (memcmp(img, out, px)) { printf("ROUND TRIP FAILED\n"); return 1; }
    double e = std::chrono::duration<double, std::milli>(t1 - t0).count() / reps, d = std::chrono::duration<double, std::milli>(t2 - t1).count() / reps;
    printf("%dx%d mode %d record %zu B: dbde_pack_frame %.3f ms/call (%.0f fps), dbde_unpack_frame %.3f ms/call (%.0f fps)\n", W, H, mode, n, e,
           1e3 / e, d, 1e3 / d);
    return 0;
}
