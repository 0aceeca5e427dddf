// Soak of the pageable drop-in path: T threads (re-created in waves, so contexts are parked and reused) call
// dbde_pack_frame / dbde_unpack_frame on malloc'd buffers of several geometries for `secs` seconds; EVERY record is
// compared with the first encoding of the same frame and EVERY decoded image with its source.
//   g++ -O2 -std=c++14 -pthread -Iinclude scratch/dropin_soak.cpp -Ldbce-video-cpp_b200 -ldbde_b200 -Wl,-rpath,$PWD/dbce-video-cpp_b200 -o scratch/dropin_soak
//   scratch/dropin_soak T secs
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <atomic>
#include <chrono>
#include <thread>
#include <vector>
#include "dbde_util.h"

static uint64_t sm64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
struct Geo { int W, H; };
int main(int argc, char **argv) {
    const int T = argc > 1 ? atoi(argv[1]) : 8;
    const double secs = argc > 2 ? atof(argv[2]) : 20.0;
    const Geo geos[] = {{2048, 2048}, {1001, 1003}, {2304, 520}, {640, 480}, {2049, 33}};
    std::atomic<long> calls{0}, bad{0};
    const auto t_end = std::chrono::steady_clock::now() + std::chrono::duration<double>(secs);
    int wave = 0;
    while (std::chrono::steady_clock::now() < t_end) {
        std::vector<std::thread> th;
        for (int t = 0; t < T; t++)
            th.emplace_back([&, t, wave] {
                const Geo g = geos[(t + wave) % 5];
                const size_t px = (size_t)g.W * g.H, wh = (size_t)((g.W + 7) / 8) * ((g.H + 7) / 8);
                uint8_t *img = (uint8_t *)malloc(px), *out = (uint8_t *)malloc(px), *rec = (uint8_t *)malloc(32 + 66 * wh), *first = (uint8_t *)malloc(32 + 66 * wh);
                for (size_t i = 0; i < px; i++) { uint64_t h = sm64(i / 8 + 131 * t + 7 * wave); img[i] = (uint8_t)(20 + (h & ((1u << (h >> 60 & 7)) - 1))); }
                size_t n0 = 0;
                const auto w_end = std::chrono::steady_clock::now() + std::chrono::milliseconds(700 + 100 * (t % 4));
                for (int it = 0; std::chrono::steady_clock::now() < w_end; it++) {
                    const size_t n = dbde_pack_frame(42, img, g.W, g.H, rec);
                    if (it == 0) { n0 = n; memcpy(first, rec, n); }
                    else if (n != n0 || memcmp(first, rec, n)) bad++;
                    memset(out, 0xAB, px);
                    uint8_t *p = rec;
                    frame_header fh = dbde_unpack_frame(&p, g.W, g.H, out);
                    if (fh.u64s != 2 || fh.index != 42 || (size_t)(p - rec) != n || memcmp(img, out, px)) bad++;
                    calls++;
                }
                free(img); free(out); free(rec); free(first);
            });
        for (auto &x : th) x.join();
        wave++;
    }
    printf("%s: %ld pack+unpack calls over %d waves of %d threads, %ld mismatches\n", bad.load() ? "FAILED" : "soak ok", calls.load(), wave, T, bad.load());
    return bad.load() ? 1 : 0;
}
