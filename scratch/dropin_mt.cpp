// The reference's single-frame entry points under a THREADED caller: T host threads, each
// with its own malloc'd frame / record / output buffers, call dbde_pack_frame and then dbde_unpack_frame in a loop
// (the reference's functions are re-entrant on disjoint buffers).  Built twice by scratch/dropin_mt.sh:
// against libdbde_b200.so (the GPU path behind the same symbols) and against the reference object (dbde_util.o).
//   dropin_mt W H reps mode T      mode 0: microscopy-like (depth ~3), 1: noise (depth 8)
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
int main(int argc, char **argv) {
    const int W = argc > 1 ? atoi(argv[1]) : 2048, H = argc > 2 ? atoi(argv[2]) : 2048, reps = argc > 3 ? atoi(argv[3]) : 50;
    const int mode = argc > 4 ? atoi(argv[4]) : 0, T = argc > 5 ? atoi(argv[5]) : 1;
    const size_t px = (size_t)W * H, wh = (size_t)((W + 7) / 8) * ((H + 7) / 8);
    std::vector<uint8_t *> img(T), out(T), rec(T);
    for (int t = 0; t < T; t++) {
        img[t] = (uint8_t *)malloc(px); out[t] = (uint8_t *)malloc(px); rec[t] = (uint8_t *)malloc(32 + 66 * wh);
        for (size_t i = 0; i < px; i++) { uint64_t h = sm64(i + 977 * t); img[t][i] = mode ? (uint8_t)h : (uint8_t)(12 + __builtin_popcountll(h & 0xFF)); }
    }
    std::atomic<int> ready{0}, go{0}, bad{0};
    std::vector<double> te(T), td(T);
    std::vector<std::thread> th;
    auto now = [] { return std::chrono::steady_clock::now(); };
    for (int t = 0; t < T; t++)
        th.emplace_back([&, t] {
            for (int i = 0; i < 3; i++) dbde_pack_frame(i, img[t], W, H, rec[t]);                    // warm-up (GPU build: context creation)
            for (int i = 0; i < 3; i++) { uint8_t *p = rec[t]; dbde_unpack_frame(&p, W, H, out[t]); }
            ready++;
            while (!go.load()) std::this_thread::yield();
            auto t0 = now();
            for (int i = 0; i < reps; i++) dbde_pack_frame(i, img[t], W, H, rec[t]);
            auto t1 = now();
            for (int i = 0; i < reps; i++) { uint8_t *p = rec[t]; frame_header fh = dbde_unpack_frame(&p, W, H, out[t]); if (fh.u64s != 2) bad++; }
            auto t2 = now();
            te[t] = std::chrono::duration<double>(t1 - t0).count();
            td[t] = std::chrono::duration<double>(t2 - t1).count();
            if (memcmp(img[t], out[t], px)) bad++;
        });
    while (ready.load() < T) std::this_thread::yield();
    go = 1;
    for (auto &x : th) x.join();
    double we = 0, wd = 0;
    for (int t = 0; t < T; t++) { if (te[t] > we) we = te[t]; if (td[t] > wd) wd = td[t]; }
    printf("%dx%d mode %d T=%d: dbde_pack_frame %.0f fps aggregate (%.3f ms/call/thread), dbde_unpack_frame %.0f fps aggregate (%.3f ms/call/thread)%s\n",
           W, H, mode, T, T * reps / we, 1e3 * we / reps, T * reps / wd, 1e3 * wd / reps, bad.load() ? "  ROUND TRIP FAILED" : "");
    return bad.load() ? 1 : 0;
}
