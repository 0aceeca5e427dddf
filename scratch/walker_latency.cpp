// The reference's streaming reader used the way its users do: dbde_start_file_walk, then dbde_walk_a_file
// once per frame into a malloc'd image, dbde_end_file_walk.  The file is written here with
// dbde_pack_video_header + dbde_pack_frame (dbde_util_test.cpp:204-211 does the same).  Built against
// libdbde_b200.so and against the reference object by scratch/dropin_latency.sh.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <chrono>
#include "dbde_util.h"

static uint64_t sm64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
int main(int argc, char **argv) {
    int W = argc > 1 ? atoi(argv[1]) : 2048, H = argc > 2 ? atoi(argv[2]) : 2048, N = argc > 3 ? atoi(argv[3]) : 256;
    int buffered = argc > 4 ? atoi(argv[4]) : 16;
    const char *path = argc > 5 ? argv[5] : "/dev/shm/walker_probe.dbde";
    size_t px = (size_t)W * H, wh = (size_t)((W + 7) / 8) * ((H + 7) / 8);
    uint8_t *img = (uint8_t *)malloc(px), *out = (uint8_t *)malloc(px), *rec = (uint8_t *)malloc(32 + 66 * wh);
    for (size_t i = 0; i < px; i++) img[i] = (uint8_t)(12 + __builtin_popcountll(sm64(i) & 0xFF));
    FILE *f = fopen(path, "wb");
    if (!f) return 3;
    uint8_t vhb[28];
    video_header vh = {3, (uint64_t)H, (uint64_t)W, 100.0};
    fwrite(vhb, 1, dbde_pack_video_header(vh, vhb), f);
    dbde_pack_frame(0, img, W, H, rec);                            // warm-up: GPU context creation is not file throughput
    auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < N; i++) {
        img[(size_t)i % px] ^= 0x55;                               // every frame differs a little
        size_t n = dbde_pack_frame(i, img, W, H, rec);
        fwrite(rec, 1, n, f);
    }
    fclose(f);
    auto t1 = std::chrono::steady_clock::now();
    video_header vr;
    dbde_file_walker w = dbde_start_file_walk(path, buffered, &vr);
    if (!w.fptr) return 4;
    frame_header fh;
    int got = 0;
    while (dbde_walk_a_file(&w, &fh, out)) {
        if (fh.index != (uint64_t)got) return 5;
        got++;
    }
    dbde_end_file_walk(&w);
    auto t2 = std::chrono::steady_clock::now();
    if (got != N || memcmp(img, out, px)) { printf("WALK FAILED: %d of %d frames\n", got, N); return 1; }
    double tw = std::chrono::duration<double>(t1 - t0).count(), tr = std::chrono::duration<double>(t2 - t1).count();
    printf("%dx%d, %d frames, %d buffered: write (pack_frame + fwrite) %.0f fps, walk (dbde_walk_a_file) %.0f fps = %.2f GB/s raw\n", W, H, N,
           buffered, N / tw, N / tr, N * (double)px / tr / 1e9);
    remove(path);
    return 0;
}
