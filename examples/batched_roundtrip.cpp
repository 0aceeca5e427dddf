// The batched binding of INTEGRATION.md section 2, complete and runnable: encode N frames held in host
// memory into a .dbde file (28-byte video header + frame records, exactly the bytes the reference's
// dbde_pack_video_header + dbde_pack_frame would write), read the file back, index and decode it.
//
//   g++ -O2 -std=c++14 -Iinclude examples/batched_roundtrip.cpp -Ldbce-video-cpp_b200 -ldbde_b200
//       -Wl,-rpath,$PWD/dbce-video-cpp_b200 -o batched_roundtrip && ./batched_roundtrip out.dbde 640 480 32
//
// tests/test_gpu_parity.py builds and runs it on the GPU box and compares the file with the oracle's bytes.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "dbde_util.h"     // the reference's own declarations (video_header, dbde_pack_video_header, ...)
#include "dbde_b200.h"     // the C ABI

// encode N frames held in host memory and append them to a .dbde file
static int write_dbde(FILE *f, const uint8_t *frames, int W, int H, int N, double hz) {
    dbde_b200_ctx *ctx;
    if (dbde_b200_create(0, &ctx)) { fprintf(stderr, "%s\n", dbde_b200_last_error()); return -1; }
    uint8_t hdr[28];
    video_header vh = {3, (uint64_t)H, (uint64_t)W, hz};
    fwrite(hdr, 1, dbde_pack_video_header(vh, hdr), f);               // dbde_util.cpp:198-209
    size_t cap = dbde_b200_stream_bound(W, H, N);
    uint8_t *out;
    std::vector<uint64_t> offs(N + 1);
    dbde_b200_host_alloc(cap, (void **)&out);                          // pinned: fastest D2H
    int rc = dbde_b200_encode_host(ctx, frames, W, H, /*first_index=*/0, N, out, cap, offs.data());
    if (!rc) fwrite(out, 1, offs[N], f);                               // records are back to back
    else fprintf(stderr, "%s\n", dbde_b200_last_error());
    dbde_b200_host_free(out);
    dbde_b200_destroy(ctx);
    return rc;
}

// decode a whole stream (file minus the 28-byte header) held in host memory
static int read_dbde(const uint8_t *stream, size_t bytes, int W, int H, uint8_t *frames, int max_frames) {
    dbde_b200_ctx *ctx;
    if (dbde_b200_create(0, &ctx)) return -1;
    std::vector<uint64_t> offs(max_frames + 1);
    long n = dbde_b200_index_stream(stream, bytes, W, H, offs.data(), max_frames);   // next = cur + 32 + 2wh + 8*n64
    if (n < 0) { dbde_b200_destroy(ctx); return -2; }
    std::vector<uint32_t> status(n ? n : 1);
    int rc = dbde_b200_decode_host(ctx, stream, bytes, offs.data(), W, H, (int)n, frames, status.data(), nullptr);
    // status[i] != 0  <=>  the reference's dbde_unpack_frame would have returned u64s == -1;
    // that frame's pixels are left untouched
    for (long i = 0; i < n && !rc; i++)
        if (status[i]) rc = -3;
    dbde_b200_destroy(ctx);
    return rc ? rc : (int)n;
}

int main(int argc, char **argv) {
    const char *path = argc > 1 ? argv[1] : "roundtrip.dbde";
    int W = argc > 2 ? atoi(argv[2]) : 640, H = argc > 3 ? atoi(argv[3]) : 480, N = argc > 4 ? atoi(argv[4]) : 32;
    const size_t px = (size_t)W * H;
    std::vector<uint8_t> frames(px * N), back(px * N, 0xCD);
    uint64_t z = 42;                                                   // a moving gradient with a little noise
    for (int f = 0; f < N; f++)
        for (int y = 0; y < H; y++)
            for (int x = 0; x < W; x++) {
                z = z * 6364136223846793005ull + 1442695040888963407ull;
                frames[px * f + (size_t)y * W + x] = (uint8_t)(((x + 3 * f) >> 3) + ((y >> 4) & 15) + ((z >> 60) & 3));
            }
    FILE *f = fopen(path, "wb");
    if (!f) return 2;
    if (write_dbde(f, frames.data(), W, H, N, 25.0)) return 3;
    fclose(f);
    f = fopen(path, "rb");
    if (!f) return 4;
    fseek(f, 0, SEEK_END);
    const size_t fbytes = (size_t)ftell(f);
    fseek(f, 0, SEEK_SET);
    std::vector<uint8_t> file(fbytes);
    if (fread(file.data(), 1, fbytes, f) != fbytes) return 5;
    fclose(f);
    uint8_t *hp = file.data();
    video_header vh = dbde_unpack_video_header(&hp);                   // dbde_util.cpp:347-359
    if (vh.u64s != 3 || vh.width != (uint64_t)W || vh.height != (uint64_t)H) return 6;
    int n = read_dbde(file.data() + 28, fbytes - 28, W, H, back.data(), N + 8);
    if (n != N) { fprintf(stderr, "decoded %d of %d frames\n", n, N); return 7; }
    if (memcmp(frames.data(), back.data(), px * N)) { fprintf(stderr, "round trip differs\n"); return 8; }
    printf("%s: %d frames of %dx%d, %zu raw bytes -> %zu file bytes (ratio %.3f), round trip exact\n", path, N, W, H, px * N, fbytes,
           (double)fbytes / (double)(px * N));
    return 0;
}
