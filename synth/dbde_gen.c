/* CPU side of the synthetic frame generators -- TEST / BENCH INFRASTRUCTURE ONLY.
 * See dbde_gen.h for the definitions. */
#include "dbde_gen.h"
#include <stddef.h>

#define GEN_API __attribute__((visibility("default")))

/* Fill `nframes` tightly packed W x H U8 frames, frame numbers f0, f0+1, ... */
GEN_API void gen_frames(int kind, uint64_t seed, uint64_t f0, int nframes, int W, int H, uint8_t *out) {
    for (int i = 0; i < nframes; i++) {
        uint64_t f = f0 + (uint64_t)i;
        uint64_t fkey = gen_frame_key(seed, f);
        gen_blob_t blobs[GEN_NBLOBS];
        if (kind == GEN_MICRO)
            for (int b = 0; b < GEN_NBLOBS; b++) blobs[b] = gen_blob(seed, f, b, W, H);
        uint8_t *dst = out + (size_t)i * W * H;
        for (int y = 0; y < H; y++)
            for (int x = 0; x < W; x++) dst[(size_t)y * W + x] = gen_pixel(kind, seed, f, fkey, y, x, W, H, blobs);
    }
}
