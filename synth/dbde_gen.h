/* Deterministic integer-only synthetic frame generators.
 * TEST / BENCH INFRASTRUCTURE ONLY -- shared by the CPU generator (dbde_gen.c)
 * and the CUDA generator (synth_gpu.cu) so both emit identical bytes.
 *
 * kinds: 0 noise   v = hpix & 0xFF                      (every tile depth 8; dbde_util_test.cpp:329 style)
 *        1 micro   microscopy-like: dim popcount noise, masked border, 96 drifting blobs
 *        2 mix     tile class (tx+3ty+f)%9 -> depth 0..8 evenly
 *        3 low     70/20/10 % depth 0/1/2
 */
#ifndef DBDE_GEN_H
#define DBDE_GEN_H
#include <stdint.h>

#ifdef __CUDACC__
#define GEN_HD __host__ __device__ __forceinline__
#else
#define GEN_HD static inline
#endif

#define GEN_NOISE 0
#define GEN_MICRO 1
#define GEN_MIX 2
#define GEN_LOW 3
#define GEN_NBLOBS 96
#define GEN_BLOB_R 20

GEN_HD uint64_t gen_sm64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
GEN_HD uint64_t gen_frame_key(uint64_t seed, uint64_t f) { return gen_sm64(seed * 0x100000001B3ull + f); }
GEN_HD uint64_t gen_hpix(uint64_t fkey, uint64_t y, uint64_t x) { return gen_sm64(fkey ^ (y << 24) ^ x); }
GEN_HD uint64_t gen_htile(uint64_t seed, uint64_t f, uint64_t ty, uint64_t tx) {
    return gen_sm64(seed ^ (ty << 32) ^ tx ^ (f << 48));
}
GEN_HD int gen_popc8(uint32_t v) {
    v &= 0xFF;
    v = (v & 0x55) + ((v >> 1) & 0x55);
    v = (v & 0x33) + ((v >> 2) & 0x33);
    return (int)((v & 0x0F) + (v >> 4));
}

typedef struct { int cx, cy, amp; } gen_blob_t;

/* blob b of frame f: centre drifts by 0..3 px per frame in x */
GEN_HD gen_blob_t gen_blob(uint64_t seed, uint64_t f, int b, int W, int H) {
    uint64_t h = gen_sm64(seed * 7919ull + (uint64_t)b);
    gen_blob_t o;
    o.cx = (int)(((h & 0xFFFFF) % (uint64_t)W + f * ((h >> 50) & 3)) % (uint64_t)W);
    o.cy = (int)(((h >> 20) & 0xFFFFF) % (uint64_t)H);
    o.amp = 40 + (int)((h >> 40) % 200);
    return o;
}

GEN_HD uint8_t gen_pixel(int kind, uint64_t seed, uint64_t f, uint64_t fkey, int y, int x, int W, int H,
                         const gen_blob_t *blobs) {
    uint64_t hp = gen_hpix(fkey, (uint64_t)y, (uint64_t)x);
    if (kind == GEN_NOISE) return (uint8_t)(hp & 0xFF);
    if (kind == GEN_MICRO) {
        if (x < W / 16 || y < H / 16) return 0;
        int v = 12 + gen_popc8((uint32_t)hp);
        const int R2 = GEN_BLOB_R * GEN_BLOB_R;
        for (int b = 0; b < GEN_NBLOBS; b++) {
            int dx = x - blobs[b].cx, dy = y - blobs[b].cy;
            int d2 = dx * dx + dy * dy;
            if (d2 < R2) {
                v += blobs[b].amp * (R2 - d2) / R2;
                if (v > 255) v = 255;
            }
        }
        return (uint8_t)v;
    }
    uint64_t ht = gen_htile(seed, f, (uint64_t)(y >> 3), (uint64_t)(x >> 3));
    if (kind == GEN_MIX) {
        int k = (int)(((uint64_t)(x >> 3) + 3ull * (uint64_t)(y >> 3) + f) % 9ull);
        uint32_t range = (1u << k) - 1u;
        uint32_t mn = (uint32_t)(ht % (uint64_t)(256u - range));
        uint32_t v = mn + ((uint32_t)hp & range);
        if ((x & 7) == 0 && (y & 7) == 0) v = mn;
        if ((x & 7) == 1 && (y & 7) == 0) v = mn + range;
        return (uint8_t)v;
    }
    /* GEN_LOW */
    {
        int c = (int)(ht % 10ull);
        int k = c < 7 ? 0 : (c < 9 ? 1 : 2);
        uint32_t mn = (uint32_t)((ht >> 8) % 200ull);
        return (uint8_t)(mn + ((uint32_t)hp & ((1u << k) - 1u)));
    }
}

#endif
