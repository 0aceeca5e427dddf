/* DBDE oracle -- TEST INFRASTRUCTURE ONLY.
 *
 * A plain scalar-C restatement of the reference DBDE frame codec
 * (Ichoran/dbce-video-cpp, dbde_util.cpp).  It exists so that tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg can check the CUDA
 * path byte for byte.  Nothing in the product path (dbce-video-cpp_b200/) may
 * include, link or call this file.
 *
 * PARITY PINNED: tests/test_oracle.py checks this file against
 *   - the reference's own KAT stream (dbde_util_test.cpp:135-178),
 *   - the README example (README.md:75-84, words 1-6 and both planes),
 *   - golden vectors produced by the unmodified reference compiled here
 *     (oracle/_ref, see oracle/Makefile; fixtures in tests/golden/),
 *   - differential runs against oracle/_ref on random frames when it is built.
 *
 * Every function cites the reference lines it restates.  The arithmetic is
 * written from the format definition (README.md:50-67), not from the SSE code:
 * a tile's 64 values (pixel - min) are concatenated LSB-first, `depth` bits
 * each, into `depth` little-endian U64 words.
 */
#include <stdint.h>
#include <stddef.h>
#include <string.h>

#define ORACLE_API __attribute__((visibility("default")))

/* The reference's two compile-time format variants, selectable at run time here so one
 * checker serves both builds:
 *   DBDE_INVERT_ENDIAN (dbde_util.cpp:15-19,24-27,246-270): every 8-pixel tile row is byte-reversed
 *     before packing and after unpacking (pixel column c travels as column 7-c);
 *   DBDE_HZ_AS_INTEGER (dbde_util.cpp:203-204,352-353): the video header's frame_hz travels as a
 *     rounded U64 instead of an IEEE-754 double. */
static int g_invert_endian = 0, g_hz_as_integer = 0;
ORACLE_API void oracle_set_variants(int invert_endian, int hz_as_integer) {
    g_invert_endian = invert_endian;
    g_hz_as_integer = hz_as_integer;
}

static void put_le(uint8_t *p, uint64_t v, int nbytes) {
    for (int i = 0; i < nbytes; i++) p[i] = (uint8_t)(v >> (8 * i));
}
static uint64_t get_le(const uint8_t *p, int nbytes) {
    uint64_t v = 0;
    for (int i = 0; i < nbytes; i++) v |= (uint64_t)p[i] << (8 * i);
    return v;
}

/* bits(max-min): 0 for a flat tile, else position of the top set bit + 1.
 * Restates the branch tree at dbde_util.cpp:48,57,66-68. */
static int bit_length_u8(unsigned r) {
    int k = 0;
    while (r) { k++; r >>= 1; }
    return k;
}

/* One full 8x8 tile, rows `stride` apart.  dbde_util.cpp:22-103.
 * Writes exactly 8*depth bytes; returns (depth << 8) | min. */
ORACLE_API uint32_t oracle_pack_8x8(const uint8_t *image, int stride, uint8_t *target) {
    unsigned lo = 255, hi = 0;
    for (int r = 0; r < 8; r++)
        for (int c = 0; c < 8; c++) {
            unsigned p = image[r * stride + c];
            if (p < lo) lo = p;
            if (p > hi) hi = p;
        }
    int k = bit_length_u8(hi - lo);                 /* :48 depth 0 writes nothing */
    if (k == 0) return lo;
    memset(target, 0, (size_t)(8 * k));
    for (int i = 0; i < 64; i++) {
        int c = g_invert_endian ? 7 - (i & 7) : (i & 7);                            /* ENDIAN(), :15-19,24-27 */
        unsigned q = (unsigned)(uint8_t)(image[(i >> 3) * stride + c] - lo);        /* :51-55 */
        int bit = k * i;                            /* field i occupies bits [k*i, k*i+k) */
        for (int b = 0; b < k; b++, bit++)
            if ((q >> b) & 1u) target[bit >> 3] |= (uint8_t)(1u << (bit & 7));
    }
    return ((uint32_t)k << 8) | lo;
}

/* Clamp-to-edge pad a (downmargin x rightmargin) corner to 8x8, then pack.
 * dbde_util.cpp:105-135. */
ORACLE_API uint32_t oracle_pack_8x8_partial(const uint8_t *image, int stride, int rightmargin,
                                            int downmargin, uint8_t *target) {
    uint8_t full[64];
    for (int r = 0; r < 8; r++) {
        int rr = r < downmargin ? r : downmargin - 1;
        for (int c = 0; c < 8; c++) {
            int cc = c < rightmargin ? c : rightmargin - 1;
            full[8 * r + c] = image[rr * stride + cc];
        }
    }
    return oracle_pack_8x8(full, 8, target);
}

/* Whole-frame encode.  dbde_util.cpp:137-180.
 * Layout: I32 wh | U8 depth[wh] | I32 wh | U8 min[wh] | I32 n64 | U64 words[n64]. */
ORACLE_API size_t oracle_pack_image(const uint8_t *image, int W, int H, uint8_t *target) {
    int w = (W + 7) / 8, h = (H + 7) / 8, wh = w * h;
    uint8_t *bd = target + 4, *mi = target + 8 + wh, *words = target + 12 + 2 * (size_t)wh;
    put_le(target, (uint32_t)wh, 4);
    put_le(target + 4 + wh, (uint32_t)wh, 4);
    uint32_t n64 = 0;
    for (int ty = 0; ty < h; ty++)
        for (int tx = 0; tx < w; tx++) {
            int rm = W - 8 * tx; if (rm > 8) rm = 8;
            int dm = H - 8 * ty; if (dm > 8) dm = 8;
            const uint8_t *src = image + (size_t)8 * ty * W + 8 * tx;
            uint32_t r = (rm == 8 && dm == 8) ? oracle_pack_8x8(src, W, words + 8 * (size_t)n64)
                                              : oracle_pack_8x8_partial(src, W, rm, dm, words + 8 * (size_t)n64);
            *bd++ = (uint8_t)(r >> 8);
            *mi++ = (uint8_t)(r & 0xFF);
            n64 += r >> 8;
        }
    put_le(target + 8 + 2 * (size_t)wh, n64, 4);
    return 12 + 2 * (size_t)wh + 8 * (size_t)n64;
}

/* 20-byte frame header: I32 u64s | U64 index | F64 (double)elapsed_ns.
 * dbde_util.cpp:182-188 (elapsed_ns travels as an IEEE-754 double). */
ORACLE_API size_t oracle_pack_frame_header(uint32_t u64s, uint64_t index, uint64_t elapsed_ns, uint8_t *target) {
    double e = (double)elapsed_ns;
    uint64_t ebits;
    memcpy(&ebits, &e, 8);
    put_le(target, u64s, 4);
    put_le(target + 4, index, 8);
    put_le(target + 12, ebits, 8);
    return 20;
}

/* dbde_util.cpp:190-196: header {2, index, 0} then the image block. */
ORACLE_API size_t oracle_pack_frame(uint64_t index, const uint8_t *image, int W, int H, uint8_t *target) {
    size_t sz = oracle_pack_frame_header(2, index, 0, target);
    return sz + oracle_pack_image(image, W, H, target + sz);
}

/* 28-byte video header: I32 u64s | U64 height | U64 width | F64 frame_hz.
 * dbde_util.cpp:198-209 (frame_hz as a rounded U64 under the DBDE_HZ_AS_INTEGER variant). */
ORACLE_API size_t oracle_pack_video_header(uint32_t u64s, uint64_t height, uint64_t width, double frame_hz,
                                           uint8_t *target) {
    uint64_t hz;
    memcpy(&hz, &frame_hz, 8);
    if (g_hz_as_integer) hz = (uint64_t)(long long)(frame_hz + 0.5);            /* :203-204 */
    put_le(target, u64s, 4);
    put_le(target + 4, height, 8);
    put_le(target + 12, width, 8);
    put_le(target + 20, hz, 8);
    return 28;
}

/* One tile decode into 8 rows `stride` apart.  dbde_util.cpp:216-279.
 * depth >= 8 is treated as raw bytes (:229), as the reference does. */
ORACLE_API void oracle_unpack_8x8(uint8_t depth, uint8_t minval, const uint8_t *packed, size_t stride,
                                  uint8_t *image) {
    for (int i = 0; i < 64; i++) {
        unsigned q = 0;
        if (depth >= 8) q = packed[i];
        else {
            int bit = depth * i;
            for (int b = 0; b < depth; b++, bit++) q |= ((packed[bit >> 3] >> (bit & 7)) & 1u) << b;
        }
        int c = g_invert_endian ? 7 - (i & 7) : (i & 7);                      /* ENDIAN(), :246-270 */
        image[(size_t)(i >> 3) * stride + c] = (uint8_t)(q + minval);         /* wrapping add, :246 */
    }
}

/* dbde_util.cpp:281-289: decode, keep only the valid crop. */
ORACLE_API void oracle_unpack_8x8_partial(uint8_t depth, uint8_t minval, const uint8_t *packed, size_t stride,
                                          int rightmargin, int downmargin, uint8_t *image) {
    uint8_t img[64];
    oracle_unpack_8x8(depth, minval, packed, 8, img);
    for (int y = 0; y < downmargin; y++)
        for (int x = 0; x < rightmargin; x++) image[stride * y + x] = img[8 * y + x];
}

/* Whole-frame decode.  dbde_util.cpp:291-328.  Returns bytes consumed, or 0
 * (image untouched) when nb != wh, nm != wh or sum(depth) != n64. */
ORACLE_API size_t oracle_unpack_image(const uint8_t *packed, int W, int H, uint8_t *image) {
    int w = (W + 7) / 8, h = (H + 7) / 8, wh = w * h;
    const uint8_t *pack = packed;
    int32_t nb = (int32_t)get_le(pack, 4); pack += 4;
    if (nb != wh) return 0;
    const uint8_t *b = pack; pack += nb;
    int32_t nm = (int32_t)get_le(pack, 4); pack += 4;
    if (nm != wh) return 0;
    const uint8_t *m = pack; pack += nm;
    int32_t n64 = (int32_t)get_le(pack, 4); pack += 4;
    for (int i = 0; i < wh; i++) n64 -= b[i];
    if (n64 != 0) return 0;
    for (int ty = 0; ty < h; ty++)
        for (int tx = 0; tx < w; tx++) {
            int rm = W - 8 * tx; if (rm > 8) rm = 8;
            int dm = H - 8 * ty; if (dm > 8) dm = 8;
            uint8_t bi = *b++, mn = *m++;
            uint8_t *dst = image + (size_t)8 * ty * W + 8 * tx;
            if (rm == 8 && dm == 8) oracle_unpack_8x8(bi, mn, pack, (size_t)W, dst);
            else oracle_unpack_8x8_partial(bi, mn, pack, (size_t)W, rm, dm, dst);
            pack += 8 * (size_t)bi;
        }
    return (size_t)(pack - packed);
}

/* dbde_util.cpp:330-337.  out[0]=u64s (0xFFFFFFFF when != 2), out[1]=index,
 * out[2]=elapsed_ns (double -> u64).  Always consumes 20 bytes. */
ORACLE_API size_t oracle_unpack_frame_header(const uint8_t *packed, uint64_t out[3]) {
    uint32_t u = (uint32_t)get_le(packed, 4);
    uint64_t ebits = get_le(packed + 12, 8);
    double e;
    memcpy(&e, &ebits, 8);
    out[0] = (u != 2) ? 0xFFFFFFFFu : u;
    out[1] = get_le(packed + 4, 8);
    out[2] = (uint64_t)e;
    return 20;
}

/* dbde_util.cpp:339-345.  Returns bytes consumed INCLUDING the 20-byte header;
 * when the image block is invalid out[0] = 0xFFFFFFFF and only 20 is returned
 * (the reference leaves *packed just after the header). */
ORACLE_API size_t oracle_unpack_frame(const uint8_t *packed, int W, int H, uint8_t *image, uint64_t out[3]) {
    size_t used = oracle_unpack_frame_header(packed, out);
    size_t n = oracle_unpack_image(packed + used, W, H, image);
    if (n == 0) out[0] = 0xFFFFFFFFu;
    else used += n;
    return used;
}

/* dbde_util.cpp:347-359.  out_u[0]=u64s (0xFFFFFFFF when != 3), [1]=height, [2]=width. */
ORACLE_API size_t oracle_unpack_video_header(const uint8_t *packed, uint64_t out_u[3], double *frame_hz) {
    uint32_t u = (uint32_t)get_le(packed, 4);
    uint64_t hz = get_le(packed + 20, 8);
    if (g_hz_as_integer) *frame_hz = (double)hz;                                /* :352-353 */
    else memcpy(frame_hz, &hz, 8);
    out_u[0] = (u != 3) ? 0xFFFFFFFFu : u;
    out_u[1] = get_le(packed + 4, 8);
    out_u[2] = get_le(packed + 12, 8);
    return 28;
}

/* Convenience for tests/bench: encode `n` frames back to back (the byte stream a
 * writer would fwrite after the 28-byte video header, cf. dbde_util_test.cpp:204-211).
 * sizes[i] receives each frame record's size.  Returns total bytes. */
ORACLE_API size_t oracle_pack_frames(const uint8_t *frames, int W, int H, uint64_t first_index, int n,
                                     uint8_t *target, uint64_t *sizes) {
    size_t off = 0;
    for (int i = 0; i < n; i++) {
        size_t s = oracle_pack_frame(first_index + (uint64_t)i, frames + (size_t)i * W * H, W, H, target + off);
        if (sizes) sizes[i] = s;
        off += s;
    }
    return off;
}

/* Decode `n` frame records laid back to back.  Returns bytes consumed, or 0 on the
 * first invalid frame (status of the reference loop in dbde_walk_a_file, :415-416). */
ORACLE_API size_t oracle_unpack_frames(const uint8_t *stream, int W, int H, int n, uint8_t *frames,
                                       uint64_t *indices) {
    size_t off = 0;
    for (int i = 0; i < n; i++) {
        uint64_t hdr[3];
        size_t used = oracle_unpack_frame(stream + off, W, H, frames + (size_t)i * W * H, hdr);
        if (hdr[0] != 2) return 0;
        if (indices) indices[i] = hdr[1];
        off += used;
    }
    return off;
}

/* ======================================================================================
 * DBDE16 -- the 16-bit extension the format note hints at (README.md:65: "could expand
 * size to handle higher bit depth images").  NOT IN THE REFERENCE CODE:
 * there is nothing to be bit-exact with, so this restatement DEFINES the extension and is
 * labelled "parity unpinned".  What pins it anyway (tests/test_oracle.py):
 *   - embedding: a frame whose pixels all fit in 8 bits encodes to the same depth plane, the
 *     same U64 words and the same minima as the reference-pinned 8-bit codec above -- only
 *     the minimum plane is two bytes per tile;
 *   - hand-computed words for small tiles, round trips, clamp padding as at dbde_util.cpp:105-135.
 * Layout of a frame record (20-byte frame header unchanged, dbde_util.cpp:182-188):
 *   I32 wh | U8 depth[wh] (0..16) | I32 2*wh | U16 min[wh] little-endian | I32 n64 | U64 words[n64]
 * A tile's 64 values (pixel - min) are concatenated LSB-first, `depth` bits each, into `depth`
 * little-endian U64 words -- the rule of README.md:54-56 with a wider pixel.  A reader tells
 * the two layouts apart by the length of the minimum plane (wh vs 2*wh).
 * ====================================================================================== */
static int bit_length_u16(unsigned r) {
    int k = 0;
    while (r) { k++; r >>= 1; }
    return k;
}

/* one clamp-padded tile; returns depth, *mn = minimum; writes 8*depth bytes */
static int pack16_tile(const uint16_t *image, int stride, int rm, int dm, uint16_t *mn, uint8_t *target) {
    uint16_t full[64];
    for (int r = 0; r < 8; r++) {
        int rr = r < dm ? r : dm - 1;
        for (int c = 0; c < 8; c++) {
            int cc = c < rm ? c : rm - 1;
            full[8 * r + c] = image[(size_t)rr * stride + cc];
        }
    }
    unsigned lo = 65535, hi = 0;
    for (int i = 0; i < 64; i++) {
        if (full[i] < lo) lo = full[i];
        if (full[i] > hi) hi = full[i];
    }
    int k = bit_length_u16(hi - lo);
    *mn = (uint16_t)lo;
    if (k == 0) return 0;
    memset(target, 0, (size_t)(8 * k));
    for (int i = 0; i < 64; i++) {
        unsigned q = (unsigned)full[i] - lo;
        int bit = k * i;
        for (int b = 0; b < k; b++, bit++)
            if ((q >> b) & 1u) target[bit >> 3] |= (uint8_t)(1u << (bit & 7));
    }
    return k;
}

ORACLE_API size_t oracle_frame_record_bound16(int W, int H) {
    size_t wh = (size_t)((W + 7) / 8) * ((H + 7) / 8);
    return 32 + 3 * wh + 128 * wh;
}

/* image: H x W u16, tightly packed rows.  Returns the record size (20-byte header included). */
ORACLE_API size_t oracle_pack_frame16(uint64_t index, const uint16_t *image, int W, int H, uint8_t *target) {
    int w = (W + 7) / 8, h = (H + 7) / 8, wh = w * h;
    size_t sz = oracle_pack_frame_header(2, index, 0, target);
    uint8_t *t = target + sz;
    uint8_t *bd = t + 4, *mi = t + 8 + wh, *words = t + 12 + 3 * (size_t)wh;
    put_le(t, (uint32_t)wh, 4);
    put_le(t + 4 + wh, (uint32_t)(2 * wh), 4);
    uint32_t n64 = 0;
    for (int ty = 0; ty < h; ty++)
        for (int tx = 0; tx < w; tx++) {
            int rm = W - 8 * tx; if (rm > 8) rm = 8;
            int dm = H - 8 * ty; if (dm > 8) dm = 8;
            uint16_t mn;
            int k = pack16_tile(image + (size_t)8 * ty * W + 8 * tx, W, rm, dm, &mn, words + 8 * (size_t)n64);
            *bd++ = (uint8_t)k;
            put_le(mi, mn, 2); mi += 2;
            n64 += (uint32_t)k;
        }
    put_le(t + 8 + 3 * (size_t)wh, n64, 4);
    return sz + 12 + 3 * (size_t)wh + 8 * (size_t)n64;
}

/* Returns bytes consumed INCLUDING the 20-byte header; on a malformed block (nb != wh, nm != 2*wh,
 * sum(depth) != n64, a depth > 16 -- the checks of dbde_util.cpp:295-303 carried over) out[0] =
 * 0xFFFFFFFF, only 20 is returned and the image is untouched. */
ORACLE_API size_t oracle_unpack_frame16(const uint8_t *packed, int W, int H, uint16_t *image, uint64_t out[3]) {
    int w = (W + 7) / 8, h = (H + 7) / 8, wh = w * h;
    size_t used = oracle_unpack_frame_header(packed, out);
    const uint8_t *pack = packed + used;
    int32_t nb = (int32_t)get_le(pack, 4); pack += 4;
    if (nb != wh) { out[0] = 0xFFFFFFFFu; return used; }
    const uint8_t *b = pack; pack += nb;
    int32_t nm = (int32_t)get_le(pack, 4); pack += 4;
    if (nm != 2 * wh) { out[0] = 0xFFFFFFFFu; return used; }
    const uint8_t *m = pack; pack += nm;
    int64_t n64 = (int64_t)(uint32_t)get_le(pack, 4); pack += 4;
    for (int i = 0; i < wh; i++) {
        if (b[i] > 16) { out[0] = 0xFFFFFFFFu; return used; }
        n64 -= b[i];
    }
    if (n64 != 0) { out[0] = 0xFFFFFFFFu; return used; }
    for (int ty = 0; ty < h; ty++)
        for (int tx = 0; tx < w; tx++) {
            int rm = W - 8 * tx; if (rm > 8) rm = 8;
            int dm = H - 8 * ty; if (dm > 8) dm = 8;
            int k = *b++;
            unsigned mn = (unsigned)get_le(m, 2); m += 2;
            for (int i = 0; i < 64; i++) {
                unsigned q = 0;
                int bit = k * i;
                for (int bb = 0; bb < k; bb++, bit++) q |= ((pack[bit >> 3] >> (bit & 7)) & 1u) << bb;
                int y = i >> 3, x = i & 7;
                if (y < dm && x < rm) image[(size_t)(8 * ty + y) * W + 8 * tx + x] = (uint16_t)(q + mn);
            }
            pack += 8 * (size_t)k;
        }
    return (size_t)(pack - packed);
}

ORACLE_API size_t oracle_pack_frames16(const uint16_t *frames, int W, int H, uint64_t first_index, int n,
                                       uint8_t *target, uint64_t *sizes) {
    size_t off = 0;
    for (int i = 0; i < n; i++) {
        size_t s = oracle_pack_frame16(first_index + (uint64_t)i, frames + (size_t)i * W * H, W, H, target + off);
        if (sizes) sizes[i] = s;
        off += s;
    }
    return off;
}
