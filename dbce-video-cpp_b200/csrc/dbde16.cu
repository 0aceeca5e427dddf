// DBDE16 on B200: the 16-bit extension of the DBDE frame codec that the format note hints at
// (reference README.md:65 "could expand size to handle higher bit depth images").
// The reference has no such code, so nothing here replaces a reference function; the layout is the
// reference's with a two-byte minimum plane (restated for the tests in the checker's "DBDE16" section):
//   I32 2 | U64 index | F64 0.0 | I32 wh | U8 depth[wh] (0..16) | I32 2*wh | U16 min[wh] | I32 n64 | U64 words[n64]
// and a frame whose pixels fit in 8 bits gets the 8-bit codec's depth plane and words.
//
// Simpler than the 8-bit kernels (this is the "next" row, not the headline path), same ideas:
// one lane == one 8x8 tile, min/max with VIMNMX.U16x2, depth = 32 - clz(max - min), a partition of 256
// consecutive tiles per CTA iteration, frame-interleaved tickets + the single-word decoupled look-back for
// the word offsets, words staged in shared memory and copied out coalesced.  Pixels are read and written
// with 16-byte accesses when rows allow it (W % 8 == 0, aligned base), element-wise with clamp/crop otherwise.
#include "dbde_device.cuh"
#include "dbde_kernels.h"

namespace dbde {

constexpr int kT16 = 256;                      // tiles per partition == threads per CTA
constexpr int kWordBytes16 = 128 * kT16;       // worst case: 16 words per tile
// The staged words are addressed as 32-bit units with one unit of padding after every 32: a tile's words
// start 2k units after its neighbour's, and without the padding equal-depth neighbours at k = 8 / 16 land
// on 2 / 1 of the 32 banks (measured: all-depth-16 noise 2.0 / 1.1 TB/s before).
__device__ __forceinline__ uint32_t pad16(uint32_t unit) { return unit + (unit >> 5); }
constexpr int kWordBytesPadded16 = kWordBytes16 + kWordBytes16 / 32 + 64;

// ------------------------------------------------------------------ tile <-> registers
// px[4r + j] = pixels (2j, 2j+1) of tile row r, low half first.  Clamp-to-edge padding as
// dbde_pack_8x8_partial (dbde_util.cpp:105-135).
__device__ __forceinline__ void load_tile16(const uint16_t *frame, int W, int H, int ty, int tx, bool aligned, uint32_t (&px)[32]) {
    const int rows_valid = min(8, H - 8 * ty), cols_valid = min(8, W - 8 * tx);
    if (aligned && rows_valid == 8 && cols_valid == 8) {          // the common case: one pointer, eight 16-byte loads
        const uint4 *row = reinterpret_cast<const uint4 *>(frame + (size_t)(8 * ty) * W + 8 * tx);
        const size_t pitch = (size_t)W / 8;                       // in 16-byte units
#pragma unroll
        for (int r = 0; r < 8; r++) {
            const uint4 v = __ldcs(row + r * pitch);
            px[4 * r] = v.x; px[4 * r + 1] = v.y; px[4 * r + 2] = v.z; px[4 * r + 3] = v.w;
        }
        return;
    }
#pragma unroll
    for (int r = 0; r < 8; r++) {
        const uint16_t *row = frame + (size_t)(8 * ty + min(r, rows_valid - 1)) * W + 8 * tx;
        if (aligned && cols_valid == 8) {
            const uint4 v = __ldcs(reinterpret_cast<const uint4 *>(row));
            px[4 * r] = v.x; px[4 * r + 1] = v.y; px[4 * r + 2] = v.z; px[4 * r + 3] = v.w;
        } else if (cols_valid == 8) {
            // a full row of an odd-width frame: 4-byte accesses (the row starts on an even or an odd pixel)
            if (((uintptr_t)row & 3) == 0) {
                const uint32_t *r32 = reinterpret_cast<const uint32_t *>(row);
#pragma unroll
                for (int j = 0; j < 4; j++) px[4 * r + j] = r32[j];
            } else {
                const uint32_t *r32 = reinterpret_cast<const uint32_t *>(row + 1);
                const uint32_t a0 = row[0], w0 = r32[0], w1 = r32[1], w2 = r32[2], a7 = row[7];
                px[4 * r] = a0 | (w0 << 16);
                px[4 * r + 1] = __funnelshift_r(w0, w1, 16);
                px[4 * r + 2] = __funnelshift_r(w1, w2, 16);
                px[4 * r + 3] = (w2 >> 16) | (a7 << 16);
            }
        } else {
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const uint32_t a = row[min(2 * j, cols_valid - 1)], b = row[min(2 * j + 1, cols_valid - 1)];
                px[4 * r + j] = a | (b << 16);
            }
        }
    }
}

// crop as dbde_unpack_8x8_partial (dbde_util.cpp:281-289): only rows < H and columns < W are written
__device__ __forceinline__ void store_tile16(uint16_t *frame, int W, int H, int ty, int tx, bool aligned, const uint32_t (&px)[32]) {
    const int rows_valid = min(8, H - 8 * ty), cols_valid = min(8, W - 8 * tx);
    if (aligned && rows_valid == 8 && cols_valid == 8) {
        uint4 *row = reinterpret_cast<uint4 *>(frame + (size_t)(8 * ty) * W + 8 * tx);
        const size_t pitch = (size_t)W / 8;
#pragma unroll
        for (int r = 0; r < 8; r++) st_stream_v4u32(row + r * pitch, make_uint4(px[4 * r], px[4 * r + 1], px[4 * r + 2], px[4 * r + 3]));
        return;
    }
#pragma unroll
    for (int r = 0; r < 8; r++) {
        if (r >= rows_valid) break;
        uint16_t *row = frame + (size_t)(8 * ty + r) * W + 8 * tx;
        if (aligned && cols_valid == 8) {
            st_stream_v4u32(row, make_uint4(px[4 * r], px[4 * r + 1], px[4 * r + 2], px[4 * r + 3]));
        } else if (cols_valid == 8) {
            if (((uintptr_t)row & 3) == 0) {
                uint32_t *r32 = reinterpret_cast<uint32_t *>(row);
#pragma unroll
                for (int j = 0; j < 4; j++) r32[j] = px[4 * r + j];
            } else {
                uint32_t *r32 = reinterpret_cast<uint32_t *>(row + 1);
                row[0] = (uint16_t)px[4 * r];
                r32[0] = __funnelshift_r(px[4 * r], px[4 * r + 1], 16);
                r32[1] = __funnelshift_r(px[4 * r + 1], px[4 * r + 2], 16);
                r32[2] = __funnelshift_r(px[4 * r + 2], px[4 * r + 3], 16);
                row[7] = (uint16_t)(px[4 * r + 3] >> 16);
            }
        } else {
#pragma unroll
            for (int c = 0; c < 8; c++)
                if (c < cols_valid) row[c] = (uint16_t)(px[4 * r + (c >> 1)] >> (16 * (c & 1)));
        }
    }
}

// ------------------------------------------------------------------ encoder
__global__ void __launch_bounds__(kT16, 4) dbde16_encode_kernel(const Enc16Params P) {
    extern __shared__ __align__(16) uint8_t s_words[];
    __shared__ uint32_t s_ticket, s_wtot[kT16 / 32];
    __shared__ uint64_t s_excl;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const size_t fpix = (size_t)P.W * P.H;
    const size_t fixed = 32 + 3 * (size_t)P.wh;
    for (;;) {
        if (tid == 0) s_ticket = atomicAdd(P.ticket, 1u);
        __syncthreads();
        const unsigned t = s_ticket;
        if (t >= P.nparts) break;
        // frame-interleaved order (ticket t -> frame t mod N, partition t div N): a partition's predecessors
        // in its frame hold smaller tickets, so they have started and the look-back below terminates
        const unsigned q = t / (unsigned)P.nframes, f = t - q * (unsigned)P.nframes;
        const unsigned p = f * (unsigned)P.ppf + q;
        const int tile = (int)q * kT16 + tid;
        const bool valid = tile < P.wh;
        const int ty = valid ? tile / P.w : 0, tx = valid ? tile - ty * P.w : 0;
        uint32_t px[32];
        load_tile16(P.frames + (size_t)f * fpix, P.W, P.H, ty, tx, P.aligned != 0, px);
        // min / max of the 64 pixels, two per word (VIMNMX.U16x2)
        uint32_t lo2 = px[0], hi2 = px[0];
#pragma unroll
        for (int i = 1; i < 32; i++) {
            lo2 = __vminu2(lo2, px[i]);
            hi2 = __vmaxu2(hi2, px[i]);
        }
        uint32_t mn = min(lo2 & 0xffffu, lo2 >> 16), mx = max(hi2 & 0xffffu, hi2 >> 16);
        int k = 32 - __clz((int)(mx - mn));                      // bits(max - min): 0..16
        if (!valid) { k = 0; mn = 0; }
        const uint32_t incl = warp_inclusive_scan((uint32_t)k, lane);
        if (lane == 31) s_wtot[warp] = incl;
        __syncthreads();
        uint32_t wbase = 0, agg = 0;
#pragma unroll
        for (int v = 0; v < kT16 / 32; v++) {
            const uint32_t x = s_wtot[v];
            if (v < warp) wbase += x;
            agg += x;
        }
        if (warp == 0) {
            // decoupled look-back over this frame's partitions (one chain per frame: the word count restarts
            // at every frame, as n64 does at dbde_util.cpp:146)
            uint64_t excl = 0;
            if (q > 0) {
                if (lane == 0) st_relaxed_u64(P.desc + p, desc_make(kDescAggregate, agg));
                excl = lookback_exclusive<true>(P.desc, p, p - q, lane);
            }
            if (lane == 0) {
                st_relaxed_u64(P.desc + p, desc_make(kDescPrefix, excl + agg));
                s_excl = excl;
            }
        }
        // pack (p - min), k bits each, LSB first: two pixels (2k bits) per step into a 64-bit accumulator,
        // flushed to shared memory 32 bits at a time; the tile's k U64 words start at word `off` of the partition
        if (k > 0) {
            const uint32_t m2 = mn * 0x00010001u, kk = 2u * (uint32_t)k;
            const uint32_t sbase = smem_u32(s_words);
            uint32_t unit = 2u * (wbase + incl - (uint32_t)k);                   // 32-bit unit inside the partition's words
            uint32_t acc = 0, fill = 0;                                          // `fill` (< 32) valid low bits in acc
            const uint32_t kmul = 1u << k;
#pragma unroll
            for (int i = 0; i < 32; i++) {
                const uint32_t d = px[i] - m2;                                   // no borrow: every pixel >= min
                const uint32_t pair = (d >> 16) * kmul + (d & 0xffffu);          // 2k bits
                const uint32_t lo = acc | (pair << fill);
                const uint32_t hi = __funnelshift_l(pair, 0u, fill);             // what does not fit in 32 bits (0 when fill == 0)
                fill += kk;
                // a unit is complete: predicated store (no branch), carry the overflow
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ge.u32 p, %2, 32;\n\t@p st.shared.u32 [%0], %1;\n\t}"
                             ::"r"(sbase + 4u * pad16(unit)), "r"(lo), "r"(fill) : "memory");
                const bool full = fill >= 32u;
                acc = full ? hi : lo;
                unit += full ? 1u : 0u;
                fill &= 31u;
            }
        }
        __syncthreads();                                      // words staged, s_excl published
        const uint64_t excl = s_excl;
        uint8_t *rec = P.out + (size_t)f * P.slot_stride;
        if (valid) {
            rec[24 + tile] = (uint8_t)k;                                                  // depth plane
            rec[28 + (size_t)P.wh + 2 * (size_t)tile] = (uint8_t)mn;                      // minimum plane, little-endian U16
            rec[29 + (size_t)P.wh + 2 * (size_t)tile] = (uint8_t)(mn >> 8);
        }
        uint8_t *dst = rec + fixed + 8 * excl;
        const uint32_t *su = reinterpret_cast<const uint32_t *>(s_words);
        if (((uintptr_t)dst & 7) == 0) {
            for (uint32_t i = (uint32_t)tid; i < agg; i += kT16)
                st_stream_u64(dst + 8 * (size_t)i, (uint64_t)su[pad16(2u * i)] | ((uint64_t)su[pad16(2u * i + 1u)] << 32));
        } else {
            for (uint32_t i = (uint32_t)tid; i < 8u * agg; i += kT16) dst[i] = s_words[4u * pad16(i >> 2) + (i & 3u)];
        }
        if (q == (unsigned)P.ppf - 1 && warp == 0) {
            // fixed fields: {I32 2 | U64 index | F64 0.0 | I32 wh} {I32 2*wh} {I32 n64}  (dbde_util.cpp:141-146,182-191)
            const uint32_t n64 = (uint32_t)(excl + agg);
            const uint64_t index = P.first_index + f;
            uint32_t b;
            uint8_t *d8;
            if (lane < 4) { b = (2u >> (8 * lane)) & 0xff; d8 = rec + lane; }
            else if (lane < 12) { b = (uint32_t)(index >> (8 * (lane - 4))) & 0xff; d8 = rec + lane; }
            else if (lane < 20) { b = 0; d8 = rec + lane; }
            else if (lane < 24) { b = ((uint32_t)P.wh >> (8 * (lane - 20))) & 0xff; d8 = rec + lane; }
            else if (lane < 28) { b = ((uint32_t)(2 * P.wh) >> (8 * (lane - 24))) & 0xff; d8 = rec + 24 + P.wh + (lane - 24); }
            else { b = (n64 >> (8 * (lane - 28))) & 0xff; d8 = rec + 28 + 3 * (size_t)P.wh + (lane - 28); }
            *d8 = (uint8_t)b;
            if (lane == 0) {
                P.frame_offsets[f] = (uint64_t)f * P.slot_stride;
                P.frame_sizes[f] = fixed + 8ull * n64;
            }
        }
        __syncthreads();                                      // s_words / s_ticket are reused by the next partition
    }
}

// ------------------------------------------------------------------ decoder: validate + scan
// One CTA per frame: the checks of dbde_unpack_image carried over (dbde_util.cpp:295-303: nb == wh,
// nm == 2*wh, sum(depth) == n64; plus the tag, depth <= 16, bounds) and the exclusive word prefix of every
// 32-tile group (the running input pointer of :312 made explicit).
__global__ void __launch_bounds__(256) dbde16_scan_kernel(const Dec16Params P) {
    __shared__ uint32_t s_warp[8], s_flag;
    const int f = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint64_t off = P.frame_offsets[f];
    const uint64_t fixed = 32 + 3 * (uint64_t)P.wh;
    const int ng = (P.wh + 31) / 32;
    uint32_t *wp = P.wprefix + (size_t)f * (ng + 1);
    if (off + fixed > P.stream_bytes) {
        if (tid == 0) {
            P.status[f] = kStTruncated;
            if (P.indices) P.indices[f] = 0;
        }
        return;
    }
    const uint8_t *rec = P.stream + off;
    auto le32 = [](const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); };
    uint32_t status = 0;
    if (le32(rec) != 2u) status |= kStBadFrameHeader;
    if (le32(rec + 20) != (uint32_t)P.wh) status |= kStBadDepthCount;
    if (le32(rec + 24 + P.wh) != 2u * (uint32_t)P.wh) status |= kStBadMinCount;
    const uint32_t n64 = le32(rec + 28 + 3 * (size_t)P.wh);
    if (tid == 0) s_flag = 0;
    __syncthreads();
    const uint8_t *dp = rec + 24;
    bool big = false;
    for (int g0 = warp * 4; g0 < ng; g0 += 32) {              // four groups in flight per warp
        uint32_t d[4];
#pragma unroll
        for (int b = 0; b < 4; b++) {
            const int t = 32 * (g0 + b) + lane;
            d[b] = (g0 + b < ng && t < P.wh) ? dp[t] : 0u;
        }
#pragma unroll
        for (int b = 0; b < 4; b++) {
            if (d[b] > 16u) { big = true; d[b] = 16u; }
            const uint32_t sum = __reduce_add_sync(0xffffffffu, d[b]);
            if (lane == 0 && g0 + b < ng) wp[g0 + b] = sum;
        }
    }
    if (big) atomicOr(&s_flag, 1u);
    __syncthreads();
    uint32_t carry = 0;
    for (int base = 0; base < ng; base += 2048) {
        uint32_t v[8], local = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int idx = base + tid * 8 + i;
            v[i] = idx < ng ? wp[idx] : 0u;
            local += v[i];
        }
        const uint32_t incl = warp_inclusive_scan(local, lane);
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        uint32_t pre = carry + incl - local, tot = 0;
#pragma unroll
        for (int wv = 0; wv < 8; wv++) {
            const uint32_t x = s_warp[wv];
            if (wv < warp) pre += x;
            tot += x;
        }
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int idx = base + tid * 8 + i;
            if (idx < ng) wp[idx] = pre;
            pre += v[i];
        }
        carry += tot;
        __syncthreads();
    }
    if (tid == 0) {
        wp[ng] = carry;
        if (carry != n64) status |= kStBadWordCount;
        if (s_flag) status |= kStDepthTooBig;
        if (off + fixed + 8ull * carry > P.stream_bytes) status |= kStTruncated;
        P.status[f] = status;
        if (P.indices) P.indices[f] = (uint64_t)le32(rec + 4) | ((uint64_t)le32(rec + 8) << 32);
    }
}

// ------------------------------------------------------------------ decoder: unpack
__global__ void __launch_bounds__(kT16) dbde16_decode_kernel(const Dec16Params P) {
    extern __shared__ __align__(16) uint8_t s_words[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const size_t fpix = (size_t)P.W * P.H;
    const int ng = (P.wh + 31) / 32;
    for (unsigned p = blockIdx.x; p < P.nparts; p += gridDim.x) {
        const unsigned f = p / (unsigned)P.ppf, q = p - f * (unsigned)P.ppf;
        if (P.status[f] != 0) continue;                       // rejected by the scan: the image stays untouched
        const uint8_t *rec = P.stream + P.frame_offsets[f];
        const uint32_t *wp = P.wprefix + (size_t)f * (ng + 1);
        const int g0 = (int)q * (kT16 / 32), g1 = min(g0 + kT16 / 32, ng);
        const uint32_t w0 = wp[g0], agg = wp[g1] - w0;
        // the partition's words, coalesced, into shared memory
        const uint8_t *src = rec + 32 + 3 * (size_t)P.wh + 8ull * w0;
        // Padded layout (pad16): after every 32 units comes one padding unit, and that padding unit holds a
        // COPY of the unit that follows it, so the reader's "this unit and the next" is always two adjacent loads.
        uint32_t *su = reinterpret_cast<uint32_t *>(s_words);
        if (((uintptr_t)src & 7) == 0) {
            for (uint32_t i = (uint32_t)tid; i < agg; i += kT16) {
                const uint64_t v = __ldcs(reinterpret_cast<const uint64_t *>(src) + i);
                const uint32_t at = pad16(2u * i);                                  // units 2i and 2i+1 share a block of 32
                su[at] = (uint32_t)v;
                su[at + 1u] = (uint32_t)(v >> 32);
                if (((2u * i) & 31u) == 0u && i > 0u) su[at - 1u] = (uint32_t)v;    // the copy in the padding slot before it
            }
        } else {
            for (uint32_t i = (uint32_t)tid; i < 8u * agg; i += kT16) {
                const uint32_t u = i >> 2;
                s_words[4u * pad16(u) + (i & 3u)] = src[i];
                if ((u & 31u) == 0u && u > 0u) s_words[4u * (pad16(u) - 1u) + (i & 3u)] = src[i];
            }
        }
        if (tid < 2) {                                                            // the reader looks one unit past the end
            const uint32_t u = 2u * agg + (uint32_t)tid;
            su[pad16(u)] = 0u;
            if ((u & 31u) == 0u && u > 0u) su[pad16(u) - 1u] = 0u;
        }
        const int tile = (int)q * kT16 + tid;
        const bool valid = tile < P.wh;
        int k = 0;
        uint32_t mn = 0;
        if (valid) {
            k = rec[24 + tile];
            const uint8_t *mp = rec + 28 + (size_t)P.wh + 2 * (size_t)tile;
            mn = (uint32_t)mp[0] | ((uint32_t)mp[1] << 8);
        }
        const uint32_t incl = warp_inclusive_scan((uint32_t)k, lane);
        const uint32_t woff = (g0 + warp < ng ? wp[g0 + warp] - w0 : 0u) + incl - (uint32_t)k;
        __syncthreads();
        uint32_t px[32];
        const uint32_t m2 = mn * 0x00010001u;
        if (k > 0) {
            const uint32_t kk = 2u * (uint32_t)k, mk = (1u << k) - 1u;
            const uint32_t *su = reinterpret_cast<const uint32_t *>(s_words);
#pragma unroll
            for (int i = 0; i < 32; i++) {
                const uint32_t bit = kk * (uint32_t)i;                               // pair i: bits [2k i, 2k i + 2k)
                const uint32_t at = pad16(2u * woff + (bit >> 5));
                const uint32_t a = su[at], b = su[at + 1u];                          // the next unit, or its copy in the padding slot
                const uint32_t v = __funnelshift_r(a, b, bit);                       // low 2k bits = the pair
                px[i] = ((v & mk) | (((v >> k) & mk) << 16)) + m2;                    // + min, wrapping per pixel is impossible: <= 65535
            }
        } else {
#pragma unroll
            for (int i = 0; i < 32; i++) px[i] = m2;
        }
        if (valid) {
            const int ty = tile / P.w, tx = tile - ty * P.w;
            store_tile16(P.frames + (size_t)f * fpix, P.W, P.H, ty, tx, P.aligned != 0, px);
        }
        __syncthreads();                                      // s_words is reused by the next partition
    }
}

// ------------------------------------------------------------------ launches
static size_t smem16() { return (size_t)kWordBytesPadded16; }

template <typename Kern, typename Params>
static cudaError_t launch16(Kern kern, const Params &P, unsigned nparts, int num_sms, cudaStream_t stream) {
    int occ = 0;
    cudaError_t e = cached_occupancy((const void *)kern, kT16, smem16(), &occ);
    if (e != cudaSuccess) return e;
    unsigned grid = (unsigned)(num_sms * occ);
    if (grid > nparts) grid = nparts;
    if (grid == 0) return cudaSuccess;
    kern<<<grid, kT16, smem16(), stream>>>(P);
    return cudaGetLastError();
}

cudaError_t launch_encode16(const Enc16Params &P, int num_sms, cudaStream_t stream) {
    return launch16(dbde16_encode_kernel, P, P.nparts, num_sms, stream);
}
cudaError_t launch_decode16_scan(const Dec16Params &P, cudaStream_t stream) {
    if (P.nframes <= 0) return cudaSuccess;
    dbde16_scan_kernel<<<P.nframes, 256, 0, stream>>>(P);
    return cudaGetLastError();
}
cudaError_t launch_decode16(const Dec16Params &P, int num_sms, cudaStream_t stream) {
    return launch16(dbde16_decode_kernel, P, P.nparts, num_sms, stream);
}

}  // namespace dbde
