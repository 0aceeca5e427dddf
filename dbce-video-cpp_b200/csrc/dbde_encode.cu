// DBDE B200 encoder: raw U8 frames in HBM -> DBDE frame records laid back to back in HBM,
// byte-identical to what dbde_pack_frame (dbde_util.cpp:190-196) writes frame after frame.
//
// One persistent kernel, one pass over the pixels:
//   warp 8  (producer) : claims partitions from an atomic ticket and bulk-TMAs their pixel rows
//                        into a ring of shared-memory stages (mbarrier full/empty pipeline)
//   warps 0-7 (tiles)  : one lane per 8x8 tile -- min, depth = bits(max-min), bit packing; the
//                        payload words are staged in the (now dead) pixel stage and copied out
//                        with coalesced stores
//   warp 9  (scan)     : publishes the partition's depth sum, resolves its exclusive prefix with
//                        a single-pass decoupled look-back, and hands the output address to the
//                        tile warps; writes the frame's fixed fields (header, lengths, n64)
// Word offsets are per-frame quantities (the reference zeroes n64 per frame, dbde_util.cpp:146),
// so there is one look-back chain per frame, and tickets are INTERLEAVED across the frames of the
// batch (ticket t -> frame t mod N, partition t div N): the partitions in flight at any moment
// belong to different frames, a partition's predecessor was finished a whole generation earlier,
// and the look-back is one descriptor read instead of a convoy of L2 round trips.  Frame f's
// record is written to its own slot (out + f * slot_stride); sizes go to frame_sizes[].
#include "dbde_device.cuh"
#include "dbde_kernels.h"

#include <cstdlib>

namespace dbde {

constexpr int kEncStages = 4;
constexpr int kEncThreads = kTilesPerPart + 64;

struct EncCtl {                 // per-stage control block, written by the producer warp
    int part;                   // partition id, -1 = no more work
    PartInfo pi;                // its geometry (computed once, by the producer)
    uint16_t rowoff[kMaxRowsPerPart];   // byte offset of each row's first pixel inside its smem row
};
struct EncBase {                // per-stage, written by the scan warp
    uint8_t *frame;             // where this frame's record starts
    uint8_t *payload;           // where this partition's first U64 word goes
};

struct EncSmem {
    uint64_t full[kEncStages], empty[kEncStages], aggbar[kEncStages], basebar[kEncStages];
    EncCtl ctl[kEncStages];
    EncBase base[kEncStages];
    uint32_t warptot[kEncStages][kConsumerWarps];   // depth sum of each tile warp
    uint32_t wbase[kEncStages][kConsumerWarps];     // its word offset inside the partition (scan warp)
};

// FAST : 16-byte aligned rows, no partial tiles (W % 16 == 0, H % 8 == 0).
// WST  : (FAST only) every tile warp covers 32 consecutive columns of ONE band, so its own slice
//        of the pixel stage doubles as WARP-PRIVATE payload staging: no CTA barrier anywhere.
template <bool FAST, bool WST>
__global__ void __launch_bounds__(kEncThreads, 3) dbde_encode_kernel(const EncParams P) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    EncSmem &S = *reinterpret_cast<EncSmem *>(smem_raw);
    uint8_t *stages = smem_raw + ((sizeof(EncSmem) + 127) & ~127);
    const PartGeom &g = P.g;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int s = 0; s < kEncStages; s++) {
            mbar_init(&S.full[s], 1);
            mbar_init(&S.empty[s], kConsumerWarps);
            mbar_init(&S.aggbar[s], kConsumerWarps);
            mbar_init(&S.basebar[s], 1);
        }
        mbar_fence_init();
    }
    __syncthreads();

    if (warp == kConsumerWarps) {
        // ============================ producer warp ============================
        const size_t fbytes = (size_t)g.W * g.H;
        for (unsigned it = 0;; it++) {
            const int s = it % kEncStages;
            const uint32_t ph = (it / kEncStages) & 1;
            mbar_wait(&S.empty[s], ph ^ 1);
            unsigned t = 0;
            if (lane == 0) t = atomicAdd(P.ticket, 1u);
            t = __shfl_sync(0xffffffffu, t, 0);
            // frame-interleaved order: all frames' partition 0, then all frames' partition 1, ...
            const unsigned tq = t / (unsigned)P.nframes;
            const unsigned p = (t - tq * (unsigned)P.nframes) * (unsigned)g.ppf + tq;
            if (t >= P.nparts) {
                if (lane == 0) {
                    S.ctl[s].part = -1;
                    mbar_arrive(&S.full[s]);
                }
                break;
            }
            const PartInfo pi = part_info(g, p);
            const uint8_t *fptr = P.frames + (size_t)pi.f * fbytes;
            uint8_t *stage = stages + (size_t)s * g.stage_bytes;
            const int nrows = pi.nbands * 8;
            const int rowbytes = min(8 * pi.ntx, g.W - 8 * pi.tx0);
            // each lane owns rows lane, lane+32
            uint32_t mybytes = 0;
            const uint8_t *src[2];
            uint32_t len[2];
#pragma unroll
            for (int j = 0; j < 2; j++) {
                const int row = lane + 32 * j;
                len[j] = 0;
                src[j] = nullptr;
                if (row < nrows) {
                    const int y = 8 * pi.y0 + row;
                    if (y < g.H) {
                        const uint8_t *a = fptr + (size_t)y * g.W + 8 * pi.tx0;
                        const uintptr_t a0 = (uintptr_t)a & ~(uintptr_t)15;
                        const uintptr_t a1 = ((uintptr_t)a + rowbytes + 15) & ~(uintptr_t)15;
                        src[j] = (const uint8_t *)a0;
                        len[j] = (uint32_t)(a1 - a0);
                        S.ctl[s].rowoff[row] = (uint16_t)((uintptr_t)a - a0);
                    }
                }
                mybytes += len[j];
            }
            const uint32_t total = __reduce_add_sync(0xffffffffu, mybytes);
            __syncwarp();               // every lane's rowoff[] store precedes the release below
            if (lane == 0) {
                S.ctl[s].part = (int)p;
                S.ctl[s].pi = pi;
                mbar_arrive_expect_tx(&S.full[s], total);
            }
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 2; j++)
                if (len[j]) tma_load_1d(stage + (size_t)(lane + 32 * j) * g.pitch, src[j], len[j], &S.full[s]);
        }
    } else if (warp == kConsumerWarps + 1) {
        // ============================ scan warp ============================
        for (unsigned it = 0;; it++) {
            const int s = it % kEncStages;
            const uint32_t ph = (it / kEncStages) & 1;
            mbar_wait(&S.full[s], ph);
            const int part = S.ctl[s].part;
            if (part < 0) break;
            const unsigned p = (unsigned)part;
            const PartInfo pi = S.ctl[s].pi;
            mbar_wait(&S.aggbar[s], ph);
            uint32_t wt = lane < kConsumerWarps ? S.warptot[s][lane] : 0u;
            const uint64_t agg = __reduce_add_sync(0xffffffffu, wt);
            if (WST) {                               // exclusive prefix over the 8 tile warps
                const uint32_t winc = warp_inclusive_scan(wt, lane);
                if (lane < kConsumerWarps) S.wbase[s][lane] = winc - wt;
                __syncwarp();
            }
            uint64_t excl = 0;                      // U64 words of this frame before this partition
            if (pi.q == 0) {
                if (lane == 0) st_relaxed_u64(P.desc + p, desc_make(kDescPrefix, agg));
            } else {
                if (lane == 0) st_relaxed_u64(P.desc + p, desc_make(kDescAggregate, agg));
#ifdef DBDE_EXP_SKIP_LOOKBACK      // profiling-only variant (WRONG output): isolates the pipeline from the scan chain
                excl = (uint64_t)pi.q * 700;
#else
                excl = lookback_exclusive(P.desc, p, p - (unsigned)pi.q, lane);
#endif
                if (lane == 0) st_relaxed_u64(P.desc + p, desc_make(kDescPrefix, excl + agg));
            }
            const size_t fixed = 32 + 2 * (size_t)g.wh;     // frame header + lengths + planes
            uint8_t *frame = P.out + (size_t)pi.f * P.slot_stride;
            if (lane == 0) {
                S.base[s].frame = frame;
                S.base[s].payload = frame + fixed + 8 * excl;
                mbar_arrive(&S.basebar[s]);
            }
            if (pi.q == g.ppf - 1) {
                // last partition of the frame: the fixed fields (dbde_util.cpp:141-146,182-188,191)
                const uint32_t n64 = (uint32_t)(excl + agg);
                const uint64_t index = P.first_index + (uint64_t)pi.f;
                uint32_t b;      // lane i writes one byte of {I32 2 | U64 index | F64 0.0 | I32 wh} {I32 wh} {I32 n64}
                uint8_t *dst;
                if (lane < 4) { b = (2u >> (8 * lane)) & 0xff; dst = frame + lane; }
                else if (lane < 12) { b = (uint32_t)(index >> (8 * (lane - 4))) & 0xff; dst = frame + lane; }
                else if (lane < 20) { b = 0; dst = frame + lane; }
                else if (lane < 24) { b = ((uint32_t)g.wh >> (8 * (lane - 20))) & 0xff; dst = frame + lane; }
                else if (lane < 28) { b = ((uint32_t)g.wh >> (8 * (lane - 24))) & 0xff; dst = frame + 24 + g.wh + (lane - 24); }
                else { b = (n64 >> (8 * (lane - 28))) & 0xff; dst = frame + 28 + 2 * (size_t)g.wh + (lane - 28); }
                *dst = (uint8_t)b;
                if (lane == 0) {
                    P.frame_offsets[pi.f] = (uint64_t)pi.f * P.slot_stride;
                    P.frame_sizes[pi.f] = fixed + 8ull * n64;
                }
            }
        }
    } else {
        // ============================ tile warps: one lane == one 8x8 tile ============================
        // The copy-out of partition i is deferred until partition i+1 has been packed, so the
        // scan warp's look-back for i overlaps the arithmetic of i+1.
        int sb = 0, stx = tid;                  // slot -> (band within partition, tile column)
        if (g.nseg == 1 && g.G > 1) {
            sb = tid / g.w;
            stx = tid - sb * g.w;
        }
        // deferred partition: its depth/min stay in registers until its addresses are known
        int d_s = -1, d_k = 0, d_tfirst = 0;
        uint32_t d_mn = 0, d_ph = 0, d_wtot = 0;
        bool d_valid = false;

        // payload staging address of U64 word `a` (swizzled against bank conflicts).
        //   !WST: a = word index inside the partition, staged linearly from the stage base
        //    WST: a = word index inside this WARP's payload; word (32*row + col) lives where the
        //         warp's pixel row `row`, column `col` was (wb = the warp's first pixel byte)
        auto word_ptr = [&](uint8_t *wb, uint32_t a) -> uint8_t * {
            if (WST) {
                const uint32_t x = swz(a);
                return wb + (size_t)(x >> 5) * g.pitch + 8u * (x & 31u);
            }
            return wb + swz_bytes(8u * a);
        };
        // copy `n` staged words to global memory at `dst`, spread over `nthr` threads (rank `me`)
        auto copy_out = [&](uint8_t *wb, uint8_t *dst, uint32_t n, uint32_t me, uint32_t nthr) {
            const uintptr_t ga = (uintptr_t)dst;
            if ((ga & 7) == 0) {
                // 16-byte stores over the aligned middle, one 8-byte word at either end if needed
                const uint32_t head = (uint32_t)((ga >> 3) & 1);
                if (head && me == 0 && n) st_stream_u64(dst, *reinterpret_cast<const uint64_t *>(word_ptr(wb, 0)));
                const uint32_t npair = n > head ? (n - head) >> 1 : 0u;
                for (uint32_t i = me; i < npair; i += nthr) {
                    const uint32_t a = head + 2 * i;
                    st_stream_v2u64(dst + 8 * (size_t)a, *reinterpret_cast<const uint64_t *>(word_ptr(wb, a)),
                                    *reinterpret_cast<const uint64_t *>(word_ptr(wb, a + 1)));
                }
                const uint32_t done = head + 2 * npair;
                if (done < n && me == 1)
                    st_stream_u64(dst + 8 * (size_t)done, *reinterpret_cast<const uint64_t *>(word_ptr(wb, done)));
            } else if ((ga & 3) == 0) {
                for (uint32_t i = me; i < 2 * n; i += nthr)
                    st_stream_u32(dst + 4 * (size_t)i, *reinterpret_cast<const uint32_t *>(word_ptr(wb, i >> 1) + 4 * (i & 1)));
            } else {
                for (uint32_t i = me; i < 8 * n; i += nthr) dst[i] = word_ptr(wb, i >> 3)[i & 7];
            }
        };

        auto flush_deferred = [&]() {
            uint8_t *stage = stages + (size_t)d_s * g.stage_bytes;
            mbar_wait(&S.basebar[d_s], d_ph);
            uint8_t *frame = S.base[d_s].frame;
            uint8_t *payload = S.base[d_s].payload;
            // ---- depth and minimum planes (dbde_util.cpp:156-157)
            if (d_valid) {
                frame[24 + d_tfirst + tid] = (uint8_t)d_k;
                frame[28 + (size_t)g.wh + d_tfirst + tid] = (uint8_t)d_mn;
            }
            // ---- coalesced copy-out
            if (WST) {      // this warp's words, by this warp
                uint8_t *wb = stage + (size_t)(sb * 8) * g.pitch + 8 * (stx - lane);
                copy_out(wb, payload + 8 * (size_t)S.wbase[d_s][warp], d_wtot, (uint32_t)lane, 32u);
            } else {        // the partition's words, by all tile warps
                uint32_t total = 0;
#pragma unroll
                for (int wv = 0; wv < kConsumerWarps; wv++) total += S.warptot[d_s][wv];
                copy_out(stage, payload, total, (uint32_t)tid, (uint32_t)kTilesPerPart);
            }
            fence_proxy_async();        // my generic accesses to the stage precede the next TMA fill
            __syncwarp();
            if (lane == 0) mbar_arrive(&S.empty[d_s]);
            d_s = -1;
        };

        for (unsigned it = 0;; it++) {
            const int s = it % kEncStages;
            const uint32_t ph = (it / kEncStages) & 1;
            mbar_wait(&S.full[s], ph);
            const int part = S.ctl[s].part;
            if (part < 0) break;
            const PartInfo pi = S.ctl[s].pi;
            uint8_t *stage = stages + (size_t)s * g.stage_bytes;
            const bool valid = tid < pi.nt;
            const int asb = valid ? sb : 0, astx = valid ? stx : 0;   // idle lanes read (and discard) tile 0

            // ---- stage (1)->registers: 8 rows x 8 bytes
            uint32_t px[16];
            if (FAST) {
                const uint8_t *base = stage + (size_t)(asb * 8) * g.pitch + astx * 8;
#pragma unroll
                for (int r = 0; r < 8; r++) {
                    uint2 v = *reinterpret_cast<const uint2 *>(base + (size_t)r * g.pitch);
                    px[2 * r] = v.x;
                    px[2 * r + 1] = v.y;
                }
            } else {
                // clamp-to-edge padding (dbde_util.cpp:105-135): rows past H repeat the last valid
                // row, columns past W repeat the last valid pixel of the row
                const int rows_valid = min(8, g.H - 8 * (pi.y0 + sb));
                const int ncol = min(8, g.W - 8 * (pi.tx0 + stx));
#pragma unroll
                for (int r = 0; r < 8; r++) {
                    uint2 v = make_uint2(0u, 0u);
                    if (valid) {
                        const int row = sb * 8 + min(r, rows_valid - 1);
                        v = lds_u64_unaligned(stage + (size_t)row * g.pitch + S.ctl[s].rowoff[row] + stx * 8);
                        if (ncol < 8) {
                            uint64_t x = ((uint64_t)v.y << 32) | v.x;
                            const uint64_t last = (x >> (8 * (ncol - 1))) & 0xffull;
                            const uint64_t keep = (1ull << (8 * ncol)) - 1ull;
                            x = (x & keep) | ((last * 0x0101010101010101ull) & ~keep);
                            v = make_uint2((uint32_t)x, (uint32_t)(x >> 32));
                        }
                    }
                    px[2 * r] = v.x;
                    px[2 * r + 1] = v.y;
                }
            }
            // ---- stage (2): min, depth
            uint32_t mn = tile_min(px);
            int k = tile_subtract_depth(px, mn);
            if (!valid) { k = 0; mn = 0; }
            // ---- stage (3a): depth sums -> scan warp
            const uint32_t incl = warp_inclusive_scan((uint32_t)k, lane);
            const uint32_t wtot = __shfl_sync(0xffffffffu, incl, 31);
            if (lane == 31) {
                S.warptot[s][warp] = incl;
                mbar_arrive(&S.aggbar[s]);
            }
            uint32_t off = incl - (uint32_t)k;          // WST: word offset inside the warp's payload
            uint8_t *wb;
            if (WST) {
                // the shuffles above already converged the warp: every lane's pixels are in registers,
                // so the warp's own slice of the stage is dead and becomes its payload staging
                wb = stage + (size_t)(sb * 8) * g.pitch + 8 * (stx - lane);
            } else {
                // every tile of this partition is in registers (its stage may be overwritten) and every
                // payload word of the deferred partition has been staged (it may be copied out)
                bar_consumers();
#pragma unroll
                for (int wv = 0; wv < kConsumerWarps; wv++) {
                    const uint32_t t = S.warptot[s][wv];
                    if (wv < warp) off += t;
                }
                wb = stage;
            }
            // ---- stage (4): pack (p - min) into k U64 words, staged (swizzled) in the dead pixel bytes
            if (k > 0) {
                const uint32_t c1 = (1u << k) - 256u, c2 = (1u << (2 * k)) - 65536u;
                uint32_t q[16];
#pragma unroll
                for (int i = 0; i < 16; i++) q[i] = squeeze4(px[i], c1, c2);
                auto store = [&](int n, uint32_t lo, uint32_t hi) {
                    *reinterpret_cast<uint2 *>(word_ptr(wb, off + (uint32_t)n)) = make_uint2(lo, hi);
                };
                switch (k) {
                    case 1: concat_fields<1>(q, store); break;
                    case 2: concat_fields<2>(q, store); break;
                    case 3: concat_fields<3>(q, store); break;
                    case 4: concat_fields<4>(q, store); break;
                    case 5: concat_fields<5>(q, store); break;
                    case 6: concat_fields<6>(q, store); break;
                    case 7: concat_fields<7>(q, store); break;
                    default: concat_fields<8>(q, store); break;
                }
            }
            if (WST) __syncwarp();      // the warp's payload is staged before any lane copies it out later
            if (d_s >= 0) flush_deferred();
            d_s = s; d_ph = ph; d_k = k; d_mn = mn; d_tfirst = pi.tfirst; d_valid = valid; d_wtot = wtot;
        }
        if (d_s >= 0) {
            if (!WST) bar_consumers();  // the last partition's payload is fully staged
            flush_deferred();
        }
    }
}

size_t enc_smem_bytes(const PartGeom &g) {
    return ((sizeof(EncSmem) + 127) & ~(size_t)127) + (size_t)kEncStages * g.stage_bytes;
}

cudaError_t launch_encode(const EncParams &P, bool fast, int num_sms, cudaStream_t stream) {
    const size_t smem = enc_smem_bytes(P.g);
    // warp-private staging needs whole warps inside one band: every partition's ntx % 32 == 0
    // Measured on B200 (micro-2048): CTA-wide staging 1.18 ms vs warp-private 1.24 ms per 1000 frames
    // -- the barrier it removes is hidden by the other resident CTAs while its address math is not --
    // so it stays an opt-in experiment (DBDE_B200_WST=1).
    static const bool want_wst = getenv("DBDE_B200_WST") && atoi(getenv("DBDE_B200_WST")) != 0;
    const bool wst = want_wst && fast && (P.g.w % 32 == 0);
    auto kern = !fast ? dbde_encode_kernel<false, false> : (wst ? dbde_encode_kernel<true, true> : dbde_encode_kernel<true, false>);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int occ = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kEncThreads, smem);
    if (e != cudaSuccess) return e;
    if (occ < 1) return cudaErrorLaunchOutOfResources;
    unsigned grid = (unsigned)(num_sms * occ);
    if (grid > P.nparts) grid = P.nparts;
    if (grid == 0) return cudaSuccess;
    kern<<<grid, kEncThreads, smem, stream>>>(P);
    return cudaGetLastError();
}

}  // namespace dbde
