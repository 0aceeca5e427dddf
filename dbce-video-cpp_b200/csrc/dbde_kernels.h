// Internal interface between the C ABI (dbde_capi.cu) and the kernels.
#pragma once
#include "dbde_device.cuh"

namespace dbde {

struct EncParams {
    PartGeom g;
    const uint8_t *frames;      // nframes * W * H, tightly packed, device
    uint8_t *out;               // slot of frame record 0, device; record f lives at out + f * slot_stride
    uint64_t slot_stride;       // bytes between consecutive frame slots (>= 32 + 66*wh)
    uint64_t *frame_offsets;    // nframes entries (= f * slot_stride), device
    uint64_t *frame_sizes;      // nframes entries (record bytes), device
    uint64_t *desc;             // nparts look-back descriptors (index f*ppf + q), zeroed
    unsigned int *ticket;       // zeroed
    uint64_t first_index;
    int nframes;
    unsigned nparts;
    uint32_t flags;             // kFlagInvertRows
};

struct DecParams {
    PartGeom g;
    const uint8_t *stream;      // frame records, device
    uint64_t stream_bytes;
    const uint64_t *frame_offsets;   // nframes entries (record starts), device
    uint8_t *frames;            // nframes * W * H out, device
    uint32_t *status;           // nframes
    uint64_t *indices;          // nframes or null
    uint32_t *wprefix;          // nframes * (ppf*8 + 1) exclusive word prefixes per partition-warp
    int nframes;
    unsigned nparts;
    uint32_t flags;             // kFlagInvertRows
};

// The reference's DBDE_INVERT_ENDIAN build variant (dbde_util.cpp:15-19,24-27,246-270): every 8-pixel tile
// row is byte-reversed before packing and after unpacking.  A launch-uniform flag here.
enum : uint32_t { kFlagInvertRows = 1 };

// status bits reported by the decoder (0 = frame decoded)
enum : uint32_t {
    kStBadFrameHeader = 1,   // u64s != 2                 (dbde_util.cpp:335)
    kStBadDepthCount = 2,    // nb != w*h                 (dbde_util.cpp:296)
    kStBadMinCount = 4,      // nm != w*h                 (dbde_util.cpp:299)
    kStBadWordCount = 8,     // sum(depth) != n64         (dbde_util.cpp:302-303)
    kStDepthTooBig = 16,     // a depth byte > 8 (the reference does not reject this; we do)
    kStTruncated = 32,       // record runs past the end of the stream buffer
};

// ---- DBDE16, the 16-bit extension (dbde16.cu): partitions of 256 consecutive tiles
struct Enc16Params {
    const uint16_t *frames;     // nframes * W * H u16, tightly packed, device
    uint8_t *out;               // record f at out + f * slot_stride
    uint64_t slot_stride;       // >= 32 + 131 * wh
    uint64_t *frame_offsets, *frame_sizes;
    uint64_t *desc;             // nparts look-back descriptors (index f*ppf + q), zeroed
    unsigned int *ticket;       // zeroed
    uint64_t first_index;
    int W, H, w, h, wh, ppf, nframes, aligned;   // aligned: rows can be moved 16 bytes at a time
    unsigned nparts;
};
struct Dec16Params {
    const uint8_t *stream;
    uint64_t stream_bytes;
    const uint64_t *frame_offsets;
    uint16_t *frames;
    uint32_t *status;
    uint64_t *indices;          // or null
    uint32_t *wprefix;          // nframes * (ceil(wh/32) + 1) exclusive word prefixes per 32-tile group
    int W, H, w, h, wh, ppf, nframes, aligned;
    unsigned nparts;
};
cudaError_t launch_encode16(const Enc16Params &P, int num_sms, cudaStream_t stream);
cudaError_t launch_decode16_scan(const Dec16Params &P, cudaStream_t stream);
cudaError_t launch_decode16(const Dec16Params &P, int num_sms, cudaStream_t stream);

// The persistent kernels' launch configuration (dynamic shared memory opt-in + resident CTAs per SM) is a
// property of (device, kernel, shared-memory size): looked up once and remembered, because the two runtime
// calls cost as much as the launch itself when a caller encodes one frame per call.
cudaError_t cached_occupancy(const void *kernel, int threads, size_t smem, int *occ);

size_t enc_smem_bytes(const PartGeom &g);
size_t dec_smem_bytes(const PartGeom &g);
cudaError_t launch_encode(const EncParams &P, bool fast, int num_sms, cudaStream_t stream);
cudaError_t launch_compact(const uint8_t *slots, uint64_t slot_stride, const uint64_t *sizes, int n, uint8_t *dst,
                           cudaStream_t stream);
cudaError_t launch_decode_scan(const DecParams &P, cudaStream_t stream);
cudaError_t launch_decode(const DecParams &P, bool fast, int num_sms, cudaStream_t stream);

}  // namespace dbde
