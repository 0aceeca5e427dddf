// CUDA synthetic-frame generator -- TEST / BENCH INFRASTRUCTURE ONLY (not part of the product
// library).  Emits exactly the bytes of gen_frames() in dbde_gen.c (same integer arithmetic,
// shared through dbde_gen.h) so bench.py can fill HBM with BASELINE.json's synthetic video
// without a minutes-long CPU loop.  tests/test_gpu_parity.py checks it against the CPU generator.
#include "dbde_gen.h"
#include <cuda_runtime.h>

__global__ void synth_kernel(int kind, uint64_t seed, uint64_t f0, int nframes, int W, int H, uint8_t *out) {
    __shared__ gen_blob_t blobs[GEN_NBLOBS];
    const int fi = blockIdx.y;
    const uint64_t f = f0 + (uint64_t)fi;
    if (kind == GEN_MICRO) {
        for (int b = threadIdx.x; b < GEN_NBLOBS; b += blockDim.x) blobs[b] = gen_blob(seed, f, b, W, H);
        __syncthreads();
    }
    const uint64_t fkey = gen_frame_key(seed, f);
    const size_t npix = (size_t)W * H;
    uint8_t *dst = out + (size_t)fi * npix;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += (size_t)gridDim.x * blockDim.x) {
        int y = (int)(i / (size_t)W), x = (int)(i % (size_t)W);
        dst[i] = gen_pixel(kind, seed, f, fkey, y, x, W, H, blobs);
    }
}

// `out` is a device pointer; `stream` a cudaStream_t (0 = default).  Returns a cudaError_t.
extern "C" __attribute__((visibility("default")))
int synth_frames_device(int kind, uint64_t seed, uint64_t f0, int nframes, int W, int H, uint8_t *out, void *stream) {
    if (nframes <= 0) return 0;
    dim3 grid(256, nframes);
    synth_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(kind, seed, f0, nframes, W, H, out);
    return (int)cudaGetLastError();
}
