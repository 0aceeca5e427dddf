// Feasibility probe: byte-granular TMA through a rank-1 tensor map.  Can 256-byte boxes, stored from (or
// loaded into) 128-byte-aligned shared memory to ARBITRARY byte offsets in global memory, keep up with HBM?
// Each CTA moves "rows" of W bytes (W odd) as ceil(W16/256) boxes, 16 rows per batch, like the odd-size codec would.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scratch/tma_probe scratch/tma_probe.cu   (no -lcuda: entry point looked up at run time)
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s -> %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128) store_probe(const __grid_constant__ CUtensorMap tm256, const __grid_constant__ CUtensorMap tmrest,
                                                   int W, int rows_per_batch, long long nbatches, int rest, int check) {
    extern __shared__ __align__(128) uint8_t sm[];
    const int PA = (W + 127) & ~127;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // fill shared rows with a pattern that encodes (row, col) so the host can verify placement
    for (int i = threadIdx.x; i < rows_per_batch * PA; i += blockDim.x) sm[i] = (uint8_t)((i % PA) * 7 + (i / PA) * 13 + check);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    const int nfull = (W & ~15) / 256;
    for (long long b = blockIdx.x; b < nbatches; b += gridDim.x) {
        if (warp == 0) {
            const long long g0 = b * (long long)rows_per_batch * W;          // byte offset of the batch in the global tensor
            for (int r = lane; r < rows_per_batch * (nfull + 1); r += 32) {
                const int row = r / (nfull + 1), bx = r % (nfull + 1);
                const int x = bx * 256;
                const int c0 = (int)(g0 + (long long)row * W + x);
                if (bx < nfull) {
                    asm volatile("cp.async.bulk.tensor.1d.global.shared::cta.bulk_group [%0, {%1}], [%2];" ::"l"(&tm256), "r"(c0),
                                 "r"(smem_u32(sm + row * PA + x)) : "memory");
                } else if (rest) {
                    asm volatile("cp.async.bulk.tensor.1d.global.shared::cta.bulk_group [%0, {%1}], [%2];" ::"l"(&tmrest), "r"(c0),
                                 "r"(smem_u32(sm + row * PA + x)) : "memory");
                }
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(128) load_probe(const __grid_constant__ CUtensorMap tm256, const __grid_constant__ CUtensorMap tmrest, int W,
                                                  int rows_per_batch, long long nbatches, int rest, unsigned long long *sink) {
    extern __shared__ __align__(128) uint8_t sm[];
    __shared__ __align__(8) uint64_t bar;
    const int PA = (W + 127) & ~127;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int nfull = (W & ~15) / 256;
    unsigned long long acc = 0;
    uint32_t phase = 0;
    for (long long b = blockIdx.x; b < nbatches; b += gridDim.x) {
        if (warp == 0) {
            const long long g0 = b * (long long)rows_per_batch * W;
            if (lane == 0) {
                const uint32_t bytes = (uint32_t)rows_per_batch * (uint32_t)(nfull * 256 + rest);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bytes) : "memory");
            }
            __syncwarp();
            for (int r = lane; r < rows_per_batch * (nfull + 1); r += 32) {
                const int row = r / (nfull + 1), bx = r % (nfull + 1);
                const int x = bx * 256;
                const int c0 = (int)(g0 + (long long)row * W + x);
                if (bx < nfull) {
                    asm volatile("cp.async.bulk.tensor.1d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2}], [%3];" ::"r"(
                                     smem_u32(sm + row * PA + x)), "l"(&tm256), "r"(c0), "r"(smem_u32(&bar)) : "memory");
                } else if (rest) {
                    asm volatile("cp.async.bulk.tensor.1d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2}], [%3];" ::"r"(
                                     smem_u32(sm + row * PA + x)), "l"(&tmrest), "r"(c0), "r"(smem_u32(&bar)) : "memory");
                }
            }
        }
        // everyone waits for the batch
        asm volatile("{\n.reg .pred P1;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra D;\nbra W;\nD:\n}" ::"r"(smem_u32(&bar)), "r"(phase) : "memory");
        phase ^= 1;
        acc += sm[(threadIdx.x * 37) % (rows_per_batch * PA)];
        __syncthreads();
    }
    if (acc == 0x123456789ull) *sink = acc;
}

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                             const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static CUtensorMap make_map(EncodeFn enc, void *base, uint64_t bytes, uint32_t box) {
    CUtensorMap m;
    cuuint64_t dim[1] = {bytes};
    cuuint64_t stride[1] = {0};
    cuuint32_t bx[1] = {box}, es[1] = {1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, base, dim, stride, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled(box %u) failed: %d\n", box, (int)r); exit(2); }
    return m;
}

int main(int argc, char **argv) {
    int W = argc > 1 ? atoi(argv[1]) : 1001, rows = argc > 2 ? atoi(argv[2]) : 16;
    long long nbatches = argc > 3 ? atoll(argv[3]) : 63000;
    int shift = argc > 4 ? atoi(argv[4]) : 5;              // misalignment of the tensor's first byte inside the allocation
    int mode = argc > 5 ? atoi(argv[5]) : 3;               // 1 = stores only, 2 = loads only, 3 = both
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qr;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr));
    EncodeFn enc = (EncodeFn)fn;
    const uint64_t total = (uint64_t)nbatches * rows * W;
    uint8_t *buf;
    CK(cudaMalloc(&buf, total + 4096));
    CK(cudaMemset(buf, 0xEE, total + 4096));
    const int W16 = W & ~15, rest = W16 % 256;
    // the tensor starts at a 16-byte aligned address; the data starts `shift` bytes later (coordinates carry the shift)
    CUtensorMap tm256 = make_map(enc, buf, total + 2048, 256);
    CUtensorMap tmrest = make_map(enc, buf, total + 2048, rest ? rest : 16);
    int dev = 0, sms = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int PA = (W + 127) & ~127;
    const size_t smem = (size_t)rows * PA;
    CK(cudaFuncSetAttribute(store_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(load_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    (void)shift;
    for (int ctas = 2; ctas <= 8; ctas *= 2) {
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        float ms;
        const double moved = (double)nbatches * rows * (W16);
        if (mode & 1) {
        store_probe<<<sms * ctas, 128, smem>>>(tm256, tmrest, W, rows, nbatches, rest, 0);
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0));
        for (int i = 0; i < 5; i++) store_probe<<<sms * ctas, 128, smem>>>(tm256, tmrest, W, rows, nbatches, rest, i);
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= 5;
        printf("store: W=%d rows=%d ctas/SM=%d  %.3f ms  -> %.0f GB/s (aligned part of each row)\n", W, rows, ctas, ms, moved / ms / 1e6);
        }
        if (!(mode & 2)) continue;
        unsigned long long *sink; CK(cudaMalloc(&sink, 8));
        load_probe<<<sms * ctas, 128, smem>>>(tm256, tmrest, W, rows, nbatches, rest, sink);
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0));
        for (int i = 0; i < 5; i++) load_probe<<<sms * ctas, 128, smem>>>(tm256, tmrest, W, rows, nbatches, rest, sink);
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= 5;
        printf("load : W=%d rows=%d ctas/SM=%d  %.3f ms  -> %.0f GB/s\n", W, rows, ctas, ms, moved / ms / 1e6);
    }
    // placement check of the last store pass (check = 4): byte (row, col) of batch b must be ((col)*7 + row*13 + 4) & 255
    std::vector<uint8_t> h(4 * (size_t)rows * W + 64);
    CK(cudaMemcpy(h.data(), buf, h.size(), cudaMemcpyDeviceToHost));
    long bad = 0;
    for (int b = 0; b < 4; b++)
        for (int r = 0; r < rows; r++)
            for (int c = 0; c < W16; c++)
                if (h[((size_t)b * rows + r) * W + c] != (uint8_t)(c * 7 + r * 13 + 4)) bad++;
    printf("placement check: %ld wrong bytes in the first 4 batches (tail bytes of each row are not written by this probe)\n", bad);
    return 0;
}
