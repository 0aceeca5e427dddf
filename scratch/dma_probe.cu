// How fast are small pinned DMAs on this box?  4 MiB moved as k pieces, H2D and D2H, with the host-side cost of
// the enqueue calls and the time until a trailing 4-byte flag copy lands (what the drop-in path waits on).
//   nvcc -O2 -o scratch/dma_probe scratch/dma_probe.cu && scratch/dma_probe
#include <cuda_runtime.h>
#include <stdio.h>
#include <chrono>
#include <vector>
static double now_us() { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main() {
    const size_t n = 4u << 20;
    uint8_t *h, *d; volatile uint32_t *flag; uint32_t *d_one;
    cudaHostAlloc(&h, n, cudaHostAllocDefault); cudaMalloc(&d, n);
    cudaHostAlloc((void **)&flag, 64, cudaHostAllocDefault); cudaMalloc(&d_one, 4);
    uint32_t one = 1; cudaMemcpy(d_one, &one, 4, cudaMemcpyHostToDevice);
    cudaStream_t st; cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
    for (int dir = 0; dir < 2; dir++)
        for (size_t piece : {n, n / 4, n / 16, n / 64}) {
            double best_total = 1e9, best_enq = 1e9;
            for (int rep = 0; rep < 30; rep++) {
                *flag = 0;
                const double t0 = now_us();
                for (size_t o = 0; o < n; o += piece) {
                    if (dir == 0) cudaMemcpyAsync(d + o, h + o, piece, cudaMemcpyHostToDevice, st);
                    else cudaMemcpyAsync(h + o, d + o, piece, cudaMemcpyDeviceToHost, st);
                }
                cudaMemcpyAsync((void *)flag, d_one, 4, cudaMemcpyDeviceToHost, st);
                const double t1 = now_us();
                while (*flag == 0) {}
                const double t2 = now_us();
                if (t2 - t0 < best_total) best_total = t2 - t0;
                if (t1 - t0 < best_enq) best_enq = t1 - t0;
            }
            printf("%s 4 MiB as %3zu pieces of %4zu KiB: enqueue %.1f us, flag seen after %.1f us (%.1f GB/s)\n", dir ? "D2H" : "H2D", n / piece,
                   piece >> 10, best_enq, best_total, n / best_total / 1e3);
        }
    // an empty kernel-less round trip: flag copy alone
    double best = 1e9;
    for (int rep = 0; rep < 50; rep++) {
        *flag = 0;
        const double t0 = now_us();
        cudaMemcpyAsync((void *)flag, d_one, 4, cudaMemcpyDeviceToHost, st);
        while (*flag == 0) {}
        const double t2 = now_us();
        if (t2 - t0 < best) best = t2 - t0;
    }
    printf("flag copy alone: %.1f us\n", best);
    cudaEvent_t ev; cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    best = 1e9;
    for (int rep = 0; rep < 50; rep++) {
        const double t0 = now_us();
        cudaMemcpyAsync(d, h, 4096, cudaMemcpyHostToDevice, st);
        cudaEventRecord(ev, st);
        cudaEventSynchronize(ev);
        const double t2 = now_us();
        if (t2 - t0 < best) best = t2 - t0;
    }
    printf("4 KiB H2D + event sync: %.1f us\n", best);
    return 0;
}
