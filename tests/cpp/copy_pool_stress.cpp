// CPU stress of the copy pool (no CUDA): T caller threads relay buffers through "bounce" memory exactly the
// way relay_h2d / relay_d2h do -- jobs with per-piece counters on the caller's stack, helping waits on the
// jobs and on flags that another thread (standing in for the GPU's flag copies) sets later.
//   copy_pool_stress T reps        (DBDE_B200_COPY_THREADS picks the pool size; 0 = callers only)
#include <stdio.h>

#include "../../dbce-video-cpp_b200/csrc/dbde_copy_pool.h"

using namespace dbde;

int main(int argc, char **argv) {
    const int T = argc > 1 ? atoi(argv[1]) : 8, reps = argc > 2 ? atoi(argv[2]) : 200;
    const size_t n = (size_t)3 << 20, job = 256u << 10;
    std::atomic<int> bad{0};
    std::vector<std::thread> th;
    // the "device": completes flags a little later, in order
    struct Flag { volatile uint32_t v; };
    std::vector<std::vector<Flag>> flags(T, std::vector<Flag>(16));
    std::atomic<bool> stop{false};
    std::mutex fm;
    std::deque<volatile uint32_t *> fq;
    std::thread device([&] {
        while (!stop.load()) {
            volatile uint32_t *f = nullptr;
            {
                std::lock_guard<std::mutex> lk(fm);
                if (!fq.empty()) { f = fq.front(); fq.pop_front(); }
            }
            if (f) *f = 1u; else std::this_thread::yield();
        }
    });
    for (int t = 0; t < T; t++)
        th.emplace_back([&, t] {
            std::vector<uint8_t> src(n), bounce(n), dst(n);
            CopyPool &pool = CopyPool::get();
            for (int r = 0; r < reps; r++) {
                for (size_t i = 0; i < n; i += 4099) src[i] = (uint8_t)(i + r + t);
                // relay_h2d pattern
                const int nj = (int)((n + job - 1) / job);
                std::atomic<int> pend[64];
                CopyPool::Job jobs[64];
                for (int i = 0; i < nj; i++) {
                    const size_t o = (size_t)i * job;
                    pend[i].store(1);
                    jobs[i] = CopyPool::Job{bounce.data() + o, src.data() + o, n - o < job ? n - o : job, &pend[i], (uint8_t)((r & 1) ? CopyPool::kStreamingStores : CopyPool::kPlain)};
                }
                pool.submit(jobs, nj);
                for (int i = 0; i < nj; i++) pool.help_until([&] { return pend[i].load(std::memory_order_acquire) == 0; });
                // "kernel": flag set by the device thread
                flags[t][15].v = 0;
                { std::lock_guard<std::mutex> lk(fm); fq.push_back(&flags[t][15].v); }
                pool.help_until([&] { return flags[t][15].v != 0u; });
                // relay_d2h pattern: 3 pieces, flags, then jobs on one counter
                for (int k = 0; k < 3; k++) {
                    flags[t][k].v = 0;
                    std::lock_guard<std::mutex> lk(fm);
                    fq.push_back(&flags[t][k].v);
                }
                std::atomic<int> pd{0};
                for (int k = 0; k < 3; k++) {
                    pool.help_until([&] { return flags[t][k].v != 0u; });
                    CopyPool::Job js[8];
                    int c = 0;
                    const size_t o = (size_t)k << 20;
                    for (size_t q = 0; q < ((size_t)1 << 20); q += job) js[c++] = CopyPool::Job{dst.data() + o + q, bounce.data() + o + q, job, &pd};
                    pd.fetch_add(c);
                    pool.submit(js, c);
                }
                pool.help_until([&] { return pd.load(std::memory_order_acquire) == 0; });
                if (memcmp(src.data(), dst.data(), n)) bad++;
            }
        });
    for (auto &x : th) x.join();
    stop = true;
    device.join();
    printf("%s: %d threads x %d reps\n", bad.load() ? "MISMATCH" : "ok", T, reps);
    return bad.load() ? 1 : 0;
}
