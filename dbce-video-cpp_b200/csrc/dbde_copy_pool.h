// The process-wide pool that moves bytes between pageable caller memory and pinned bounce buffers
// (see "pageable host memory" in dbde_capi.cu).  Host-only and CUDA-free, so it is stress-tested on its
// own (tests/test_copy_pool.py builds scratch-free C++ around this header).
#pragma once
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#if defined(__x86_64__)
#include <emmintrin.h>
#endif

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>

namespace dbde {
inline void cpu_pause() {
#if defined(__x86_64__) || defined(__i386__)
    __builtin_ia32_pause();
#else
    std::this_thread::yield();
#endif
}

// memcpy whose stores bypass the cache (dst 16-byte aligned): for bytes a DMA engine reads next.  Lines that
// sit dirty in a CPU cache make the device's reads slower than lines in DRAM (measured on the drop-in path's
// bounce buffers).
inline void copy_streaming(uint8_t *dst, const uint8_t *src, size_t n) {
#if defined(__x86_64__)
    if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
        size_t i = 0;
        for (; i + 64 <= n; i += 64) {
            const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + i));
            const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + i + 16));
            const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + i + 32));
            const __m128i d = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + i + 48));
            _mm_stream_si128(reinterpret_cast<__m128i *>(dst + i), a);
            _mm_stream_si128(reinterpret_cast<__m128i *>(dst + i + 16), b);
            _mm_stream_si128(reinterpret_cast<__m128i *>(dst + i + 32), c);
            _mm_stream_si128(reinterpret_cast<__m128i *>(dst + i + 48), d);
        }
        if (i < n) memcpy(dst + i, src + i, n - i);
        _mm_sfence();
        return;
    }
#endif
    memcpy(dst, src, n);
}

class CopyPool {
  public:
    enum Kind : uint8_t { kPlain = 0, kStreamingStores = 1 };
    struct Job {
        uint8_t *dst;
        const uint8_t *src;
        size_t n;
        std::atomic<int> *pending;       // decremented when the job is done (the submitter adds the jobs first)
        uint8_t kind = kPlain;           // kStreamingStores: the destination is read by a DMA engine next
    };
    static CopyPool &get() {
        static CopyPool p;
        return p;
    }
    void submit(const Job *jobs, int n) {
        if (n <= 0) return;
        {
            std::lock_guard<std::mutex> lk(m_);
            for (int i = 0; i < n; i++) q_.push_back(jobs[i]);
        }
        queued_.fetch_add(n, std::memory_order_release);
        if (sleepers_.load(std::memory_order_acquire) > 0) cv_.notify_all();
    }
    bool run_one() {
        if (queued_.load(std::memory_order_acquire) <= 0) return false;
        Job j;
        {
            std::lock_guard<std::mutex> lk(m_);
            if (q_.empty()) return false;
            j = q_.front();
            q_.pop_front();
        }
        queued_.fetch_sub(1, std::memory_order_relaxed);
        if (j.kind == kStreamingStores) copy_streaming(j.dst, j.src, j.n);
        else memcpy(j.dst, j.src, j.n);
        j.pending->fetch_sub(1, std::memory_order_release);
        return true;
    }
    // wait until done() holds, copying queued pieces (anyone's) meanwhile.  done() may be a driver query:
    // it is evaluated once per round, and an idle waiter backs off to a yield so that it does not starve
    // the threads it is waiting for when there are more callers than cores.
    template <class Done>
    void help_until(Done done) {
        helpers_.fetch_add(1, std::memory_order_relaxed);
        for (unsigned idle = 0; !done();) {
            if (run_one()) {
                idle = 0;
                continue;
            }
            // nothing to copy: a short spin for the low-latency case, then give the core away -- first to
            // whoever is runnable, and after ~a millisecond of that for a timed sleep, so that a crowd of
            // waiting callers never keeps every core busy while the thread they wait for needs one
            if (++idle < 32) cpu_pause();
            else if (idle < 2048) std::this_thread::yield();
            else std::this_thread::sleep_for(std::chrono::microseconds(50));
        }
        helpers_.fetch_sub(1, std::memory_order_relaxed);
    }
    int helpers() const { return helpers_.load(std::memory_order_relaxed); }

  private:
    CopyPool() {
        const int hc = (int)std::thread::hardware_concurrency();
        crowd_ = hc / 2 > 4 ? hc / 2 : 4;
        if (const char *e = getenv("DBDE_B200_COPY_CROWD")) crowd_ = atoi(e);
        int n = hc / 2 - 1;
        if (n > 6) n = 6;
        if (const char *e = getenv("DBDE_B200_COPY_THREADS")) n = atoi(e);
        if (n < 0) n = 0;
        for (int i = 0; i < n; i++) th_.emplace_back([this] { work(); });
    }
    ~CopyPool() {
        {
            std::lock_guard<std::mutex> lk(m_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto &t : th_) t.join();
    }
    void work() {
        using clock = std::chrono::steady_clock;
        for (;;) {
            // a caller in a loop finds the threads still awake (they poll for half a millisecond after
            // their last job); an idle process finds them asleep
            auto last = clock::now();
            for (unsigned spins = 0;; spins++) {
                // with that many callers already helping, the pool's threads would only add contention
                // for the cores (measured at 16 callers on 16 cores): stand back
                if (helpers_.load(std::memory_order_relaxed) >= crowd_) {
                    std::this_thread::sleep_for(std::chrono::microseconds(200));
                    last = clock::now();
                    continue;
                }
                if (run_one()) {
                    last = clock::now();
                    spins = 0;
                    continue;
                }
                cpu_pause();
                if ((spins & 255u) == 255u && clock::now() - last > std::chrono::microseconds(500)) break;
            }
            std::unique_lock<std::mutex> lk(m_);
            sleepers_.fetch_add(1, std::memory_order_release);
            cv_.wait(lk, [&] { return stop_ || !q_.empty(); });
            sleepers_.fetch_sub(1, std::memory_order_release);
            if (stop_) return;
        }
    }
    std::mutex m_;
    std::condition_variable cv_;
    std::deque<Job> q_;
    std::atomic<int> queued_{0}, sleepers_{0}, helpers_{0};
    int crowd_ = 8;
    std::vector<std::thread> th_;
    bool stop_ = false;
};

}  // namespace dbde
