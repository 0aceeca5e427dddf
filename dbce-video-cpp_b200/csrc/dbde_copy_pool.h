// The process-wide pool that moves bytes between pageable caller memory and pinned bounce buffers
// (see "pageable host memory" in dbde_capi.cu).  Host-only and CUDA-free, so it is stress-tested on its
// own (tests/test_copy_pool.py builds scratch-free C++ around this header).
#pragma once
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

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

class CopyPool {
  public:
    struct Job {
        uint8_t *dst;
        const uint8_t *src;
        size_t n;
        std::atomic<int> *pending;       // decremented when the job is done (the submitter adds the jobs first)
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
        memcpy(j.dst, j.src, j.n);
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
