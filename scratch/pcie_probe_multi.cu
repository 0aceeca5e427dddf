// Host<->device copy ceiling of a multi-GPU box, with no codec in the way.
//
//   scratch/pcie_probe_multi [--gpus 1,2,4,8] [--mib 1024] [--chunk 64] [--reps 4] [--modes proc,thread]
//
// For every N in --gpus and every mode, N workers (forked processes, or threads of one process)
// each own one GPU, `mib` MiB of pinned input, `mib` MiB of pinned output and two streams, and move
// the buffers in `chunk`-MiB cudaMemcpyAsync pieces: H2D alone, D2H alone, and both at once (the
// codec's end-to-end leg is a duplex stream: raw frames one way, records the other, per direction
// of the codec).  All workers start from a common barrier; the figure is total bytes / the slowest
// worker's time.  One JSON line per measurement.  Also: a multi-threaded host memcpy figure, because
// every byte of the e2e path is also read or written by host DRAM.
#include <cuda_runtime.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/wait.h>
#include <unistd.h>

#include <chrono>
#include <string>
#include <thread>
#include <vector>

#define CK(x)                                                                                  \
    do {                                                                                       \
        cudaError_t e_ = (x);                                                                  \
        if (e_ != cudaSuccess) {                                                               \
            fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); \
            _exit(3);                                                                          \
        }                                                                                      \
    } while (0)

struct Shared {
    pthread_barrier_t bar;
    double secs[3][16];      // [h2d, d2h, duplex][worker]
};

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

static void worker(int dev, Shared *sh, int slot, size_t bytes, size_t chunk, int reps) {
    CK(cudaSetDevice(dev));
    uint8_t *h_in, *h_out, *d_in, *d_out;
    CK(cudaHostAlloc(&h_in, bytes, cudaHostAllocDefault));
    CK(cudaHostAlloc(&h_out, bytes, cudaHostAllocDefault));
    CK(cudaMalloc(&d_in, bytes));
    CK(cudaMalloc(&d_out, bytes));
    memset(h_in, 1, bytes);
    memset(h_out, 2, bytes);
    CK(cudaMemset(d_out, 3, bytes));
    cudaStream_t s1, s2;
    CK(cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
    auto pass = [&](bool h2d, bool d2h) {
        for (size_t o = 0; o < bytes; o += chunk) {
            const size_t n = bytes - o < chunk ? bytes - o : chunk;
            if (h2d) CK(cudaMemcpyAsync(d_in + o, h_in + o, n, cudaMemcpyHostToDevice, s1));
            if (d2h) CK(cudaMemcpyAsync(h_out + o, d_out + o, n, cudaMemcpyDeviceToHost, s2));
        }
    };
    pass(true, true);
    CK(cudaDeviceSynchronize());
    for (int m = 0; m < 3; m++) {
        const bool h2d = m != 1, d2h = m != 0;
        pthread_barrier_wait(&sh->bar);
        const double t0 = now();
        for (int r = 0; r < reps; r++) pass(h2d, d2h);
        CK(cudaStreamSynchronize(s1));
        CK(cudaStreamSynchronize(s2));
        sh->secs[m][slot] = now() - t0;
        pthread_barrier_wait(&sh->bar);
    }
    CK(cudaFreeHost(h_in));
    CK(cudaFreeHost(h_out));
    CK(cudaFree(d_in));
    CK(cudaFree(d_out));
}

static std::vector<int> ints(const char *s) {
    std::vector<int> v;
    for (const char *p = s; *p;) {
        v.push_back(atoi(p));
        while (*p && *p != ',') p++;
        if (*p) p++;
    }
    return v;
}

static void host_memcpy_probe(int threads, size_t bytes) {
    std::vector<uint8_t *> a(threads), b(threads);
    for (int t = 0; t < threads; t++) {
        a[t] = (uint8_t *)malloc(bytes);
        b[t] = (uint8_t *)malloc(bytes);
        memset(a[t], 1, bytes);
        memset(b[t], 2, bytes);
    }
    const double t0 = now();
    std::vector<std::thread> th;
    for (int t = 0; t < threads; t++) th.emplace_back([&, t] { for (int r = 0; r < 4; r++) memcpy(b[t], a[t], bytes); });
    for (auto &x : th) x.join();
    const double dt = now() - t0;
    printf("{\"probe\": \"host_memcpy\", \"threads\": %d, \"copy_GBps\": %.1f, \"note\": \"bytes copied per second; DRAM traffic is 2-3x that\"}\n",
           threads, 4.0 * threads * bytes / dt / 1e9);
    for (int t = 0; t < threads; t++) { free(a[t]); free(b[t]); }
}

int main(int argc, char **argv) {
    std::vector<int> gpus = {1, 2, 4, 8};
    std::vector<std::string> modes = {"proc", "thread"};
    size_t mib = 1024, chunk = 64;
    int reps = 4;
    for (int i = 1; i + 1 < argc; i += 2) {
        if (!strcmp(argv[i], "--gpus")) gpus = ints(argv[i + 1]);
        else if (!strcmp(argv[i], "--mib")) mib = atol(argv[i + 1]);
        else if (!strcmp(argv[i], "--chunk")) chunk = atol(argv[i + 1]);
        else if (!strcmp(argv[i], "--reps")) reps = atoi(argv[i + 1]);
        else if (!strcmp(argv[i], "--modes")) {
            modes.clear();
            std::string s = argv[i + 1];
            size_t p = 0;
            while (p <= s.size()) {
                size_t q = s.find(',', p);
                if (q == std::string::npos) q = s.size();
                modes.push_back(s.substr(p, q - p));
                p = q + 1;
            }
        }
    }
    // the device count is taken in a child: the parent must not hold a CUDA context across fork()
    int ndev = 0;
    {
        int pfd[2];
        if (pipe(pfd)) return 2;
        pid_t c = fork();
        if (c == 0) {
            int n = 0;
            if (cudaGetDeviceCount(&n) != cudaSuccess) n = 0;
            if (write(pfd[1], &n, sizeof n) != sizeof n) _exit(1);
            _exit(0);
        }
        if (read(pfd[0], &ndev, sizeof ndev) != sizeof ndev) ndev = 0;
        waitpid(c, nullptr, 0);
    }
    const long cpus = sysconf(_SC_NPROCESSORS_ONLN);
    printf("{\"probe\": \"box\", \"gpus_visible\": %d, \"host_cpus\": %ld, \"mib_per_direction_per_gpu\": %zu, \"chunk_mib\": %zu, \"reps\": %d}\n",
           ndev, cpus, mib, chunk, reps);
    fflush(stdout);
    for (int t : {1, 4, 16}) if (t <= cpus) host_memcpy_probe(t, (size_t)256 << 20);
    fflush(stdout);
    const size_t bytes = mib << 20, cbytes = chunk << 20;
    for (int N : gpus) {
        if (N > ndev || N > 16) continue;
        for (const std::string &mode : modes) {
            Shared *sh = (Shared *)mmap(nullptr, sizeof(Shared), PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
            pthread_barrierattr_t at;
            pthread_barrierattr_init(&at);
            pthread_barrierattr_setpshared(&at, PTHREAD_PROCESS_SHARED);
            pthread_barrier_init(&sh->bar, &at, N);
            if (mode == "proc") {
                std::vector<pid_t> kids;
                for (int g = 0; g < N; g++) {
                    pid_t c = fork();
                    if (c == 0) {
                        worker(g, sh, g, bytes, cbytes, reps);
                        _exit(0);
                    }
                    kids.push_back(c);
                }
                bool bad = false;
                for (pid_t c : kids) {
                    int st = 0;
                    waitpid(c, &st, 0);
                    if (!WIFEXITED(st) || WEXITSTATUS(st)) bad = true;
                }
                if (bad) { printf("{\"probe\": \"pcie\", \"mode\": \"proc\", \"gpus\": %d, \"error\": \"a worker failed\"}\n", N); continue; }
            } else {
                pid_t c = fork();                      // a fresh process per configuration: no context survives
                if (c == 0) {
                    std::vector<std::thread> th;
                    for (int g = 0; g < N; g++) th.emplace_back(worker, g, sh, g, bytes, cbytes, reps);
                    for (auto &x : th) x.join();
                    _exit(0);
                }
                int st = 0;
                waitpid(c, &st, 0);
                if (!WIFEXITED(st) || WEXITSTATUS(st)) { printf("{\"probe\": \"pcie\", \"mode\": \"thread\", \"gpus\": %d, \"error\": \"failed\"}\n", N); continue; }
            }
            const char *names[3] = {"h2d", "d2h", "duplex"};
            for (int m = 0; m < 3; m++) {
                double worst = 0, best = 1e30;
                for (int g = 0; g < N; g++) {
                    if (sh->secs[m][g] > worst) worst = sh->secs[m][g];
                    if (sh->secs[m][g] < best) best = sh->secs[m][g];
                }
                const double per_dir = (double)N * bytes * reps / worst / 1e9;
                printf("{\"probe\": \"pcie\", \"mode\": \"%s\", \"gpus\": %d, \"dir\": \"%s\", \"GBps_per_direction\": %.1f, \"GBps_total\": %.1f, "
                       "\"slowest_worker_s\": %.4f, \"fastest_worker_s\": %.4f}\n",
                       mode.c_str(), N, names[m], per_dir, m == 2 ? 2 * per_dir : per_dir, worst, best);
            }
            fflush(stdout);
            munmap(sh, sizeof(Shared));
        }
    }
    return 0;
}
