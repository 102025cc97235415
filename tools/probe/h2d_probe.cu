// h2d_probe.cu -- what bounds the end-to-end path of an N-GPU box: aggregate pinned host <-> device copy bandwidth,
// independent of torch and of the library.  One process with N threads (what cli/bin/hdr2yuv --devices N does) or N
// processes (what bench.py under torchrun does); host memory from cudaHostAlloc (default / write-combined) or malloc +
// cudaHostRegister; optional CPU pinning of each worker to a block of cores.  Every worker copies H2D and D2H in the 2:1
// ratio of the forward path (6 B/px in, 3 B/px out) on two streams, as h2y_forward_host does.
//
//   h2d_probe <ngpus> <threads|procs> <pinned|wc|registered> <MB per copy> <iterations> [cores per worker, 0 = no pinning]
//
// Prints one line: aggregate H2D and D2H GB/s over all workers (wall clock between two barriers).
#include <cuda_runtime.h>
#include <pthread.h>
#include <sched.h>
#include <sys/mman.h>
#include <sys/wait.h>
#include <unistd.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(2); } } while (0)

struct Shared { pthread_barrier_t start, stop; };

static void worker(int dev, const char *alloc, size_t bytes, int iters, int cores, Shared *sh)
{
    if (cores > 0) {
        cpu_set_t set;
        CPU_ZERO(&set);
        for (int c = 0; c < cores; c++) CPU_SET(dev * cores + c, &set);
        sched_setaffinity(0, sizeof(set), &set);
    }
    CK(cudaSetDevice(dev));
    void *hin = nullptr, *hout = nullptr, *din, *dout;
    const size_t obytes = bytes / 2;
    if (!strcmp(alloc, "pinned")) { CK(cudaHostAlloc(&hin, bytes, cudaHostAllocDefault)); CK(cudaHostAlloc(&hout, obytes, cudaHostAllocDefault)); }
    else if (!strcmp(alloc, "wc")) { CK(cudaHostAlloc(&hin, bytes, cudaHostAllocWriteCombined)); CK(cudaHostAlloc(&hout, obytes, cudaHostAllocDefault)); }
    else {
        hin = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
        hout = mmap(nullptr, obytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
        madvise(hin, bytes, MADV_HUGEPAGE); madvise(hout, obytes, MADV_HUGEPAGE);
        memset(hin, 1, bytes); memset(hout, 1, obytes);
        CK(cudaHostRegister(hin, bytes, cudaHostRegisterDefault)); CK(cudaHostRegister(hout, obytes, cudaHostRegisterDefault));
    }
    memset(hin, 3, bytes);
    CK(cudaMalloc(&din, bytes)); CK(cudaMalloc(&dout, obytes));
    cudaStream_t s1, s2;
    CK(cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
    for (int i = 0; i < 2; i++) { CK(cudaMemcpyAsync(din, hin, bytes, cudaMemcpyHostToDevice, s1)); CK(cudaMemcpyAsync(hout, dout, obytes, cudaMemcpyDeviceToHost, s2)); }
    CK(cudaDeviceSynchronize());
    pthread_barrier_wait(&sh->start);
    for (int i = 0; i < iters; i++) { CK(cudaMemcpyAsync(din, hin, bytes, cudaMemcpyHostToDevice, s1)); CK(cudaMemcpyAsync(hout, dout, obytes, cudaMemcpyDeviceToHost, s2)); }
    CK(cudaDeviceSynchronize());
    pthread_barrier_wait(&sh->stop);
}

int main(int argc, char **argv)
{
    if (argc < 6) { fprintf(stderr, "usage: h2d_probe ngpus threads|procs pinned|wc|registered MB iters [cores per worker]\n"); return 1; }
    const int n = atoi(argv[1]), iters = atoi(argv[5]), cores = argc > 6 ? atoi(argv[6]) : 0;
    const bool procs = !strcmp(argv[2], "procs");
    const char *alloc = argv[3];
    const size_t bytes = (size_t)atoi(argv[4]) << 20;
    Shared *sh = (Shared *)mmap(nullptr, sizeof(Shared), PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
    pthread_barrierattr_t at;
    pthread_barrierattr_init(&at);
    pthread_barrierattr_setpshared(&at, PTHREAD_PROCESS_SHARED);
    pthread_barrier_init(&sh->start, &at, n + 1);
    pthread_barrier_init(&sh->stop, &at, n + 1);
    std::vector<std::thread> th;
    std::vector<pid_t> kids;
    for (int d = 0; d < n; d++) {
        if (procs) {
            pid_t p = fork();
            if (p == 0) { worker(d, alloc, bytes, iters, cores, sh); _exit(0); }
            kids.push_back(p);
        } else th.emplace_back(worker, d, alloc, bytes, iters, cores, sh);
    }
    pthread_barrier_wait(&sh->start);
    const auto t0 = std::chrono::steady_clock::now();
    pthread_barrier_wait(&sh->stop);
    const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    for (auto &t : th) t.join();
    for (pid_t p : kids) { int st; waitpid(p, &st, 0); }
    const double h2d = (double)n * iters * bytes / s * 1e-9, d2h = h2d / 2;
    printf("{\"gpus\": %d, \"mode\": \"%s\", \"alloc\": \"%s\", \"cores_per_worker\": %d, \"mb_per_copy\": %d, \"h2d_gbs_total\": %.1f, \"d2h_gbs_total\": %.1f, \"h2d_gbs_per_gpu\": %.1f}\n",
           n, argv[2], alloc, cores, atoi(argv[4]), h2d, d2h, h2d / n);
    return 0;
}
