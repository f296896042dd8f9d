// Probe: what limits the device -> pinned-host stream of the condensed matrix when several GPUs of one box copy at once?
// (Round-1 verdict: end-to-end weak-scaling efficiency 0.23 at 8 GPUs; the limiter is the cudaMemcpyAsync D2H stream of
// ct_rmsd_rows, csrc/engine.cu, into one host's memory.)  Separates machine from code:
//   * n = 1, 2, 4, 8 GPUs copying concurrently, each into its own pinned buffer, one host thread per GPU (the in-process
//     layout of clusttraj_b200.distmat.condensed_matrix) and one PROCESS per GPU (the torchrun layout of bench.py);
//   * pinned buffers: cudaHostAlloc default / write-combined / malloc + first touch by a thread bound to the GPU's local
//     CPUs (sysfs numa_node / local_cpulist of the PCI device) + cudaHostRegister;
//   * copy engine (cudaMemcpyAsync) against a kernel that stores straight into mapped host memory.
// Output: one JSON object on stdout (kept as profiles/r02_d2h_scaling.json).
//   nvcc -O3 -std=c++17 -o tools/d2h_probe tools/d2h_probe.cu -lpthread && tools/d2h_probe
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>
#include <sched.h>
#include <sys/wait.h>
#include <unistd.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)

static const size_t CHUNK = (size_t)256 << 20;   // bytes per copy (the engine streams ~256 MiB chunks)
static const int NCOPIES = 8;                    // copies per measurement: 2 GiB per GPU

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
static double wall() { return std::chrono::duration<double>(std::chrono::system_clock::now().time_since_epoch()).count(); }

__global__ void store_kernel(const double* __restrict__ src, double* __restrict__ dst, size_t n) {
    const size_t stride = (size_t)gridDim.x * blockDim.x * 2;
    for (size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 2; i + 1 < n; i += stride)
        *reinterpret_cast<double2*>(dst + i) = *reinterpret_cast<const double2*>(src + i);
}

static std::string pci_sysfs(int dev) {
    char id[32];
    CK(cudaDeviceGetPCIBusId(id, sizeof id, dev));
    for (char* p = id; *p; ++p) *p = (char)tolower(*p);
    return std::string("/sys/bus/pci/devices/") + id;
}
static std::string read_line(const std::string& path) {
    std::ifstream f(path);
    std::string s;
    if (f) std::getline(f, s);
    return s;
}
static std::vector<int> parse_cpulist(const std::string& s) {
    std::vector<int> out;
    std::stringstream ss(s);
    std::string tok;
    while (std::getline(ss, tok, ',')) {
        int a, b;
        if (sscanf(tok.c_str(), "%d-%d", &a, &b) == 2) for (int c = a; c <= b; ++c) out.push_back(c);
        else if (sscanf(tok.c_str(), "%d", &a) == 1) out.push_back(a);
    }
    return out;
}
static void bind_to(const std::vector<int>& cpus) {
    if (cpus.empty()) return;
    cpu_set_t set;
    CPU_ZERO(&set);
    for (int c : cpus) CPU_SET(c, &set);
    sched_setaffinity(0, sizeof set, &set);
}

enum Kind { PINNED = 0, WRITE_COMBINED = 1, NUMA_REGISTERED = 2, KERNEL_STORES = 3 };
static const char* kind_name[] = {"pinned_default", "write_combined", "first_touch_local_cpus_registered", "kernel_stores_mapped"};

struct Slot {
    int dev;
    double* d = nullptr;
    double* h = nullptr;
    void* raw = nullptr;
    cudaStream_t s;
    double seconds = 0.0;
};

static void setup(Slot& sl, Kind kind) {
    CK(cudaSetDevice(sl.dev));
    CK(cudaMalloc(&sl.d, CHUNK));
    CK(cudaMemset(sl.d, 1, CHUNK));
    CK(cudaStreamCreateWithFlags(&sl.s, cudaStreamNonBlocking));
    if (kind == PINNED) CK(cudaHostAlloc((void**)&sl.h, CHUNK * 2, cudaHostAllocPortable));
    else if (kind == WRITE_COMBINED) CK(cudaHostAlloc((void**)&sl.h, CHUNK * 2, cudaHostAllocPortable | cudaHostAllocWriteCombined));
    else if (kind == KERNEL_STORES) CK(cudaHostAlloc((void**)&sl.h, CHUNK * 2, cudaHostAllocPortable | cudaHostAllocMapped));
    else {
        bind_to(parse_cpulist(read_line(pci_sysfs(sl.dev) + "/local_cpulist")));
        if (posix_memalign(&sl.raw, 1 << 21, CHUNK * 2)) exit(2);
        memset(sl.raw, 0, CHUNK * 2);   // first touch on the bound CPUs
        CK(cudaHostRegister(sl.raw, CHUNK * 2, cudaHostRegisterPortable));
        sl.h = (double*)sl.raw;
    }
    CK(cudaDeviceSynchronize());
}
static void teardown(Slot& sl, Kind kind) {
    CK(cudaSetDevice(sl.dev));
    if (kind == NUMA_REGISTERED) { CK(cudaHostUnregister(sl.raw)); free(sl.raw); }
    else CK(cudaFreeHost(sl.h));
    CK(cudaFree(sl.d));
    CK(cudaStreamDestroy(sl.s));
}
static void copies(Slot& sl, Kind kind) {
    CK(cudaSetDevice(sl.dev));
    for (int c = 0; c < NCOPIES; ++c) {
        double* dst = sl.h + (c & 1) * (CHUNK / sizeof(double));
        if (kind == KERNEL_STORES) {
            double* dmap;
            CK(cudaHostGetDevicePointer((void**)&dmap, dst, 0));
            store_kernel<<<148 * 4, 256, 0, sl.s>>>(sl.d, dmap, CHUNK / sizeof(double));
        } else {
            CK(cudaMemcpyAsync(dst, sl.d, CHUNK, cudaMemcpyDeviceToHost, sl.s));
        }
    }
    CK(cudaStreamSynchronize(sl.s));
}

// n GPUs, one host thread each, one process
static void run_threads(int n, Kind kind, double* agg, double* slowest, double* fastest) {
    std::vector<Slot> slots(n);
    std::vector<std::thread> th;
    std::atomic<int> ready{0}, go{0};
    double t_begin[16], t_end[16];
    for (int k = 0; k < n; ++k) {
        slots[k].dev = k;
        th.emplace_back([&, k] {
            setup(slots[k], kind);
            copies(slots[k], kind);   // warm-up
            ready.fetch_add(1);
            while (!go.load()) std::this_thread::yield();
            t_begin[k] = now();
            copies(slots[k], kind);
            t_end[k] = now();
            teardown(slots[k], kind);
        });
    }
    while (ready.load() < n) std::this_thread::yield();
    go.store(1);
    for (auto& t : th) t.join();
    double b = t_begin[0], e = t_end[0];
    *slowest = 1e30; *fastest = 0.0;
    for (int k = 0; k < n; ++k) {
        b = std::min(b, t_begin[k]); e = std::max(e, t_end[k]);
        const double r = (double)CHUNK * NCOPIES / (t_end[k] - t_begin[k]) / 1e9;
        *slowest = std::min(*slowest, r); *fastest = std::max(*fastest, r);
    }
    *agg = (double)CHUNK * NCOPIES * n / (e - b) / 1e9;
}

// child of the multi-process variant: device k, start at wall-clock time t_start; prints "begin end" (wall clock)
static int child_main(int k, double t_start, int kind_i) {
    Slot sl; sl.dev = k;
    const Kind kind = (Kind)kind_i;
    setup(sl, kind);
    copies(sl, kind);
    while (wall() < t_start) {}
    const double b = wall();
    copies(sl, kind);
    const double e = wall();
    printf("%.6f %.6f\n", b, e);
    teardown(sl, kind);
    return 0;
}

static bool run_processes(const char* self, int n, Kind kind, double* agg, double* slowest) {
    const double t_start = wall() + 6.0 + 0.5 * n;   // CUDA start-up + pinning of every child must fit before this
    std::vector<FILE*> pipes;
    for (int k = 0; k < n; ++k) {
        char cmd[512];
        snprintf(cmd, sizeof cmd, "%s --child %d %.6f %d", self, k, t_start, (int)kind);
        FILE* p = popen(cmd, "r");
        if (!p) return false;
        pipes.push_back(p);
    }
    double b = 1e300, e = 0.0;
    *slowest = 1e30;
    bool ok = true;
    for (FILE* p : pipes) {
        double cb, ce;
        if (fscanf(p, "%lf %lf", &cb, &ce) == 2) {
            b = std::min(b, cb); e = std::max(e, ce);
            *slowest = std::min(*slowest, (double)CHUNK * NCOPIES / (ce - cb) / 1e9);
            if (cb > t_start + 0.25) ok = false;   // a child was late: the copies did not overlap as intended
        } else ok = false;
        pclose(p);
    }
    *agg = (double)CHUNK * NCOPIES * n / (e - b) / 1e9;
    return ok;
}

int main(int argc, char** argv) {
    if (argc >= 5 && !strcmp(argv[1], "--child")) return child_main(atoi(argv[2]), atof(argv[3]), atoi(argv[4]));
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    printf("{\"devices\": %d, \"host_cpus\": %ld, \"numa_nodes_online\": \"%s\",\n", ndev, sysconf(_SC_NPROCESSORS_ONLN),
           read_line("/sys/devices/system/node/online").c_str());
    printf(" \"bytes_per_gpu_per_measurement\": %zu, \"chunk_bytes\": %zu,\n \"gpus\": [\n", CHUNK * NCOPIES, CHUNK);
    for (int k = 0; k < ndev; ++k) {
        const std::string base = pci_sysfs(k);
        printf("  {\"device\": %d, \"pci\": \"%s\", \"numa_node\": \"%s\", \"local_cpulist\": \"%s\"}%s\n", k, base.c_str() + 21,
               read_line(base + "/numa_node").c_str(), read_line(base + "/local_cpulist").c_str(), k + 1 < ndev ? "," : "");
    }
    printf(" ],\n \"threads_one_process\": [\n");
    bool first = true;
    for (int kind = 0; kind < 4; ++kind) {
        for (int n = 1; n <= ndev; n *= 2) {
            double agg, slow, fast;
            run_threads(n, (Kind)kind, &agg, &slow, &fast);
            printf("  %s{\"buffers\": \"%s\", \"n_gpus\": %d, \"aggregate_gbs\": %.1f, \"slowest_gpu_gbs\": %.1f, \"fastest_gpu_gbs\": %.1f}\n",
                   first ? "" : ",", kind_name[kind], n, agg, slow, fast);
            fflush(stdout);
            first = false;
        }
    }
    printf(" ],\n \"one_process_per_gpu\": [\n");
    first = true;
    // (buffers, GPUs): one GPU alone, all GPUs, all GPUs with locally first-touched buffers
    const int plan[3][2] = {{PINNED, 1}, {PINNED, ndev}, {NUMA_REGISTERED, ndev}};
    for (int t = 0; t < 3; ++t) {
        if (t > 0 && ndev == 1) break;
        double agg, slow;
        const bool ok = run_processes(argv[0], plan[t][1], (Kind)plan[t][0], &agg, &slow);
        printf("  %s{\"buffers\": \"%s\", \"n_gpus\": %d, \"aggregate_gbs\": %.1f, \"slowest_gpu_gbs\": %.1f, \"overlapped\": %s}\n",
               first ? "" : ",", kind_name[plan[t][0]], plan[t][1], agg, slow, ok ? "true" : "false");
        fflush(stdout);
        first = false;
    }
    printf(" ]\n}\n");
    return 0;
}
