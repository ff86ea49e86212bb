// numa_probe.cu — where does page-locked host memory have to live for each GPU's PCIe copies to run at link speed,
// and what does this box let a process find out / control?  Evidence for DESIGN.md §7 (host path placement).
//   build: nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o build/numa_probe tools/numa_probe.cu
//   run:   build/numa_probe [log2_bytes=29]
// Prints: allowed CPUs / memory nodes, cgroup limits, per-device PCI id + sysfs numa_node + cudaDevAttrHostNumaId,
// whether mbind / set_mempolicy / move_pages work, then copy-only GB/s per device x placement (default cudaHostAlloc,
// bound to each node, interleaved), alone and with all devices running concurrently.
#include <cuda_runtime.h>
#include <sys/mman.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <cctype>
#include <cerrno>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#define CK(x)                                                                                         \
    do {                                                                                              \
        cudaError_t e_ = (x);                                                                         \
        if (e_ != cudaSuccess) { std::printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); std::exit(2); } \
    } while (0)

static const int MPOL_DEFAULT_ = 0, MPOL_PREFERRED_ = 1, MPOL_BIND_ = 2, MPOL_INTERLEAVE_ = 3;
static long sys_mbind(void* p, size_t len, int mode, const unsigned long* mask, unsigned long maxnode, unsigned flags) {
    return syscall(SYS_mbind, p, len, mode, mask, maxnode, flags);
}
static long sys_set_mempolicy(int mode, const unsigned long* mask, unsigned long maxnode) {
    return syscall(SYS_set_mempolicy, mode, mask, maxnode);
}
static long sys_move_pages(int pid, unsigned long n, void** pages, const int* nodes, int* status, int flags) {
    return syscall(SYS_move_pages, pid, n, pages, nodes, status, flags);
}

static std::string slurp(const std::string& path) {
    FILE* f = std::fopen(path.c_str(), "r");
    if (!f) return "<unreadable>";
    char buf[4096];
    size_t n = std::fread(buf, 1, sizeof(buf) - 1, f);
    std::fclose(f);
    buf[n] = 0;
    while (n && (buf[n - 1] == '\n' || buf[n - 1] == ' ')) buf[--n] = 0;
    return buf;
}
static void grep_status(const char* key) {
    FILE* f = std::fopen("/proc/self/status", "r");
    if (!f) return;
    char line[1024];
    while (std::fgets(line, sizeof(line), f))
        if (std::strncmp(line, key, std::strlen(key)) == 0) std::printf("  %s", line);
    std::fclose(f);
}

struct HostBuf {
    void* p = nullptr;
    size_t bytes = 0;
    bool mapped = false;  // mmap + cudaHostRegister (else cudaHostAlloc)
    std::string label;
};

// node < 0: -1 default policy cudaHostAlloc, -2 interleave over all nodes
static HostBuf make_buf(size_t bytes, int node, int n_nodes, const char* label) {
    HostBuf b;
    b.bytes = bytes;
    b.label = label;
    if (node == -1) {
        CK(cudaHostAlloc(&b.p, bytes, cudaHostAllocDefault));
        std::memset(b.p, 1, bytes);
        return b;
    }
    void* p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (p == MAP_FAILED) { std::printf("mmap failed\n"); std::exit(2); }
    unsigned long mask = 0;
    int mode = MPOL_BIND_;
    if (node == -2) { mask = (1ul << n_nodes) - 1; mode = MPOL_INTERLEAVE_; }
    else mask = 1ul << node;
    long r = sys_mbind(p, bytes, mode, &mask, 64, 0);
    if (r != 0) std::printf("  mbind(%s) failed: errno %d (%s)\n", label, errno, std::strerror(errno));
    std::memset(p, 1, bytes);
    CK(cudaHostRegister(p, bytes, cudaHostRegisterDefault));
    b.p = p;
    b.mapped = true;
    return b;
}
static void where_is(const HostBuf& b) {
    const int n = 8;
    void* pages[n];
    int status[n];
    for (int i = 0; i < n; ++i) pages[i] = static_cast<char*>(b.p) + (b.bytes / n) * i;
    long r = sys_move_pages(0, n, pages, nullptr, status, 0);
    std::printf("  %-12s pages on nodes:", b.label.c_str());
    if (r != 0) { std::printf(" move_pages failed errno %d\n", errno); return; }
    for (int i = 0; i < n; ++i) std::printf(" %d", status[i]);
    std::printf("\n");
}

struct Dev {
    int id;
    void* d = nullptr;
    cudaStream_t s_in = nullptr, s_out = nullptr;
    cudaEvent_t e0, e1, f0, f1;
};

// every listed device copies `bytes` D2H (and optionally bytes/2 H2D on a second stream) reps times; returns wall seconds
static double run(std::vector<Dev>& devs, const std::vector<int>& which, const std::vector<void*>& host, size_t bytes, int reps,
                  bool d2h, bool h2d) {
    for (int k : which) { CK(cudaSetDevice(devs[k].id)); CK(cudaDeviceSynchronize()); }
    auto t0 = std::chrono::steady_clock::now();
    for (int r = 0; r < reps; ++r)
        for (size_t i = 0; i < which.size(); ++i) {
            Dev& dv = devs[which[i]];
            CK(cudaSetDevice(dv.id));
            char* h = static_cast<char*>(host[i]);
            if (d2h) CK(cudaMemcpyAsync(h, dv.d, bytes, cudaMemcpyDeviceToHost, dv.s_out));
            if (h2d) CK(cudaMemcpyAsync(static_cast<char*>(dv.d) + bytes, h + bytes, d2h ? bytes / 2 : bytes, cudaMemcpyHostToDevice, dv.s_in));
        }
    for (int k : which) { CK(cudaSetDevice(devs[k].id)); CK(cudaDeviceSynchronize()); }
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

int main(int argc, char** argv) {
    const int lg = argc > 1 ? std::atoi(argv[1]) : 29;
    const size_t bytes = size_t(1) << lg;  // per-direction copy size; host buffers are 2x (second half = H2D source)
    std::printf("== process\n");
    grep_status("Cpus_allowed_list");
    grep_status("Mems_allowed_list");
    std::printf("  cpu.max: %s\n  cpuset.cpus.effective: %s\n  cpuset.mems.effective: %s\n", slurp("/sys/fs/cgroup/cpu.max").c_str(),
                slurp("/sys/fs/cgroup/cpuset.cpus.effective").c_str(), slurp("/sys/fs/cgroup/cpuset.mems.effective").c_str());
    std::printf("  nodes online: %s   possible: %s   has_memory: %s\n", slurp("/sys/devices/system/node/online").c_str(),
                slurp("/sys/devices/system/node/possible").c_str(), slurp("/sys/devices/system/node/has_memory").c_str());
    int n_nodes = 0;
    for (int n = 0; n < 16; ++n) {
        std::string c = slurp("/sys/devices/system/node/node" + std::to_string(n) + "/cpulist");
        if (c == "<unreadable>") break;
        std::string m = slurp("/sys/devices/system/node/node" + std::to_string(n) + "/meminfo");
        size_t a = m.find("MemTotal"), f = m.find("MemFree");
        std::printf("  node%d cpus %s | %s | %s\n", n, c.c_str(), a == std::string::npos ? "?" : m.substr(a, m.find('\n', a) - a).c_str(),
                    f == std::string::npos ? "?" : m.substr(f, m.find('\n', f) - f).c_str());
        n_nodes = n + 1;
    }
    if (n_nodes == 0) n_nodes = 1;

    std::printf("== syscalls\n");
    {
        unsigned long mask = 1;
        long r = sys_set_mempolicy(MPOL_PREFERRED_, &mask, 64);
        std::printf("  set_mempolicy(PREFERRED,node0): %s\n", r == 0 ? "ok" : std::strerror(errno));
        sys_set_mempolicy(MPOL_DEFAULT_, nullptr, 0);
        void* p = mmap(nullptr, 1 << 21, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
        r = sys_mbind(p, 1 << 21, MPOL_BIND_, &mask, 64, 0);
        std::printf("  mbind(BIND,node0): %s\n", r == 0 ? "ok" : std::strerror(errno));
        munmap(p, 1 << 21);
    }

    int n_dev = 0;
    CK(cudaGetDeviceCount(&n_dev));
    std::printf("== devices: %d\n", n_dev);
    std::vector<Dev> devs(n_dev);
    for (int i = 0; i < n_dev; ++i) {
        Dev& dv = devs[i];
        dv.id = i;
        CK(cudaSetDevice(i));
        char id[32] = {0};
        CK(cudaDeviceGetPCIBusId(id, sizeof(id), i));
        for (char* c = id; *c; ++c) *c = char(std::tolower(*c));
        int host_numa = -99, numa_id = -99, numa_cfg = -99;
        cudaError_t e1 = cudaDeviceGetAttribute(&host_numa, cudaDevAttrHostNumaId, i);
        cudaError_t e2 = cudaDeviceGetAttribute(&numa_id, cudaDevAttrNumaId, i);
        cudaError_t e3 = cudaDeviceGetAttribute(&numa_cfg, cudaDevAttrNumaConfig, i);
        (void)cudaGetLastError();
        std::printf("  dev%d pci %s sysfs numa_node=%s local_cpulist=%s | HostNumaId=%d(%s) NumaId=%d(%s) NumaConfig=%d(%s)\n", i, id,
                    slurp(std::string("/sys/bus/pci/devices/") + id + "/numa_node").c_str(),
                    slurp(std::string("/sys/bus/pci/devices/") + id + "/local_cpulist").c_str(), host_numa, cudaGetErrorName(e1), numa_id,
                    cudaGetErrorName(e2), numa_cfg, cudaGetErrorName(e3));
        CK(cudaMalloc(&dv.d, 2 * bytes));
        CK(cudaMemset(dv.d, 3, 2 * bytes));
        CK(cudaStreamCreateWithFlags(&dv.s_in, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&dv.s_out, cudaStreamNonBlocking));
    }

    // placements: default, each node, interleave.  One buffer per (device, placement) so that concurrent runs do not share pages.
    struct Placement { int node; std::string label; };
    std::vector<Placement> pls;
    pls.push_back({-1, "default"});
    for (int n = 0; n < n_nodes; ++n) pls.push_back({n, "node" + std::to_string(n)});
    if (n_nodes > 1) pls.push_back({-2, "interleave"});
    std::printf("== host buffers (2 x %zu MiB each)\n", bytes >> 20);
    std::vector<std::vector<HostBuf>> bufs(n_dev);
    for (int i = 0; i < n_dev; ++i) {
        CK(cudaSetDevice(i));
        for (auto& pl : pls) {
            bufs[i].push_back(make_buf(2 * bytes, pl.node, n_nodes, pl.label.c_str()));
            if (i == 0) where_is(bufs[i].back());
        }
    }

    const int reps = 4;
    std::printf("== one device at a time, GB/s  [D2H | H2D | D2H with H2D/2 concurrently (D2H rate)]\n");
    for (int i = 0; i < n_dev; ++i)
        for (size_t p = 0; p < pls.size(); ++p) {
            std::vector<int> w{i};
            std::vector<void*> h{bufs[i][p].p};
            run(devs, w, h, bytes, 1, true, true);
            double a = run(devs, w, h, bytes, reps, true, false);
            double b = run(devs, w, h, bytes, reps, false, true);
            double c = run(devs, w, h, bytes, reps, true, true);
            std::printf("  dev%d %-11s D2H %6.1f | H2D %6.1f | both: D2H %6.1f\n", i, pls[p].label.c_str(), reps * bytes / a / 1e9,
                        reps * bytes / b / 1e9, reps * bytes / c / 1e9);
        }
    if (n_dev > 1) {
        std::printf("== all %d devices concurrently, aggregate GB/s [D2H | H2D | D2H with H2D/2]\n", n_dev);
        std::vector<int> all;
        for (int i = 0; i < n_dev; ++i) all.push_back(i);
        for (size_t p = 0; p < pls.size(); ++p) {
            std::vector<void*> h;
            for (int i = 0; i < n_dev; ++i) h.push_back(bufs[i][p].p);
            run(devs, all, h, bytes, 1, true, true);
            double a = run(devs, all, h, bytes, reps, true, false);
            double b = run(devs, all, h, bytes, reps, false, true);
            double c = run(devs, all, h, bytes, reps, true, true);
            std::printf("  all on %-11s D2H %6.1f | H2D %6.1f | both: D2H %6.1f\n", pls[p].label.c_str(), n_dev * reps * bytes / a / 1e9,
                        n_dev * reps * bytes / b / 1e9, n_dev * reps * bytes / c / 1e9);
        }
        // spread: device i on node (i * n_nodes / n_dev) and the reverse
        if (n_nodes > 1) {
            for (int rev = 0; rev < 2; ++rev) {
                std::vector<void*> h;
                for (int i = 0; i < n_dev; ++i) {
                    int node = i * n_nodes / n_dev;
                    if (rev) node = n_nodes - 1 - node;
                    h.push_back(bufs[i][1 + node].p);
                }
                run(devs, all, h, bytes, 1, true, true);
                double a = run(devs, all, h, bytes, reps, true, false);
                double c = run(devs, all, h, bytes, reps, true, true);
                std::printf("  spread %-10s D2H %6.1f | both: D2H %6.1f\n", rev ? "reversed" : "by-index", n_dev * reps * bytes / a / 1e9,
                            n_dev * reps * bytes / c / 1e9);
            }
        }
        // scaling of the default placement with the number of active devices
        for (int k = 1; k <= n_dev; k *= 2) {
            std::vector<int> w;
            std::vector<void*> h;
            for (int i = 0; i < k; ++i) { w.push_back(i); h.push_back(bufs[i][0].p); }
            double a = run(devs, w, h, bytes, reps, true, false);
            std::printf("  first %d devices, default placement: D2H aggregate %6.1f\n", k, k * reps * bytes / a / 1e9);
        }
    }
    std::printf("done\n");
    return 0;
}
