// fl_oracle.cpp — TEST INFRASTRUCTURE (see fl_oracle.h).  ISA dispatch + threading + C ABI.
#include "fl_oracle.h"

#include <algorithm>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

extern "C" {
#define FLO_DECL(ns, tb)                                                                              \
    void ns##_run_u##tb(int, unsigned, size_t, size_t, const void*, void*, const void*, const void*, \
                        uint64_t);                                                                    \
    uint64_t ns##_single_u##tb(unsigned, const void*, size_t);
#define FLO_DECL_NS(ns) FLO_DECL(ns, 8) FLO_DECL(ns, 16) FLO_DECL(ns, 32) FLO_DECL(ns, 64)
FLO_DECL_NS(flo_v2)
FLO_DECL_NS(flo_v3)
FLO_DECL_NS(flo_v4)
}

namespace {

using run_fn = void (*)(int, unsigned, size_t, size_t, const void*, void*, const void*, const void*, uint64_t);
using single_fn = uint64_t (*)(unsigned, const void*, size_t);

struct IsaTable {
    run_fn run[4];
    single_fn single[4];
    const char* name;
};

#define FLO_TAB(ns, nm)                                                                 \
    IsaTable { {ns##_run_u8, ns##_run_u16, ns##_run_u32, ns##_run_u64},                  \
               {ns##_single_u8, ns##_single_u16, ns##_single_u32, ns##_single_u64}, nm }

const IsaTable kTabs[3] = {FLO_TAB(flo_v2, "x86-64-v2"), FLO_TAB(flo_v3, "x86-64-v3"),
                           FLO_TAB(flo_v4, "x86-64-v4")};

int max_level() {
    __builtin_cpu_init();
    const bool v3 = __builtin_cpu_supports("avx2") && __builtin_cpu_supports("bmi2") &&
                    __builtin_cpu_supports("fma");
    const bool v4 = v3 && __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") &&
                    __builtin_cpu_supports("avx512dq") && __builtin_cpu_supports("avx512vl") &&
                    __builtin_cpu_supports("avx512cd");
    return v4 ? 4 : (v3 ? 3 : 2);
}

int g_level = 0;
const IsaTable& tab() {
    if (g_level == 0) g_level = max_level();
    return kTabs[g_level - 2];
}

// Persistent worker pool: the timed CPU baseline must not pay a thread spawn per call (a 128-thread
// spawn costs milliseconds, comparable to the work of one width of the bench sample).
class Pool {
public:
    static Pool& get() { static Pool* p = new Pool; return *p; }  // leaked on purpose: workers outlive exit()
    // run fn(t) for t in [0, n) on n workers (the calling thread takes t = 0)
    void run(size_t n, const std::function<void(size_t)>& fn) {
        std::lock_guard<std::mutex> serial(run_mu_);
        ensure(n - 1);
        {
            std::lock_guard<std::mutex> lk(mu_);
            fn_ = &fn; active_ = n - 1; pending_ = n - 1; ++gen_;
        }
        cv_.notify_all();
        fn(0);
        std::unique_lock<std::mutex> lk(mu_);
        done_.wait(lk, [&] { return pending_ == 0; });
        fn_ = nullptr;
    }
private:
    void ensure(size_t n) {
        while (workers_.size() < n) {
            const size_t id = workers_.size();
            workers_.emplace_back([this, id] { loop(id); });
            workers_.back().detach();
        }
    }
    void loop(size_t id) {
        unsigned long seen = 0;
        for (;;) {
            const std::function<void(size_t)>* fn;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return gen_ != seen; });
                seen = gen_;
                if (id >= active_) continue;
                fn = fn_;
            }
            (*fn)(id + 1);
            {
                std::lock_guard<std::mutex> lk(mu_);
                if (--pending_ == 0) done_.notify_one();
            }
        }
    }
    std::mutex run_mu_, mu_;
    std::condition_variable cv_, done_;
    std::vector<std::thread> workers_;
    const std::function<void(size_t)>* fn_ = nullptr;
    size_t active_ = 0, pending_ = 0;
    unsigned long gen_ = 0;
};

int type_slot(int tbits) {
    switch (tbits) {
        case 8: return 0;
        case 16: return 1;
        case 32: return 2;
        case 64: return 3;
        default: return -1;
    }
}

}  // namespace

extern "C" {

const char* flo_isa(void) { return tab().name; }

int flo_set_isa_level(int level) {
    const int mx = max_level();
    g_level = std::max(2, std::min(level, mx));
    return g_level;
}

int flo_hardware_threads(void) {
    const unsigned n = std::thread::hardware_concurrency();
    return n ? int(n) : 1;
}

int flo_run(int tbits, int op, unsigned width, size_t n_blocks, const void* in, void* out, const void* base,
            const void* refs, uint64_t ref_scalar, int n_threads) {
    const int slot = type_slot(tbits);
    if (slot < 0) return FLO_ERR_TYPE;
    if (op < FLO_OP_PACK || op > FLO_OP_UNFOR_FILTER) return FLO_ERR_TYPE;
    const bool width_op = (op == FLO_OP_PACK || op == FLO_OP_UNPACK || op == FLO_OP_FOR_PACK ||
                           op == FLO_OP_UNFOR_PACK || op == FLO_OP_UNDELTA_PACK || op == FLO_OP_UNFOR_FILTER);
    if (width_op && width > unsigned(tbits)) return FLO_ERR_WIDTH;  // bitpacking.rs:93,126
    if (!width_op) width = 0;
    if (n_blocks == 0) return FLO_OK;
    const bool needs_in = !(width == 0 && (op == FLO_OP_UNPACK || op == FLO_OP_UNFOR_PACK || op == FLO_OP_UNDELTA_PACK ||
                                          op == FLO_OP_UNFOR_FILTER));
    const bool needs_out = !(width == 0 && (op == FLO_OP_PACK || op == FLO_OP_FOR_PACK));
    if ((needs_in && !in) || (needs_out && !out)) return FLO_ERR_NULL;
    if ((op == FLO_OP_DELTA || op == FLO_OP_UNDELTA || op == FLO_OP_UNDELTA_PACK || op == FLO_OP_UNFOR_FILTER) && !base)
        return FLO_ERR_NULL;
    const run_fn run = tab().run[slot];
    if (n_threads <= 1 || n_blocks < 2) {
        run(op, width, 0, n_blocks, in, out, base, refs, ref_scalar);
        return FLO_OK;
    }
    const size_t nt = std::min<size_t>(size_t(n_threads), n_blocks);
    Pool::get().run(nt, [=](size_t t) {
        const size_t b0 = n_blocks * t / nt, b1 = n_blocks * (t + 1) / nt;
        run(op, width, b0, b1, in, out, base, refs, ref_scalar);
    });
    return FLO_OK;
}

int flo_unpack_single(int tbits, unsigned width, const void* packed, size_t index, uint64_t* value) {
    const int slot = type_slot(tbits);
    if (slot < 0) return FLO_ERR_TYPE;
    if (width > unsigned(tbits)) return FLO_ERR_WIDTH;  // bitpacking.rs:197
    if (index >= 1024) return FLO_ERR_INDEX;            // bitpacking.rs:152
    if (!value || (width != 0 && !packed)) return FLO_ERR_NULL;
    *value = tab().single[slot](width, packed, index);
    return FLO_OK;
}

int flo_unpack_gather(int tbits, unsigned width, const void* packed, const uint64_t* global_index, size_t n,
                      void* out) {
    const int slot = type_slot(tbits);
    if (slot < 0) return FLO_ERR_TYPE;
    if (width > unsigned(tbits)) return FLO_ERR_WIDTH;
    if (n == 0) return FLO_OK;
    if (!global_index || !out || (width != 0 && !packed)) return FLO_ERR_NULL;
    const single_fn f = tab().single[slot];
    const size_t block_bytes = size_t(128) * width;
    for (size_t i = 0; i < n; ++i) {
        const uint64_t g = global_index[i];
        const uint64_t v = f(width, static_cast<const uint8_t*>(packed) + (g >> 10) * block_bytes, size_t(g & 1023));
        switch (tbits) {
            case 8: static_cast<uint8_t*>(out)[i] = uint8_t(v); break;
            case 16: static_cast<uint16_t*>(out)[i] = uint16_t(v); break;
            case 32: static_cast<uint32_t*>(out)[i] = uint32_t(v); break;
            default: static_cast<uint64_t*>(out)[i] = v; break;
        }
    }
    return FLO_OK;
}

}  // extern "C"
