// emu_block.h — block-level host emulation: one std::thread per simulated CUDA thread, std::barrier per block
// (__syncthreads) and per warp (full-mask shuffles / ballots / reductions), blocks run one after the other (so a
// `static` local stands in for __shared__).  Include after emu_shim.h.  C++20.
#pragma once
#include <atomic>
#include <barrier>
#include <functional>
#include <memory>
#include <pthread.h>
#include <thread>
#include <vector>

struct EmuWarp {
    std::barrier<> bar;
    uint32_t slot[32];
    explicit EmuWarp(int n) : bar(n) {}
};
static thread_local EmuWarp* emu_warp = nullptr;
static thread_local std::barrier<>* emu_block_bar = nullptr;
static thread_local unsigned emu_lane = 0;

alignas(16) static unsigned char emu_dyn_smem_buf[232 * 1024];
static inline void* emu_dyn_smem() { return emu_dyn_smem_buf; }

#undef __shared__
#define __shared__ static

static inline void __syncthreads() { emu_block_bar->arrive_and_wait(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { emu_warp->bar.arrive_and_wait(); }

template <typename T> static inline uint32_t emu_bits(T v) { static_assert(sizeof(T) == 4, "32-bit values only"); uint32_t u; memcpy(&u, &v, 4); return u; }
template <typename T> static inline T emu_from(uint32_t u) { T v; memcpy(&v, &u, 4); return v; }
// every lane publishes, everybody reads lane `src` (own value when src is out of range)
template <typename T> static inline T emu_exchange(T v, int src)
{
    emu_warp->slot[emu_lane] = emu_bits(v);
    emu_warp->bar.arrive_and_wait();
    const T r = (src >= 0 && src < 32) ? emu_from<T>(emu_warp->slot[src]) : v;
    emu_warp->bar.arrive_and_wait();
    return r;
}
template <typename T> static inline T __shfl_sync(unsigned, T v, int src) { return emu_exchange(v, src & 31); }
template <typename T> static inline T __shfl_xor_sync(unsigned, T v, int m) { return emu_exchange(v, (int)emu_lane ^ m); }
template <typename T> static inline T __shfl_down_sync(unsigned, T v, unsigned d) { return emu_exchange(v, (int)(emu_lane + d) < 32 ? (int)(emu_lane + d) : -1); }
template <typename T> static inline T __shfl_up_sync(unsigned, T v, unsigned d) { return emu_exchange(v, (int)emu_lane - (int)d); }
static inline unsigned __ballot_sync(unsigned, bool p)
{
    emu_warp->slot[emu_lane] = p ? 1u : 0u;
    emu_warp->bar.arrive_and_wait();
    unsigned r = 0;
    for (int i = 0; i < 32; ++i) r |= (emu_warp->slot[i] & 1u) << i;
    emu_warp->bar.arrive_and_wait();
    return r;
}
static inline bool __all_sync(unsigned m, bool p) { return __ballot_sync(m, p) == 0xffffffffu; }
static inline bool __any_sync(unsigned m, bool p) { return __ballot_sync(m, p) != 0u; }
static inline unsigned __reduce_max_sync(unsigned, unsigned v)
{
    emu_warp->slot[emu_lane] = v;
    emu_warp->bar.arrive_and_wait();
    unsigned r = 0;
    for (int i = 0; i < 32; ++i) r = emu_warp->slot[i] > r ? emu_warp->slot[i] : r;
    emu_warp->bar.arrive_and_wait();
    return r;
}
static inline unsigned __reduce_min_sync(unsigned, unsigned v)
{
    emu_warp->slot[emu_lane] = v;
    emu_warp->bar.arrive_and_wait();
    unsigned r = 0xffffffffu;
    for (int i = 0; i < 32; ++i) r = emu_warp->slot[i] < r ? emu_warp->slot[i] : r;
    emu_warp->bar.arrive_and_wait();
    return r;
}
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline float __fdividef(float a, float b) { return a / b; }       // approximate on the device: only feeds conservative bounds

// can this process have `n` more threads at once?  (containers cap pids; the tests skip instead of crashing)
static inline bool emu_probe_threads(int n)
{
    std::atomic<bool> go{false};
    std::vector<pthread_t> th;
    pthread_attr_t attr;
    pthread_attr_init(&attr);
    pthread_attr_setstacksize(&attr, 1u << 20);
    bool ok = true;
    for (int i = 0; i < n && ok; ++i) {
        pthread_t t;
        if (pthread_create(&t, &attr, [](void* p) -> void* { while (!static_cast<std::atomic<bool>*>(p)->load()) sched_yield(); return nullptr; }, &go) != 0) ok = false;
        else th.push_back(t);
    }
    go.store(true);
    for (auto& t : th) pthread_join(t, nullptr);
    pthread_attr_destroy(&attr);
    return ok;
}

// hook for emu_cluster.h: per-thread set-up / tear-down of the cluster context
struct EmuThreadHook { void (*enter)(void*); void (*leave)(void*); void* arg; };

// run `body` for every thread of ONE block (its threads concurrently; 1 MB stacks: a cluster is thousands of threads)
static inline void emu_run_block(unsigned bx, unsigned by, unsigned bz, dim3 grid, dim3 block, const std::function<void()>& body,
                                 const EmuThreadHook* hook = nullptr)
{
    const unsigned nthreads = block.x * block.y * block.z;
    std::barrier<> block_bar((std::ptrdiff_t)nthreads);
    std::vector<std::unique_ptr<EmuWarp>> warps;
    for (unsigned w = 0; w * 32 < nthreads; ++w) warps.emplace_back(new EmuWarp((int)((nthreads - w * 32) < 32 ? (nthreads - w * 32) : 32)));
    struct Start { std::function<void()> run; };
    std::vector<Start> starts(nthreads);
    std::vector<pthread_t> th(nthreads);
    pthread_attr_t attr;
    pthread_attr_init(&attr);
    pthread_attr_setstacksize(&attr, 1u << 20);
    for (unsigned t = 0; t < nthreads; ++t) {
        starts[t].run = [&, t]() {
            emu_threadIdx = { t % block.x, (t / block.x) % block.y, t / (block.x * block.y) };
            emu_blockIdx = { bx, by, bz };
            emu_blockDim = { block.x, block.y, block.z };
            emu_gridDim = { grid.x, grid.y, grid.z };
            emu_lane = t & 31;
            emu_warp = warps[t >> 5].get();
            emu_block_bar = &block_bar;
            if (hook) hook->enter(hook->arg);
            body();
            emu_warp->bar.arrive_and_drop();        // an exited thread no longer takes part in collectives
            block_bar.arrive_and_drop();
            if (hook) hook->leave(hook->arg);
        };
        if (pthread_create(&th[t], &attr, [](void* p) -> void* { static_cast<Start*>(p)->run(); return nullptr; }, &starts[t]) != 0) {
            fprintf(stderr, "tests/emu: cannot create thread %u of %u (pid limit?)\n", t, nthreads);
            abort();
        }
    }
    for (auto& x : th) pthread_join(x, nullptr);
    pthread_attr_destroy(&attr);
}

// run `body` for every thread of every block of the grid (blocks sequentially)
static inline void emu_launch(dim3 grid, dim3 block, const std::function<void()>& body)
{
    for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
            for (unsigned bx = 0; bx < grid.x; ++bx) emu_run_block(bx, by, bz, grid, block, body);
}
