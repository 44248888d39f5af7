// emu_cluster.h — a thread-block cluster on the host: every CTA of the cluster runs in its OWN copy of the emulation
// library (the driver dlopen()s N copies of the .so, so each has its own `static` = __shared__ variables), all CTAs
// concurrently, with one barrier over all their threads for cluster.sync().  Distributed shared memory: an address
// inside copy A's image maps to the same offset inside copy B's image (dladdr base).  Include after emu_block.h.
#pragma once
#include <dlfcn.h>

struct EmuClusterCtx {
    std::barrier<>* bar;          // over the threads of all CTAs
    int nranks;
    uintptr_t base[16];           // load address of the library copy that runs CTA r
};
static thread_local EmuClusterCtx* emu_cluster = nullptr;
static thread_local unsigned emu_cluster_rank = 0;

namespace cooperative_groups {
struct cluster_group {
    unsigned block_rank() const { return emu_cluster_rank; }
    unsigned num_blocks() const { return (unsigned)emu_cluster->nranks; }
    void sync() const { emu_cluster->bar->arrive_and_wait(); }
    template <typename T> T* map_shared_rank(T* p, unsigned r) const
    {
        return reinterpret_cast<T*>(reinterpret_cast<uintptr_t>(p) - emu_cluster->base[emu_cluster_rank] + emu_cluster->base[r]);
    }
};
static inline cluster_group this_cluster() { return cluster_group(); }
}

static int emu_anchor;      // any object of this image
extern "C" uintptr_t emu_lib_base()
{
    Dl_info info;
    dladdr(&emu_anchor, &info);
    return reinterpret_cast<uintptr_t>(info.dli_fbase);
}
extern "C" void* emu_cluster_create(int nranks, int threads_per_cta)
{
    EmuClusterCtx* c = new EmuClusterCtx();
    c->bar = new std::barrier<>((std::ptrdiff_t)nranks * threads_per_cta);
    c->nranks = nranks;
    for (int i = 0; i < 16; ++i) c->base[i] = 0;
    return c;
}
extern "C" void emu_cluster_set_base(void* h, int rank, uintptr_t base) { static_cast<EmuClusterCtx*>(h)->base[rank] = base; }
extern "C" void emu_cluster_destroy(void* h) { EmuClusterCtx* c = static_cast<EmuClusterCtx*>(h); delete c->bar; delete c; }
