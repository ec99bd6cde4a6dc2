// Device-code portability layer.
//
// Product build (nvcc, sm_100a): everything maps 1:1 onto CUDA.
//
// Emulation build (g++ -DVB_EMULATE, TEST INFRASTRUCTURE ONLY -- see tests/emu/): the same
// kernel sources are compiled for the host and every CUDA thread of a block becomes one OS
// thread, __syncthreads() a std::barrier, warp collectives a 32-thread barrier plus a scratch
// row.  Blocks run one after another.  It exists so that the kernel *logic* can be checked
// against the oracle in the GPU-less build container (pytest -m "not gpu"); it is never
// loaded by the product (vent_analysis_b200/_lib.py only opens the CUDA library and raises
// if it is missing).
#pragma once
#include <stdint.h>
#include <stddef.h>

#if !defined(VB_EMULATE)
// ------------------------------------------------------------------------------------ CUDA
#include <cuda_runtime.h>
#define VB_SHARED __shared__
#define VB_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#define VB_KERNEL __global__
#define VB_DEV __device__ __forceinline__
#define VB_DEV_MEMBER __device__ __forceinline__
#define VB_DEV_NOINLINE __device__ __noinline__
#define VB_RESTRICT __restrict__
#define VB_LAUNCH(kernel, grid, block, smem, stream, ...) \
    kernel<<<(grid), (block), (smem), (cudaStream_t)(stream)>>>(__VA_ARGS__)
#define VB_LAUNCH_BOUNDS(t, b) __launch_bounds__(t, b)
// launch with thread-block clusters of `k` CTAs along x (the CTAs of a cluster are co-scheduled); false when
// the device cannot co-schedule one such cluster or the launch is refused (nothing was launched then)
template <class KernelT, class... Args>
inline bool vb_launch_cluster(KernelT kernel, dim3 grid, dim3 block, size_t smem, cudaStream_t stream, int k, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)k; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int fit = 0;
    if (cudaOccupancyMaxActiveClusters(&fit, kernel, &cfg) != cudaSuccess || fit < 1) { cudaGetLastError(); return false; }
    if (cudaLaunchKernelEx(&cfg, kernel, args...) != cudaSuccess) { cudaGetLastError(); return false; }
    return true;
}
#define VB_LAUNCH_CLUSTER(kernel, grid, block, smem, stream, k, ...) \
    vb_launch_cluster(kernel, (grid), (block), (smem), (cudaStream_t)(stream), (k), __VA_ARGS__)
VB_DEV void vb_sincospi(double x, double* s, double* c) { sincospi(x, s, c); }
VB_DEV float vb_ldg(const float* p) { return __ldg(p); }
VB_DEV uint32_t vb_ldg(const uint32_t* p) { return __ldg(p); }
VB_DEV uint8_t vb_ldg(const uint8_t* p) { return __ldg(p); }
VB_DEV int vb_ldg(const int* p) { return __ldg(p); }
VB_DEV int2 vb_ldg(const int2* p) { return __ldg(p); }
VB_DEV float4 vb_ldg(const float4* p) { return __ldg(p); }
// loads that bypass L1 (served by L2): data another CTA of the same kernel may have written
VB_DEV float vb_ld_cg(const float* p) { return __ldcg(p); }
VB_DEV float4 vb_ld_cg(const float4* p) { return __ldcg(p); }
VB_DEV unsigned vb_ld_cg(const unsigned* p) { return __ldcg(p); }
VB_DEV int vb_ld_cg(const int* p) { return __ldcg(p); }
VB_DEV unsigned long long vb_ld_cg(const unsigned long long* p) { return __ldcg(p); }
VB_DEV double vb_ld_cg(const double* p) { return __ldcg(p); }
VB_DEV unsigned vb_ld_acquire(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// streaming (read-once) loads: keep them out of L1
VB_DEV float vb_ld_stream(const float* p) { return __ldcs(p); }
VB_DEV uint8_t vb_ld_stream(const uint8_t* p) { return __ldcs(p); }
VB_DEV uint32_t vb_ld_stream(const uint32_t* p) { return __ldcs(p); }
VB_DEV float4 vb_ld_stream(const float4* p) { return __ldcs(p); }
VB_DEV void vb_st_stream(float* p, float v) { __stcs(p, v); }
// named barriers (ids 1..15) over a warp-aligned subset of the block
VB_DEV void vb_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
VB_DEV int vb_bar_or(int id, int nthreads, int pred) {
    int out;
    asm volatile(
        "{\n\t.reg .pred p, q;\n\tsetp.ne.s32 q, %3, 0;\n\tbar.red.or.pred p, %1, %2, q;\n\tselp.s32 %0, 1, 0, p;\n\t}"
        : "=r"(out) : "r"(id), "r"(nthreads), "r"(pred) : "memory");
    return out;
}
VB_DEV long long vb_clock() { return clock64(); }   // SM cycles
VB_DEV unsigned vb_lanemask_lt() { unsigned m; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m)); return m; }   // lanes below this one
#define VB_OPAQUE_PTR(p) asm volatile("" : "+l"(p))
// ---- explicit shared-memory accesses through a 32-bit shared-window address ---------------------
// A pointer into shared memory that has been made opaque (VB_OPAQUE_PTR) is a GENERIC pointer to the
// compiler: every access through it is a generic LD / ST with 64-bit address arithmetic.  Hot loops
// take the 32-bit shared address once (opaque as well, so that it is not re-derived from the kernel
// parameters at each use) and go through LDS / STS with 32-bit address arithmetic instead.
typedef unsigned vb_saddr;
VB_DEV vb_saddr vb_saddr_of(const void* p) {
    unsigned a = (unsigned)__cvta_generic_to_shared(p);
    asm volatile("" : "+r"(a));
    return a;
}
VB_DEV float vb_lds_f32(vb_saddr a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }
VB_DEV int vb_lds_s32(vb_saddr a) { int v; asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
VB_DEV float4 vb_lds_f32x4(vb_saddr a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
VB_DEV void vb_sts_f32(vb_saddr a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
VB_DEV unsigned vb_atoms_add_u32(vb_saddr a, unsigned v) {        // ATOMS: returns the old value
    unsigned old;
    asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(a), "r"(v) : "memory");
    return old;
}
VB_DEV void vb_reds_add_u32(vb_saddr a, unsigned v) { asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
// ---- bulk asynchronous copies (TMA, 1-D form) with mbarrier completion --------------------------
// global -> shared through the async proxy: one thread arms the barrier with the byte count and
// issues the copy; the bytes land without occupying registers or issue slots, and every consumer
// waits on the barrier's phase.  Source, destination and size are multiples of 16 bytes.
typedef unsigned long long vb_mbar_t;
VB_DEV unsigned vb_smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
VB_DEV void vb_mbar_init(vb_mbar_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(vb_smem_addr(bar)), "r"(count) : "memory");
}
VB_DEV void vb_mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
VB_DEV void vb_mbar_expect_tx(vb_mbar_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(vb_smem_addr(bar)), "r"(bytes) : "memory");
}
VB_DEV void vb_bulk_g2s(void* dst_smem, const void* src_global, unsigned bytes, vb_mbar_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(vb_smem_addr(dst_smem)), "l"(src_global), "r"(bytes), "r"(vb_smem_addr(bar)) : "memory");
}
VB_DEV void vb_mbar_wait(vb_mbar_t* bar, unsigned parity);
// ring-of-stages helpers shared with the emulation: fill number `k` (0, 1, 2, ...) of a stage
VB_DEV void vb_stage_begin(vb_mbar_t* bar, unsigned bytes) { vb_mbar_expect_tx(bar, bytes); }
VB_DEV void vb_stage_end(vb_mbar_t*) {}
VB_DEV void vb_stage_wait(vb_mbar_t* bar, unsigned k) { vb_mbar_wait(bar, k & 1u); }
// "stage is free again" barrier: `consumers` arrivals complete a phase; the producer waits for
// the consumers of fill k - 1 before it starts fill k (nothing to wait for before fill 0)
VB_DEV void vb_free_arrive(vb_mbar_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(vb_smem_addr(bar)) : "memory");
}
VB_DEV void vb_free_wait(vb_mbar_t* bar, unsigned k, int) { if (k > 0u) vb_mbar_wait(bar, (k - 1u) & 1u); }
VB_DEV void vb_mbar_wait(vb_mbar_t* bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tVB_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra VB_DONE_%=;\n\tbra VB_WAIT_%=;\n\tVB_DONE_%=:\n\t}"
        ::"r"(vb_smem_addr(bar)), "r"(parity) : "memory");
}
#else
// ------------------------------------------------------------------------------- emulation
#include <math.h>
#include <sched.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/wait.h>
#include <unistd.h>
#include <algorithm>
#include <atomic>
#include <barrier>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#define VB_SHARED static
#define VB_DYN_SMEM(name) unsigned char* name = vbemu::g_block->dyn_smem
#define VB_KERNEL static
#define VB_DEV static inline
#define VB_DEV_MEMBER inline
#define VB_DEV_NOINLINE static
#define VB_RESTRICT __restrict__
#define VB_LAUNCH_BOUNDS(t, b)
#define __forceinline__ inline

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
struct double2 { double x, y; };
struct int2 { int x, y; };
struct uint2 { unsigned x, y; };
struct uint4 { unsigned x, y, z, w; };
static inline float4 make_float4(float a, float b, float c, float d) { return float4{a, b, c, d}; }
static inline float2 make_float2(float a, float b) { return float2{a, b}; }
static inline uint2 make_uint2(unsigned a, unsigned b) { return uint2{a, b}; }
static inline double2 make_double2(double a, double b) { return double2{a, b}; }
static inline int2 make_int2(int a, int b) { return int2{a, b}; }

namespace vbemu {
struct Idx { unsigned x, y, z; };
struct Warp {
    std::barrier<> bar;
    uint64_t scratch[32];
    explicit Warp(int n) : bar(n) {}
};
struct Block {
    std::barrier<> bar;
    std::vector<std::unique_ptr<Warp>> warps;
    unsigned char* dyn_smem;
    Block(int nthreads, size_t smem) : bar(nthreads) {
        for (int w = 0; w < (nthreads + 31) / 32; ++w)
            warps.emplace_back(new Warp(std::min(32, nthreads - 32 * w)));
        dyn_smem = (unsigned char*)aligned_alloc(128, (smem + 127) / 128 * 128 + 128);
        memset(dyn_smem, 0xCD, smem);   // poison: smem is NOT zero-initialised on a GPU
    }
    ~Block() { free(dyn_smem); }
};
inline thread_local Idx t_threadIdx, t_blockIdx;
inline Idx g_blockDim, g_gridDim;
inline Block* g_block = nullptr;
inline uint64_t g_launches = 0;

inline Warp& warp() { return *g_block->warps[t_threadIdx.x >> 5]; }
inline int lane() { return (int)(t_threadIdx.x & 31); }

template <class T> inline uint64_t pack(T v) { uint64_t u = 0; memcpy(&u, &v, sizeof(T)); return u; }
template <class T> inline T unpack(uint64_t u) { T v; memcpy(&v, &u, sizeof(T)); return v; }

// every lane publishes a value, then reads whatever it needs
template <class T, class F> inline auto collective(T v, F reader) {
    Warp& W = warp();
    W.scratch[lane()] = pack(v);
    W.bar.arrive_and_wait();
    auto r = reader(W.scratch);
    W.bar.arrive_and_wait();
    return r;
}

// one CTA: one OS thread per CUDA thread
template <class Body> void run_block(unsigned bx, unsigned by, unsigned bz, dim3 block, size_t smem, Body& body) {
    const int nthreads = (int)(block.x * block.y * block.z);
    Block blk(nthreads, smem);
    g_block = &blk;
    std::vector<std::thread> th;
    th.reserve(nthreads);
    for (int t = 0; t < nthreads; ++t)
        th.emplace_back([&, t] {
            t_threadIdx = Idx{(unsigned)t % block.x, ((unsigned)t / block.x) % block.y,
                              (unsigned)t / (block.x * block.y)};
            t_blockIdx = Idx{bx, by, bz};
            body();
        });
    for (auto& x : th) x.join();
    g_block = nullptr;
}
template <class Body> void launch(dim3 grid, dim3 block, size_t smem, Body body) {
    ++g_launches;
    g_gridDim = Idx{grid.x, grid.y, grid.z};
    g_blockDim = Idx{block.x, block.y, block.z};
    for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
            for (unsigned bx = 0; bx < grid.x; ++bx) run_block(bx, by, bz, block, smem, body);
}
// Thread-block clusters: the k CTAs of a cluster run CONCURRENTLY (they synchronise with each other
// through global memory), one forked process per CTA -- the function-local `static` arrays that stand
// for __shared__ memory are then private to a CTA, while device memory (vbrt::dev_malloc: shared
// anonymous mappings) is common to all of them.  Clusters run one after the other.
// Under AddressSanitizer device memory stays on the instrumented heap (out-of-bounds accesses of every kernel are
// caught there), which forked CTAs cannot share: a cluster launch is refused and the caller falls back.
#if defined(__SANITIZE_ADDRESS__)
#define VB_EMU_SHARED_DEVMEM 0
#else
#define VB_EMU_SHARED_DEVMEM 1
#endif
inline bool g_cluster_failed = false;
template <class Body> bool launch_cluster(dim3 grid, dim3 block, size_t smem, int k, Body body) {
    if (!VB_EMU_SHARED_DEVMEM) return false;
    ++g_launches;
    g_gridDim = Idx{grid.x, grid.y, grid.z};
    g_blockDim = Idx{block.x, block.y, block.z};
    for (unsigned c0 = 0; c0 < grid.x; c0 += (unsigned)k) {
        std::vector<pid_t> kids;
        for (unsigned r = 0; r < (unsigned)k && c0 + r < grid.x; ++r) {
            const pid_t pid = fork();
            if (pid == 0) {
                alarm(240);                                // a deadlocked cluster must not hang the test run
                run_block(c0 + r, 0, 0, block, smem, body);
                _exit(0);
            }
            if (pid < 0) { g_cluster_failed = true; break; }
            kids.push_back(pid);
        }
        for (pid_t pid : kids) {
            int st = 0;
            if (waitpid(pid, &st, 0) < 0 || !WIFEXITED(st) || WEXITSTATUS(st) != 0) g_cluster_failed = true;
        }
    }
    return true;                                           // (a CTA that died is reported by vbrt::last_launch_ok)
}
}  // namespace vbemu

#define threadIdx vbemu::t_threadIdx
#define blockIdx vbemu::t_blockIdx
#define blockDim vbemu::g_blockDim
#define gridDim vbemu::g_gridDim
#define VB_LAUNCH(kernel, grid, block, smem, stream, ...) \
    vbemu::launch((grid), (block), (smem), [=] { kernel(__VA_ARGS__); })
#define VB_LAUNCH_CLUSTER(kernel, grid, block, smem, stream, k, ...) \
    vbemu::launch_cluster((grid), (block), (smem), (k), [=] { kernel(__VA_ARGS__); })

static inline void __syncthreads() { vbemu::g_block->bar.arrive_and_wait(); }
static inline void __syncwarp(unsigned = 0xffffffffu) {
    vbemu::warp().bar.arrive_and_wait();
}
static inline int __syncthreads_or(int pred) {
    static int flag[2];
    static int phase;
    vbemu::g_block->bar.arrive_and_wait();
    const int ph = phase;
    if (pred) __atomic_store_n(&flag[ph], 1, __ATOMIC_SEQ_CST);
    vbemu::g_block->bar.arrive_and_wait();
    const int r = __atomic_load_n(&flag[ph], __ATOMIC_SEQ_CST);
    vbemu::g_block->bar.arrive_and_wait();
    if (threadIdx.x == 0) { flag[ph] = 0; phase = ph ^ 1; }
    vbemu::g_block->bar.arrive_and_wait();
    return r;
}
// named barriers: one std::barrier per id, created by the first thread that needs it (and again when a
// later use of the id has another thread count, e.g. the next launch splits its teams differently; the
// replaced barrier is kept alive: a thread may still be on its way out of it)
namespace vbemu {
struct NamedBar {
    std::barrier<> bar;
    int count;
    NamedBar(int n) : bar(n), count(n) {}
};
struct Named {
    std::mutex mu;
    std::atomic<NamedBar*> cur[16];
    std::vector<std::unique_ptr<NamedBar>> all;
    int flag[16][2];
    int phase[16];
};
inline Named g_named;
inline std::barrier<>& named(int id, int n) {
    NamedBar* q = g_named.cur[id].load(std::memory_order_acquire);
    if (q && q->count == n) return q->bar;
    std::lock_guard<std::mutex> g(g_named.mu);
    q = g_named.cur[id].load(std::memory_order_acquire);
    if (!q || q->count != n) {
        g_named.all.emplace_back(new NamedBar(n));
        q = g_named.all.back().get();
        g_named.cur[id].store(q, std::memory_order_release);
    }
    return q->bar;
}
}  // namespace vbemu
static inline void vb_bar_sync(int id, int nthreads) { vbemu::named(id, nthreads).arrive_and_wait(); }
static inline int vb_bar_or(int id, int nthreads, int pred) {
    std::barrier<>& b = vbemu::named(id, nthreads);
    b.arrive_and_wait();                                   // the flag is clear
    if (pred) __atomic_store_n(&vbemu::g_named.flag[id][0], 1, __ATOMIC_SEQ_CST);
    b.arrive_and_wait();
    const int r = __atomic_load_n(&vbemu::g_named.flag[id][0], __ATOMIC_SEQ_CST);
    b.arrive_and_wait();                                   // everybody has read it
    __atomic_store_n(&vbemu::g_named.flag[id][0], 0, __ATOMIC_SEQ_CST);
    return r;
}
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline void __threadfence_block() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }

static inline unsigned __ballot_sync(unsigned, int pred) {
    return vbemu::collective<uint32_t>(pred ? 1u : 0u, [](uint64_t* s) {
        unsigned r = 0;
        const int n = std::min(32u, blockDim.x - (threadIdx.x & ~31u));
        for (int l = 0; l < n; ++l) if (s[l]) r |= 1u << l;
        return r;
    });
}
static inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
static inline int __all_sync(unsigned m, int pred) {
    const int n = std::min(32u, blockDim.x - (threadIdx.x & ~31u));
    return __ballot_sync(m, pred) == (n == 32 ? 0xffffffffu : ((1u << n) - 1));
}
template <class T> static inline T __shfl_sync(unsigned, T v, int src) {
    return vbemu::collective<T>(v, [src](uint64_t* s) { return vbemu::unpack<T>(s[src & 31]); });
}
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int m) {
    const int me = vbemu::lane();
    return vbemu::collective<T>(v, [me, m](uint64_t* s) { return vbemu::unpack<T>(s[(me ^ m) & 31]); });
}
template <class T> static inline T __shfl_down_sync(unsigned, T v, int d) {
    const int me = vbemu::lane();
    return vbemu::collective<T>(v, [me, d](uint64_t* s) { return vbemu::unpack<T>(s[me + d < 32 ? me + d : me]); });
}
template <class T> static inline T __shfl_up_sync(unsigned, T v, int d) {
    const int me = vbemu::lane();
    return vbemu::collective<T>(v, [me, d](uint64_t* s) { return vbemu::unpack<T>(s[me - d >= 0 ? me - d : me]); });
}
static inline unsigned __reduce_or_sync(unsigned, unsigned v) {
    return vbemu::collective<uint32_t>(v, [](uint64_t* s) {
        unsigned r = 0;
        const int n = std::min(32u, blockDim.x - (threadIdx.x & ~31u));
        for (int l = 0; l < n; ++l) r |= (unsigned)s[l];
        return r;
    });
}
static inline unsigned __reduce_add_sync(unsigned, unsigned v) {
    return vbemu::collective<uint32_t>(v, [](uint64_t* s) {
        unsigned r = 0;
        const int n = std::min(32u, blockDim.x - (threadIdx.x & ~31u));
        for (int l = 0; l < n; ++l) r += (unsigned)s[l];
        return r;
    });
}
static inline unsigned __reduce_min_sync(unsigned, unsigned v) {
    return vbemu::collective<uint32_t>(v, [](uint64_t* s) {
        unsigned r = 0xffffffffu;
        const int n = std::min(32u, blockDim.x - (threadIdx.x & ~31u));
        for (int l = 0; l < n; ++l) r = std::min(r, (unsigned)s[l]);
        return r;
    });
}
static inline unsigned __reduce_max_sync(unsigned, unsigned v) {
    return vbemu::collective<uint32_t>(v, [](uint64_t* s) {
        unsigned r = 0;
        const int n = std::min(32u, blockDim.x - (threadIdx.x & ~31u));
        for (int l = 0; l < n; ++l) r = std::max(r, (unsigned)s[l]);
        return r;
    });
}

// atomics
static inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) {
    return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST);
}
static inline float atomicAdd(float* p, float v) {
    uint32_t* u = (uint32_t*)p;
    uint32_t old = __atomic_load_n(u, __ATOMIC_SEQ_CST), nw;
    do {
        float f; memcpy(&f, &old, 4); f += v; memcpy(&nw, &f, 4);
    } while (!__atomic_compare_exchange_n(u, &old, nw, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST));
    float r; memcpy(&r, &old, 4); return r;
}
static inline double atomicAdd(double* p, double v) {
    uint64_t* u = (uint64_t*)p;
    uint64_t old = __atomic_load_n(u, __ATOMIC_SEQ_CST), nw;
    do {
        double f; memcpy(&f, &old, 8); f += v; memcpy(&nw, &f, 8);
    } while (!__atomic_compare_exchange_n(u, &old, nw, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST));
    double r; memcpy(&r, &old, 8); return r;
}
static inline unsigned atomicOr(unsigned* p, unsigned v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
static inline int atomicMax(int* p, int v) {
    int old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return old;
}
static inline int atomicMin(int* p, int v) {
    int old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (old > v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return old;
}
static inline unsigned atomicMax(unsigned* p, unsigned v) {
    unsigned old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return old;
}
static inline unsigned atomicMin(unsigned* p, unsigned v) {
    unsigned old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (old > v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return old;
}
static inline unsigned atomicExch(unsigned* p, unsigned v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }

// bit / math intrinsics
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
static inline unsigned __brev(unsigned v) {
    unsigned r = 0;
    for (int i = 0; i < 32; ++i) r |= ((v >> i) & 1u) << (31 - i);
    return r;
}
static inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
static inline int __float_as_int(float f) { int u; memcpy(&u, &f, 4); return u; }
static inline float __int_as_float(int u) { float f; memcpy(&f, &u, 4); return f; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline float __frcp_rn(float a) { return 1.0f / a; }
static inline unsigned __float2uint_rn(float a) { return (unsigned)llrintf(a); }
static inline long long __float2ll_rn(float a) { return llrintf(a); }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline double __fma_rn(double a, double b, double c) { return fma(a, b, c); }
static inline void vb_sincospi(double x, double* s, double* c) { *s = sin(M_PI * x); *c = cos(M_PI * x); }
template <class T> static inline T vb_ldg(const T* p) { return *p; }
template <class T> static inline T vb_ld_cg(const T* p) { return *p; }
static inline unsigned vb_ld_acquire(const unsigned* p) { usleep(20); return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
template <class T> static inline T vb_ld_stream(const T* p) { return *p; }
static inline void vb_st_stream(float* p, float v) { *p = v; }
static inline long long vb_clock() { return 0; }
static inline unsigned vb_lanemask_lt() { return (1u << (threadIdx.x & 31u)) - 1u; }
#define VB_OPAQUE_PTR(p) ((void)0)
typedef uintptr_t vb_saddr;                  // emulation: the "shared address" is the host address
static inline vb_saddr vb_saddr_of(const void* p) { return (vb_saddr)p; }
static inline float vb_lds_f32(vb_saddr a) { return *(const float*)a; }
static inline int vb_lds_s32(vb_saddr a) { return *(const int*)a; }
static inline float4 vb_lds_f32x4(vb_saddr a) { return *(const float4*)a; }
static inline void vb_sts_f32(vb_saddr a, float v) { *(float*)a = v; }
static inline unsigned vb_atoms_add_u32(vb_saddr a, unsigned v) { return __atomic_fetch_add((unsigned*)a, v, __ATOMIC_SEQ_CST); }
static inline void vb_reds_add_u32(vb_saddr a, unsigned v) { __atomic_fetch_add((unsigned*)a, v, __ATOMIC_SEQ_CST); }
// bulk copies: the issuing thread copies synchronously and bumps a completion counter; a phase
// `parity` of the barrier is complete once the counter has passed it (phases complete in order)
typedef unsigned long long vb_mbar_t;
static inline void vb_mbar_init(vb_mbar_t* bar, int) { __atomic_store_n(bar, 0ull, __ATOMIC_SEQ_CST); }
static inline void vb_mbar_fence_init() {}
// emulation: (pending copies << 32) | completed phases
static inline void vb_mbar_expect_tx(vb_mbar_t*, unsigned) {}
static inline void vb_bulk_g2s(void* dst, const void* src, unsigned bytes, vb_mbar_t*) { memcpy(dst, src, bytes); }
static inline void vb_mbar_complete_emu(vb_mbar_t* bar) { __atomic_fetch_add(bar, 1ull, __ATOMIC_SEQ_CST); }
static inline void vb_mbar_wait_emu(vb_mbar_t* bar, unsigned long long phases_done) {
    while (__atomic_load_n(bar, __ATOMIC_SEQ_CST) < phases_done) std::this_thread::yield();
}
static inline void vb_stage_begin(vb_mbar_t*, unsigned) {}
static inline void vb_stage_end(vb_mbar_t* bar) { vb_mbar_complete_emu(bar); }
static inline void vb_stage_wait(vb_mbar_t* bar, unsigned k) { vb_mbar_wait_emu(bar, (unsigned long long)k + 1ull); }
static inline void vb_free_arrive(vb_mbar_t* bar) { vb_mbar_complete_emu(bar); }
static inline void vb_free_wait(vb_mbar_t* bar, unsigned k, int consumers) { vb_mbar_wait_emu(bar, (unsigned long long)k * (unsigned long long)consumers); }
using std::max;
using std::min;
#endif
