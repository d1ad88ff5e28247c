// platform.cuh -- the one place that knows whether this translation unit is compiled by nvcc for
// sm_100a (the product) or by g++ with -DCKKS_EMU (tests/emu: a CPU *emulation of the CUDA
// execution model* that runs the very same kernel and host-orchestration source so the
// `-m "not gpu"` tests can check it against the oracle in a container without a GPU).
//
// The emulation build is test infrastructure: it is compiled only by tests/emu/build.py into
// tests/emu/, the product loader (desilofhe/_capi.py) never looks there, and bench.py /
// __graft_entry__.smoke() refuse to run on it.  There is no CPU fallback in the product.
//
// Kernel style that makes both builds possible without a fibre scheduler:
//   * 1-D thread blocks only;
//   * per-thread code sits inside FOR_THREADS { ... } regions; BLOCK_SYNC separates regions;
//   * nothing thread-dependent lives across a BLOCK_SYNC except in __shared__ memory
//     (block-uniform values may be declared outside the regions).
// Under nvcc FOR_THREADS is empty and BLOCK_SYNC is __syncthreads(); under CKKS_EMU FOR_THREADS is a
// loop over threadIdx.x and BLOCK_SYNC is nothing (regions run to completion one after another).
#pragma once
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <stdexcept>
#include <string>

typedef unsigned long long u64;
typedef long long i64;
typedef unsigned int u32;

extern long g_launch_count;   // kernels launched by this library (bench.py reports it as gpu_launches)

#ifndef CKKS_EMU
// =================================================================================== CUDA
#include <cuda_runtime.h>

#define FOR_THREADS
#define BLOCK_SYNC __syncthreads()
#ifdef CKKS_TIME_LAUNCHES
#include <time.h>
extern double g_launch_host_ns;
struct LaunchTimer {
    timespec t0;
    LaunchTimer() { clock_gettime(CLOCK_MONOTONIC, &t0); }
    ~LaunchTimer() { timespec t1; clock_gettime(CLOCK_MONOTONIC, &t1); g_launch_host_ns += (t1.tv_sec - t0.tv_sec) * 1e9 + (t1.tv_nsec - t0.tv_nsec); }
};
#define LAUNCH(kern, grid, block, stream, ...) do { LaunchTimer _lt; ++g_launch_count; kern<<<grid, block, 0, stream>>>(__VA_ARGS__); } while (0)
#else
#define LAUNCH(kern, grid, block, stream, ...) (++g_launch_count, kern<<<grid, block, 0, stream>>>(__VA_ARGS__))
#endif
#define CKKS_SHARED __shared__
#define GRID_CONST __grid_constant__
#define DEV_MEMBER __device__ __forceinline__      // member functions of device-side helper structs
// dynamic shared memory (kernels that need more than the 48 KB static limit): one 16-byte aligned array per kernel
#define DYN_SHARED_U64(name, words) extern __shared__ __align__(16) u64 name[]
template <typename K>
inline void dyn_smem_optin(K kern, size_t bytes) {
    // once per (kernel, device)
    static thread_local struct { const void* k; int dev; } done[16] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    for (auto& d : done) {
        if (d.k == (const void*)kern && d.dev == dev) return;
        if (!d.k) {
            if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess)
                throw std::runtime_error("cudaFuncSetAttribute(MaxDynamicSharedMemorySize) failed");
            d.k = (const void*)kern;
            d.dev = dev;
            return;
        }
    }
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}
#ifdef CKKS_TIME_LAUNCHES
#define LAUNCH_DYN(kern, grid, block, smem, stream, ...) do { LaunchTimer _lt; dyn_smem_optin(kern, smem); ++g_launch_count; kern<<<grid, block, smem, stream>>>(__VA_ARGS__); } while (0)
#else
#define LAUNCH_DYN(kern, grid, block, smem, stream, ...) (dyn_smem_optin(kern, smem), ++g_launch_count, kern<<<grid, block, smem, stream>>>(__VA_ARGS__))
#endif

#define CUDA_CHECK(expr)                                                                          \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess)                                                                    \
            throw std::runtime_error(std::string("CUDA error: ") + cudaGetErrorString(_e) +       \
                                     " at " #expr);                                               \
    } while (0)

typedef cudaStream_t dev_stream;

namespace dev {
// stream capture (CUDA graphs): while a capture is open nothing may synchronise or touch host memory; the guards below
// turn such a call into an exception BEFORE the driver sees it (an illegal call would silently invalidate the capture)
extern int g_capturing;
inline bool capturing() { return g_capturing != 0; }
inline void no_capture(const char* what) {
    if (g_capturing) throw std::runtime_error(std::string(what) + " is not allowed while a graph is being captured");
}
inline const char* backend_name() { return "cuda-sm_100a"; }
inline void set_device(int id) { CUDA_CHECK(cudaSetDevice(id)); }
inline dev_stream stream_create() {
    dev_stream s;
    CUDA_CHECK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    return s;
}
inline void stream_destroy(dev_stream s) { cudaStreamDestroy(s); }
inline void pool_setup(int device) {
    cudaMemPool_t pool;
    CUDA_CHECK(cudaDeviceGetDefaultMemPool(&pool, device));
    uint64_t keep = UINT64_MAX;   // never give memory back to the driver: the arena is reused
    CUDA_CHECK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
}
inline void* alloc(size_t bytes, dev_stream s) {
    void* p = nullptr;
    if (g_capturing) {      // a stream-ordered allocation would become a node of the graph: take plain device memory
        CUDA_CHECK(cudaMalloc(&p, bytes ? bytes : 8));
        return p;
    }
    CUDA_CHECK(cudaMallocAsync(&p, bytes ? bytes : 8, s));
    return p;
}
inline void free(void* p, dev_stream s) {
    no_capture("freeing device memory");
    if (p) cudaFreeAsync(p, s);
}
inline void h2d(void* d, const void* h, size_t bytes, dev_stream s) {
    no_capture("a host-to-device copy");
    CUDA_CHECK(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, s));
}
inline void d2h(void* h, const void* d, size_t bytes, dev_stream s) {
    no_capture("a device-to-host copy");
    CUDA_CHECK(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, s));
}
inline void d2d(void* o, const void* i, size_t bytes, dev_stream s) {
    CUDA_CHECK(cudaMemcpyAsync(o, i, bytes, cudaMemcpyDeviceToDevice, s));
}
inline void zero(void* d, size_t bytes, dev_stream s) { CUDA_CHECK(cudaMemsetAsync(d, 0, bytes, s)); }
inline void sync(dev_stream s) {
    no_capture("a stream synchronisation");
    CUDA_CHECK(cudaStreamSynchronize(s));
}
inline void check_last(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) throw std::runtime_error(std::string("CUDA launch error in ") + what + ": " + cudaGetErrorString(e));
}
// order `waiter` after everything enqueued on `src` so far (fork/join of the engine's lanes)
inline void stream_wait(dev_stream waiter, dev_stream src) {
    cudaEvent_t ev;
    CUDA_CHECK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    CUDA_CHECK(cudaEventRecord(ev, src));
    CUDA_CHECK(cudaStreamWaitEvent(waiter, ev, 0));
    CUDA_CHECK(cudaEventDestroy(ev));     // released once the wait has been satisfied
}
inline void* host_alloc_pinned(size_t bytes) {
    void* p = nullptr;
    CUDA_CHECK(cudaMallocHost(&p, bytes));
    return p;
}
inline void host_free_pinned(void* p) { cudaFreeHost(p); }
// a captured sequence of launches / copies (every stream that was forked from `s` during the capture included),
// instantiated once and replayed with one call
struct Graph {
    cudaGraph_t g = nullptr;
    cudaGraphExec_t exec = nullptr;
    size_t nodes = 0;
};
inline void capture_begin(dev_stream s) {
    if (g_capturing) throw std::runtime_error("a graph capture is already open");
    // relaxed mode: cudaMalloc (arena misses) and event creation stay legal on this thread during the capture
    CUDA_CHECK(cudaStreamBeginCapture(s, cudaStreamCaptureModeRelaxed));
    g_capturing = 1;
}
inline Graph* capture_end(dev_stream s) {
    g_capturing = 0;
    Graph* G = new Graph();
    cudaError_t e = cudaStreamEndCapture(s, &G->g);
    if (e != cudaSuccess || !G->g) {
        delete G;
        cudaGetLastError();
        throw std::runtime_error(std::string("graph capture failed: ") + cudaGetErrorString(e));
    }
    cudaGraphGetNodes(G->g, nullptr, &G->nodes);
    e = cudaGraphInstantiate(&G->exec, G->g, 0);
    if (e != cudaSuccess) {
        cudaGraphDestroy(G->g);
        delete G;
        throw std::runtime_error(std::string("graph instantiation failed: ") + cudaGetErrorString(e));
    }
    return G;
}
inline void capture_abort(dev_stream s) {
    if (!g_capturing) return;
    g_capturing = 0;
    cudaGraph_t g = nullptr;
    cudaStreamEndCapture(s, &g);
    if (g) cudaGraphDestroy(g);
    cudaGetLastError();
}
inline void graph_launch(Graph* G, dev_stream s) { CUDA_CHECK(cudaGraphLaunch(G->exec, s)); }
inline void graph_destroy(Graph* G) {
    if (!G) return;
    if (G->exec) cudaGraphExecDestroy(G->exec);
    if (G->g) cudaGraphDestroy(G->g);
    delete G;
}
inline void free_plain(void* p) { if (p) cudaFree(p); }      // memory taken by alloc() during a capture
struct Timer {
    cudaEvent_t a = nullptr, b = nullptr;
    void start(dev_stream s) {
        if (!a) { CUDA_CHECK(cudaEventCreate(&a)); CUDA_CHECK(cudaEventCreate(&b)); }
        CUDA_CHECK(cudaEventRecord(a, s));
    }
    float stop_ms(dev_stream s) {
        CUDA_CHECK(cudaEventRecord(b, s));
        CUDA_CHECK(cudaEventSynchronize(b));
        float ms = 0;
        CUDA_CHECK(cudaEventElapsedTime(&ms, a, b));
        return ms;
    }
    void mark_stop(dev_stream s) { CUDA_CHECK(cudaEventRecord(b, s)); }
    float elapsed_ms() {
        CUDA_CHECK(cudaEventSynchronize(b));
        float ms = 0;
        CUDA_CHECK(cudaEventElapsedTime(&ms, a, b));
        return ms;
    }
};
}  // namespace dev

__device__ __forceinline__ u64 mulhi64(u64 a, u64 b) { return __umul64hi(a, b); }
__device__ __forceinline__ double fmul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double fadd_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double fsub_rn(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double fdiv_rn(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ double ffma_rn(double a, double b, double c) { return __fma_rn(a, b, c); }
__device__ __forceinline__ double frint(double a) { return rint(a); }
__device__ __forceinline__ u32 brev32(u32 x) { return __brev(x); }
__device__ __forceinline__ int popc64(u64 x) { return __popcll(x); }
__device__ __forceinline__ i64 d2ll_rn(double x) { return __double2ll_rn(x); }
__device__ __forceinline__ double ull2d_rn(u64 x) { return __ull2double_rn(x); }
template <typename T>
__device__ __forceinline__ T ldg(const T* p) { return __ldg(p); }
// two consecutive 64-bit words (16-byte aligned) through the read-only path as one 128-bit load
__device__ __forceinline__ void ldg_pair(const u64* p, u64& a, u64& b) {
    const ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2*>(p));
    a = v.x;
    b = v.y;
}
// asynchronous global -> shared copies (LDGSTS): issued up front, no registers held while the data is in flight
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((u32)__cvta_generic_to_shared(smem_dst)), "l"(__cvta_generic_to_global(gmem_src)) : "memory");
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"((u32)__cvta_generic_to_shared(smem_dst)), "l"(__cvta_generic_to_global(gmem_src)) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
// bulk asynchronous copies (the TMA engine's 1-D form): ONE instruction moves `bytes` (a multiple of 16, both addresses 16-byte
// aligned) global -> shared and reports completion to an mbarrier in shared memory; no per-thread 8-byte requests, no
// registers, no LSU queue slots while the data is in flight
__device__ __forceinline__ void mbar_init(u64* bar, u32 count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((u32)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(u64* bar, u32 bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((u32)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, u32 bytes, u64* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     (u32)__cvta_generic_to_shared(smem_dst)),
                 "l"(__cvta_generic_to_global(gmem_src)), "r"(bytes), "r"((u32)__cvta_generic_to_shared(bar))
                 : "memory");
}
// shared -> global bulk store of `bytes` (multiple of 16, 16-byte aligned): issued by one thread, executed by the copy engine;
// bulk_store_wait_read returns when the shared-memory source has been read (the global writes complete asynchronously)
__device__ __forceinline__ void bulk_s2g(void* gmem_dst, const void* smem_src, u32 bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(__cvta_generic_to_global(gmem_dst)),
                 "r"((u32)__cvta_generic_to_shared(smem_src)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_store_commit_wait_read() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_wait(u64* bar, u32 phase) {
    asm volatile(
        "{\n .reg .pred p;\n MBAR_WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @!p bra MBAR_WAIT_%=;\n}\n" ::"r"(
            (u32)__cvta_generic_to_shared(bar)),
        "r"(phase)
        : "memory");
}
// pull a line towards the SM ahead of its use (no register, no shared memory held meanwhile)
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(__cvta_generic_to_global(p))); }
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(__cvta_generic_to_global(p))); }

#else
// =================================================================================== emulation (tests only)
#include <math.h>
#include <time.h>

struct emu_uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
extern thread_local emu_uint3 blockIdx, threadIdx;
extern thread_local dim3 blockDim, gridDim;

#define __global__ static
#define __device__ static
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define CKKS_SHARED static thread_local
#define GRID_CONST
#define DEV_MEMBER inline
#define DYN_SHARED_U64(name, words) static thread_local u64 name[words]
#define FOR_THREADS for (threadIdx.x = 0; threadIdx.x < blockDim.x; ++threadIdx.x)
#define BLOCK_SYNC ((void)0)

typedef int dev_stream;

#include <functional>
#include <vector>
namespace emu {
// graph capture under emulation: launches and device copies are RECORDED (not run) while a capture is open, exactly as
// CUDA does, and replayed in order by graph_launch -- so the host-side capture logic (static buffers, private arena,
// memoised level alignments) is testable without a GPU
extern std::vector<std::function<void()>>* g_record;
template <typename F>
inline void run_grid(dim3 grid, dim3 block, F body);
template <typename F>
inline void launch(dim3 grid, dim3 block, F body) {
    if (g_record) { g_record->push_back([=]() { run_grid(grid, block, body); }); return; }
    run_grid(grid, block, body);
}
template <typename F>
inline void run_grid(dim3 grid, dim3 block, F body) {
    const long total = (long)grid.x * grid.y * grid.z;
#pragma omp parallel for schedule(static)
    for (long b = 0; b < total; b++) {
        gridDim = grid;
        blockDim = block;
        blockIdx.x = (unsigned)(b % grid.x);
        blockIdx.y = (unsigned)((b / grid.x) % grid.y);
        blockIdx.z = (unsigned)(b / ((long)grid.x * grid.y));
        threadIdx.x = threadIdx.y = threadIdx.z = 0;
        body();
    }
}
}  // namespace emu
#define LAUNCH(kern, grid, block, stream, ...) (++g_launch_count, emu::launch(grid, block, [=]() { kern(__VA_ARGS__); }))
#define LAUNCH_DYN(kern, grid, block, smem, stream, ...) LAUNCH(kern, grid, block, stream, __VA_ARGS__)

namespace dev {
extern int g_capturing;
inline bool capturing() { return g_capturing != 0; }
inline void no_capture(const char* what) {
    if (g_capturing) throw std::runtime_error(std::string(what) + " is not allowed while a graph is being captured");
}
inline const char* backend_name() { return "emulation (tests only)"; }
inline void set_device(int) {}
inline dev_stream stream_create() { return 0; }
inline void stream_destroy(dev_stream) {}
inline void pool_setup(int) {}
inline void* alloc(size_t bytes, dev_stream) {
    void* p = ::malloc(bytes ? bytes : 8);
    if (!p) throw std::runtime_error("emulation: out of memory");
    return p;
}
inline void free(void* p, dev_stream) { no_capture("freeing device memory"); ::free(p); }
inline void h2d(void* d, const void* h, size_t bytes, dev_stream) { no_capture("a host-to-device copy"); memcpy(d, h, bytes); }
inline void d2h(void* h, const void* d, size_t bytes, dev_stream) { no_capture("a device-to-host copy"); memcpy(h, d, bytes); }
inline void d2d(void* o, const void* i, size_t bytes, dev_stream) {
    if (emu::g_record) { emu::g_record->push_back([=]() { memmove(o, i, bytes); }); return; }
    memmove(o, i, bytes);
}
inline void zero(void* d, size_t bytes, dev_stream) {
    if (emu::g_record) { emu::g_record->push_back([=]() { memset(d, 0, bytes); }); return; }
    memset(d, 0, bytes);
}
inline void sync(dev_stream) { no_capture("a stream synchronisation"); }
struct Graph {
    std::vector<std::function<void()>> ops;
    size_t nodes = 0;
};
inline void capture_begin(dev_stream) {
    if (g_capturing) throw std::runtime_error("a graph capture is already open");
    g_capturing = 1;
    emu::g_record = new std::vector<std::function<void()>>();
}
inline Graph* capture_end(dev_stream) {
    Graph* G = new Graph();
    G->ops.swap(*emu::g_record);
    G->nodes = G->ops.size();
    delete emu::g_record;
    emu::g_record = nullptr;
    g_capturing = 0;
    return G;
}
inline void capture_abort(dev_stream) {
    delete emu::g_record;
    emu::g_record = nullptr;
    g_capturing = 0;
}
inline void graph_launch(Graph* G, dev_stream) { for (auto& f : G->ops) f(); }
inline void graph_destroy(Graph* G) { delete G; }
inline void free_plain(void* p) { ::free(p); }
inline void check_last(const char*) {}
inline void stream_wait(dev_stream, dev_stream) {}
inline void* host_alloc_pinned(size_t bytes) { return ::malloc(bytes); }
inline void host_free_pinned(void* p) { ::free(p); }
struct Timer {
    struct timespec t0;
    void start(dev_stream) { clock_gettime(CLOCK_MONOTONIC, &t0); }
    struct timespec t1s;
    float stop_ms(dev_stream s) { mark_stop(s); return elapsed_ms(); }
    void mark_stop(dev_stream) { clock_gettime(CLOCK_MONOTONIC, &t1s); }
    float elapsed_ms() { return (float)((t1s.tv_sec - t0.tv_sec) * 1e3 + (t1s.tv_nsec - t0.tv_nsec) * 1e-6); }
};
}  // namespace dev

static inline u64 mulhi64(u64 a, u64 b) { return (u64)(((unsigned __int128)a * b) >> 64); }
static inline double fmul_rn(double a, double b) { return a * b; }   // built with -ffp-contract=off
static inline double fadd_rn(double a, double b) { return a + b; }
static inline double fsub_rn(double a, double b) { return a - b; }
static inline double fdiv_rn(double a, double b) { return a / b; }
static inline double ffma_rn(double a, double b, double c) { return fma(a, b, c); }     // exact single rounding (libm)
static inline double frint(double a) { return nearbyint(a); }
static inline u32 brev32(u32 x) {
    x = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
    x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
    x = ((x >> 4) & 0x0F0F0F0Fu) | ((x & 0x0F0F0F0Fu) << 4);
    x = ((x >> 8) & 0x00FF00FFu) | ((x & 0x00FF00FFu) << 8);
    return (x >> 16) | (x << 16);
}
static inline int popc64(u64 x) { return __builtin_popcountll(x); }
static inline i64 d2ll_rn(double x) { return (i64)nearbyint(x); }
static inline double ull2d_rn(u64 x) { return (double)x; }
template <typename T>
static inline T ldg(const T* p) { return *p; }
static inline void ldg_pair(const u64* p, u64& a, u64& b) { a = p[0]; b = p[1]; }
static inline void cp_async8(void* d, const void* s) { memcpy(d, s, 8); }
static inline void cp_async16(void* d, const void* s) { memcpy(d, s, 16); }
static inline void cp_async_wait_all() {}
static inline void mbar_init(u64*, u32) {}
static inline void mbar_expect_tx(u64*, u32) {}
static inline void bulk_g2s(void* d, const void* s, u32 bytes, u64*) { memcpy(d, s, bytes); }
static inline void mbar_wait(u64*, u32) {}
static inline void bulk_s2g(void* d, const void* s, u32 bytes) { memcpy(d, s, bytes); }
static inline void bulk_store_commit_wait_read() {}
static inline void fence_proxy_async() {}
static inline void prefetch_l1(const void*) {}
static inline void prefetch_l2(const void*) {}
#endif
