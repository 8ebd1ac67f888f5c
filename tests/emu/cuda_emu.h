// cuda_emu.h — TEST INFRASTRUCTURE ONLY: a host-side lockstep emulator of the CUDA execution model, just wide
// enough to compile torchdriveenv_b200/csrc/*.cu{,h} with g++ and run the very same kernel source on the CPU.
//
// Why it exists: this repository is developed in a container without a GPU.  The kernels are warp-cooperative
// (shuffles, ballots, shared-memory atomics, named barriers, bulk-async copies), so their *logic* can be checked
// against the CPU oracle before a GPU lease is spent on them.  tests/emu/build.sh compiles csrc/tde_b200.cu with
// -DTDE_HOST_EMU into tests/emu/libtde_emu.so (same C ABI, host pointers instead of device pointers); the tests under
// tests/test_emu_*.py drive it through tests/emu_engine.py.  NOTHING under torchdriveenv_b200/ loads, links or
// mentions this library: the product has no CPU path (DESIGN.md §1 "No fallback").  It proves nothing about speed,
// races or PTX semantics - the `-m gpu` tests on a B200 remain the parity tests proper.
//
// Model: one kernel launch = blocks run one after the other; the threads of a block are fibers (hand-rolled x86-64
// context switch) scheduled round-robin on ONE OS thread, so shared memory is a plain static buffer, atomics are
// plain read-modify-writes and every run is deterministic.  A warp collective (shfl / ballot / any / redux /
// syncwarp) is a rendezvous of the lanes named in its mask: a lane deposits its value and yields until all have
// arrived.  __syncthreads and named barriers rendezvous the block (exited threads count as arrived).  A sweep over
// all fibers without progress is reported as a deadlock (e.g. a collective whose mask names a lane that diverged).
#pragma once
#ifndef TDE_HOST_EMU
#error "cuda_emu.h is only for -DTDE_HOST_EMU builds (tests/emu)"
#endif
#if !defined(__x86_64__)
#error "the fiber switch of cuda_emu.h is written for x86-64"
#endif

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <vector>

// ---------------------------------------------------------------- vector types (layout-compatible with CUDA's)
struct alignas(8) float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(8) int2 { int x, y; };
struct alignas(16) int4 { int x, y, z, w; };
struct alignas(8) uint2 { unsigned x, y; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
struct uint3 { unsigned x, y, z; };
struct alignas(16) ulonglong2 { unsigned long long x, y; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline int2 make_int2(int x, int y) { return int2{x, y}; }
static inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
static inline ulonglong2 make_ulonglong2(unsigned long long x, unsigned long long y) { return ulonglong2{x, y}; }

// ---------------------------------------------------------------- qualifiers
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)
#define __restrict__
#define __align__(n) alignas(n)
#define __shared__ static   // one block runs at a time on one OS thread: a static buffer IS the block's shared memory

namespace emu {

// ---------------------------------------------------------------- fibers
extern "C" void emu_switch(void** save_sp, void* new_sp);
__asm__(
    ".text\n.globl emu_switch\n.type emu_switch,@function\nemu_switch:\n"
    "  pushq %rbp\n  pushq %rbx\n  pushq %r12\n  pushq %r13\n  pushq %r14\n  pushq %r15\n"
    "  movq %rsp, (%rdi)\n  movq %rsi, %rsp\n"
    "  popq %r15\n  popq %r14\n  popq %r13\n  popq %r12\n  popq %rbx\n  popq %rbp\n  ret\n"
    ".size emu_switch,.-emu_switch\n");

struct Cell {            // rendezvous of the lanes of one mask (or of a block barrier)
    unsigned arrived = 0, toread = 0;
    bool drain = false;
    unsigned long long vals[32];
};
struct Warp {
    std::map<unsigned, Cell> cells;
};
struct BlockBarrier { int arrived = 0, generation = 0; };
struct Fiber {
    void* sp = nullptr;
    char* stack = nullptr;
    bool done = true;
    uint3 tid;
    // what the fiber is waiting for (checked by the scheduler, so that a blocked fiber costs no context switch)
    int wkind = 0;            // 0 runnable, 1 cell no longer draining, 2 cell ready for this lane, 3 block barrier, 4 memory word
    const volatile unsigned long long* wword = nullptr;   // kind 4: runnable once (*wword >> 32 & 1) != wparity
    unsigned wparity = 0;
    Cell* wcell = nullptr;
    unsigned wbit = 0;
    BlockBarrier* wbar = nullptr;
    int wgen = 0, wcount = 0, wid = 0;
};

struct Runtime {
    std::vector<Fiber> fibers;
    std::vector<Warp> warps;
    std::map<int, BlockBarrier> barriers;   // key -1 = __syncthreads, >= 0 named barriers
    void* sched_sp = nullptr;
    int current = -1, alive = 0, nthreads = 0;
    unsigned long long progress = 0;
    void (*body)(void*) = nullptr;
    void* body_arg = nullptr;
    std::vector<unsigned char> dyn_smem;
    size_t stack_bytes = 256 * 1024;
};
inline Runtime& rt() { static Runtime r; return r; }

}  // namespace emu

// the CUDA built-ins: rewritten by the scheduler every time a fiber is resumed
inline uint3 threadIdx, blockIdx;
inline dim3 blockDim, gridDim;

namespace emu {

inline void yield() {
    Runtime& r = rt();
    Fiber& f = r.fibers[r.current];
    emu_switch(&f.sp, r.sched_sp);
}
inline bool runnable(Runtime& r, Fiber& f) {
    switch (f.wkind) {
        case 1: return !f.wcell->drain;
        case 2: return f.wcell->drain && (f.wcell->toread & f.wbit);
        case 3: return f.wbar->generation != f.wgen || f.wbar->arrived >= (f.wid < 0 ? r.alive : f.wcount);
        case 4: return (((unsigned)(*f.wword >> 32)) & 1u) != f.wparity;
        default: return true;
    }
}
inline void fiber_entry() {
    Runtime& r = rt();
    r.body(r.body_arg);
    Fiber& f = r.fibers[r.current];
    f.done = true;
    r.alive--;
    r.progress++;
    emu_switch(&f.sp, r.sched_sp);
    std::abort();   // a finished fiber is never resumed
}
inline void prepare(Fiber& f, size_t stack_bytes) {
    if (!f.stack) f.stack = (char*)std::aligned_alloc(64, stack_bytes);
    uintptr_t top = ((uintptr_t)f.stack + stack_bytes) & ~(uintptr_t)15;
    void** s = (void**)top;
    *--s = nullptr;                 // fake return address of fiber_entry (never used)
    *--s = (void*)&fiber_entry;     // popped by the `ret` of emu_switch
    for (int k = 0; k < 6; ++k) *--s = nullptr;   // rbp rbx r12 r13 r14 r15
    f.sp = (void*)s;
    f.done = false;
}

[[noreturn]] inline void die(const char* what) {
    Runtime& r = rt();
    std::fprintf(stderr, "cuda_emu: %s (block %u,%u thread %d of %d)\n", what, blockIdx.x, blockIdx.y, r.current, r.nthreads);
    std::abort();
}

// run one block to completion
inline void run_block(dim3 bid, dim3 bdim, dim3 gdim) {
    Runtime& r = rt();
    const int n = (int)(bdim.x * bdim.y * bdim.z);
    r.nthreads = n;
    if ((int)r.fibers.size() < n) r.fibers.resize(n);
    r.warps.assign((n + 31) / 32, Warp());
    r.barriers.clear();
    for (int t = 0; t < n; ++t) {
        prepare(r.fibers[t], r.stack_bytes);
        r.fibers[t].tid = uint3{(unsigned)t % bdim.x, ((unsigned)t / bdim.x) % bdim.y, (unsigned)t / (bdim.x * bdim.y)};
    }
    r.alive = n;
    blockIdx = uint3{bid.x, bid.y, bid.z};
    blockDim = bdim; gridDim = gdim;
    // Scheduling policy (TDE_EMU_SCHED): 0 = round robin over all threads; 1 = the lowest-numbered warp that can make
    // progress always runs first; 2 = the highest-numbered one.  1 and 2 let one warp run as far ahead of the others as the
    // barriers allow, which exposes inter-warp ordering assumptions (a warp reading what another has already overwritten).
    static const int policy = std::getenv("TDE_EMU_SCHED") ? std::atoi(std::getenv("TDE_EMU_SCHED")) : 0;
    const int nw = (n + 31) / 32;
    while (r.alive > 0) {
        bool ran_any = false;   // every wait names its condition (Fiber::wkind), so a thread that is resumed does make progress
        for (int wi = 0; wi < nw; ++wi) {
            const int w = policy == 2 ? nw - 1 - wi : wi;
            bool ran = false;
            for (int t = w * 32; t < std::min(n, w * 32 + 32); ++t) {
                Fiber& f = r.fibers[t];
                if (f.done || !runnable(r, f)) continue;
                f.wkind = 0;
                r.current = t;
                threadIdx = f.tid;
                emu_switch(&r.sched_sp, f.sp);
                ran = true;
            }
            ran_any = ran_any || ran;
            if (ran && policy != 0) break;   // start over from the preferred end
        }
        if (r.alive > 0 && !ran_any) die("deadlock: no thread of the block can make progress (divergent collective or barrier?)");
    }
    r.current = -1;
}

template <class F>
inline void run_grid(dim3 g, dim3 b, size_t smem, F&& f) {
    Runtime& r = rt();
    if (r.current >= 0) die("nested kernel launch");
    r.dyn_smem.assign(smem + 64, 0);
    auto tramp = [](void* p) { (*static_cast<F*>(p))(); };
    r.body = tramp; r.body_arg = (void*)&f;
    for (unsigned z = 0; z < g.z; ++z)
        for (unsigned y = 0; y < g.y; ++y)
            for (unsigned x = 0; x < g.x; ++x) run_block(dim3(x, y, z), b, g);
}
template <class K>
struct Launcher {
    dim3 g, b; size_t smem; K k;
    template <class... A> void operator()(A... a) { run_grid(g, b, smem, [&] { k(a...); }); }
};
template <class K>
inline Launcher<K> make_launcher(dim3 g, dim3 b, size_t smem, K k) { return Launcher<K>{g, b, smem, k}; }

inline unsigned char* dyn_smem() {
    Runtime& r = rt();
    return (unsigned char*)(((uintptr_t)r.dyn_smem.data() + 63) & ~(uintptr_t)63);
}
inline int flat_tid() { return rt().current; }
inline int lane_id() { return rt().current & 31; }

// warp rendezvous: returns the deposited values of all lanes in `mask` (valid until this lane calls release())
inline Cell& collect(unsigned mask, unsigned long long v) {
    Runtime& r = rt();
    const int lane = r.current & 31;
    const unsigned bit = 1u << lane;
    if (!(mask & bit)) die("warp collective called by a lane that is not in its mask");
    Cell& c = r.warps[r.current >> 5].cells[mask];
    Fiber& f = r.fibers[r.current];
    while (c.drain) { f.wkind = 1; f.wcell = &c; yield(); }   // the previous collective of this mask is still being read
    c.vals[lane] = v;
    c.arrived |= bit;
    r.progress++;
    if (c.arrived == mask) { c.drain = true; c.toread = mask; }
    else while (!(c.drain && (c.toread & bit))) { f.wkind = 2; f.wcell = &c; f.wbit = bit; yield(); }
    return c;
}
inline void release(Cell& c) {
    const unsigned bit = 1u << lane_id();
    c.toread &= ~bit;
    rt().progress++;
    if (c.toread == 0) { c.arrived = 0; c.drain = false; }
}
// block until the phase bit (bit 32) of an emulated mbarrier word differs from `parity`
inline void wait_phase(const unsigned long long* word, unsigned parity) {
    Runtime& r = rt();
    Fiber& f = r.fibers[r.current];
    while ((((unsigned)(*(const volatile unsigned long long*)word >> 32)) & 1u) == parity) {
        f.wkind = 4; f.wword = word; f.wparity = parity;
        yield();
    }
}
inline void block_barrier(int id, int count) {
    Runtime& r = rt();
    BlockBarrier& b = r.barriers[id];
    Fiber& f = r.fibers[r.current];
    const int gen = b.generation;
    b.arrived++;
    r.progress++;
    for (;;) {
        if (b.generation != gen) return;
        // __syncthreads: threads that have exited count as arrived
        const int need = id < 0 ? r.alive : count;
        if (b.arrived >= need) { b.arrived = 0; b.generation++; r.progress++; return; }
        f.wkind = 3; f.wbar = &b; f.wgen = gen; f.wcount = count; f.wid = id;
        yield();
    }
}

}  // namespace emu

// ---------------------------------------------------------------- warp / block intrinsics
template <class T>
static inline unsigned long long emu_bits(T v) { unsigned long long u = 0; static_assert(sizeof(T) <= 8, "payload"); std::memcpy(&u, &v, sizeof(T)); return u; }
template <class T>
static inline T emu_unbits(unsigned long long u) { T v; std::memcpy(&v, &u, sizeof(T)); return v; }

template <class T>
static inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) {
    (void)width;
    emu::Cell& c = emu::collect(mask, emu_bits(v));
    const int s = src & 31;
    T out = (mask >> s) & 1u ? emu_unbits<T>(c.vals[s]) : v;
    emu::release(c);
    return out;
}
template <class T>
static inline T __shfl_up_sync(unsigned mask, T v, unsigned delta, int width = 32) {
    (void)width;
    emu::Cell& c = emu::collect(mask, emu_bits(v));
    const int s = emu::lane_id() - (int)delta;
    T out = (s >= 0 && ((mask >> s) & 1u)) ? emu_unbits<T>(c.vals[s]) : v;
    emu::release(c);
    return out;
}
template <class T>
static inline T __shfl_down_sync(unsigned mask, T v, unsigned delta, int width = 32) {
    (void)width;
    emu::Cell& c = emu::collect(mask, emu_bits(v));
    const int s = emu::lane_id() + (int)delta;
    T out = (s < 32 && ((mask >> s) & 1u)) ? emu_unbits<T>(c.vals[s]) : v;
    emu::release(c);
    return out;
}
template <class T>
static inline T __shfl_xor_sync(unsigned mask, T v, int lanemask, int width = 32) {
    (void)width;
    emu::Cell& c = emu::collect(mask, emu_bits(v));
    const int s = emu::lane_id() ^ lanemask;
    T out = (s < 32 && ((mask >> s) & 1u)) ? emu_unbits<T>(c.vals[s]) : v;
    emu::release(c);
    return out;
}
static inline unsigned __ballot_sync(unsigned mask, int pred) {
    emu::Cell& c = emu::collect(mask, pred ? 1ull : 0ull);
    unsigned out = 0;
    for (int l = 0; l < 32; ++l) if (((mask >> l) & 1u) && c.vals[l]) out |= 1u << l;
    emu::release(c);
    return out;
}
static inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0u; }
static inline int __all_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) == mask; }
static inline unsigned __reduce_or_sync(unsigned mask, unsigned v) {
    emu::Cell& c = emu::collect(mask, v);
    unsigned out = 0;
    for (int l = 0; l < 32; ++l) if ((mask >> l) & 1u) out |= (unsigned)c.vals[l];
    emu::release(c);
    return out;
}
static inline unsigned __reduce_min_sync(unsigned mask, unsigned v) {
    emu::Cell& c = emu::collect(mask, v);
    unsigned out = 0xffffffffu;
    for (int l = 0; l < 32; ++l) if ((mask >> l) & 1u) out = std::min(out, (unsigned)c.vals[l]);
    emu::release(c);
    return out;
}
static inline unsigned __reduce_max_sync(unsigned mask, unsigned v) {
    emu::Cell& c = emu::collect(mask, v);
    unsigned out = 0u;
    for (int l = 0; l < 32; ++l) if ((mask >> l) & 1u) out = std::max(out, (unsigned)c.vals[l]);
    emu::release(c);
    return out;
}
static inline unsigned __reduce_add_sync(unsigned mask, unsigned v) {
    emu::Cell& c = emu::collect(mask, v);
    unsigned out = 0u;
    for (int l = 0; l < 32; ++l) if ((mask >> l) & 1u) out += (unsigned)c.vals[l];
    emu::release(c);
    return out;
}
static inline void __syncwarp(unsigned mask = 0xffffffffu) {
    emu::Cell& c = emu::collect(mask, 0ull);
    emu::release(c);
}
static inline void __syncthreads() { emu::block_barrier(-1, 0); }
static inline void emu_named_barrier(int id, int count) { emu::block_barrier(id, count); }
static inline void __threadfence() {}
static inline void __threadfence_block() {}

// ---------------------------------------------------------------- atomics (one OS thread: plain read-modify-write)
template <class T> static inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
static inline unsigned atomicOr(unsigned* p, unsigned v) { unsigned o = *p; *p = o | v; return o; }
static inline unsigned atomicMin(unsigned* p, unsigned v) { unsigned o = *p; *p = o < v ? o : v; return o; }
static inline int atomicOr(int* p, int v) { int o = *p; *p = o | v; return o; }
static inline unsigned long long atomicOr(unsigned long long* p, unsigned long long v) { unsigned long long o = *p; *p = o | v; return o; }
static inline unsigned atomicMax(unsigned* p, unsigned v) { unsigned o = *p; *p = std::max(o, v); return o; }
static inline int atomicMax(int* p, int v) { int o = *p; *p = std::max(o, v); return o; }
static inline unsigned atomicExch(unsigned* p, unsigned v) { unsigned o = *p; *p = v; return o; }

// ---------------------------------------------------------------- scalar intrinsics (IEEE binary32, round to nearest even)
template <class T> static inline T __ldg(const T* p) { return *p; }
static inline float __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
static inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
static inline float __fdividef(float a, float b) { return a / b; }
static inline float __frcp_rn(float a) { return 1.0f / a; }
static inline int __float2int_rd(float f) { return (int)std::floor(f); }
static inline int __float2int_rn(float f) { return (int)std::nearbyint(f); }
static inline int __float2int_ru(float f) { return (int)std::ceil(f); }
static inline int __float2int_rz(float f) { return (int)f; }
static inline unsigned __float2uint_rz(float f) { return f <= 0.0f ? 0u : (f >= 4294967296.0f ? 0xffffffffu : (unsigned)f); }
static inline long long __double2ll_rd(double d) { return (long long)std::floor(d); }
static inline double __int2double_rn(int v) { return (double)v; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __clz(int v) { return v == 0 ? 32 : __builtin_clz((unsigned)v); }
static inline unsigned __brev(unsigned v) { unsigned r = 0; for (int i = 0; i < 32; ++i) r |= ((v >> i) & 1u) << (31 - i); return r; }
static inline unsigned __byte_perm(unsigned a, unsigned b, unsigned s) {
    const unsigned long long src = ((unsigned long long)b << 32) | a;
    unsigned out = 0;
    for (int k = 0; k < 4; ++k) {
        const unsigned sel = (s >> (4 * k)) & 0xfu;
        unsigned byte = (unsigned)(src >> (8 * (sel & 7u))) & 0xffu;
        if (sel & 8u) byte = (byte & 0x80u) ? 0xffu : 0x00u;   // sign-replication mode
        out |= byte << (8 * k);
    }
    return out;
}
static inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned s) { s &= 31u; return s ? (hi << s) | (lo >> (32 - s)) : hi; }
static inline unsigned __funnelshift_lc(unsigned lo, unsigned hi, unsigned s) { s = std::min(s, 32u); return s == 32 ? lo : (s ? (hi << s) | (lo >> (32 - s)) : hi); }
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned s) { s &= 31u; return s ? (lo >> s) | (hi << (32 - s)) : lo; }
static inline unsigned __funnelshift_rc(unsigned lo, unsigned hi, unsigned s) { s = std::min(s, 32u); return s == 32 ? hi : (s ? (lo >> s) | (hi << (32 - s)) : lo); }
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
static inline int __float_as_int(float f) { return emu_unbits<int>(emu_bits(f)); }
static inline unsigned __float_as_uint(float f) { return emu_unbits<unsigned>(emu_bits(f)); }
static inline float __int_as_float(int v) { return emu_unbits<float>(emu_bits(v)); }
static inline float __uint_as_float(unsigned v) { return emu_unbits<float>(emu_bits(v)); }
static inline long long __double_as_longlong(double d) { return emu_unbits<long long>(emu_bits(d)); }
static inline double __longlong_as_double(long long v) { return emu_unbits<double>(emu_bits(v)); }
using std::max;
using std::min;
static inline unsigned min(unsigned a, int b) { return std::min(a, (unsigned)b); }
static inline long long min(long long a, int b) { return std::min(a, (long long)b); }
static inline long long max(long long a, int b) { return std::max(a, (long long)b); }

// ---------------------------------------------------------------- CUDA runtime API (host memory, everything synchronous)
typedef int cudaError_t;
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1 };
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize, cudaFuncAttributePreferredSharedMemoryCarveout };
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount };
struct cudaDeviceProp { int major = 10, minor = 0, multiProcessorCount = 4; char name[64] = "cuda_emu (host)"; };
struct cudaFuncAttributes { size_t sharedSizeBytes = 0; int numRegs = 0; };
#ifndef TDE_EMU_SM_COUNT
#define TDE_EMU_SM_COUNT 4   // small grids: blocks run one after the other anyway
#endif
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) { *p = cudaDeviceProp(); p->multiProcessorCount = TDE_EMU_SM_COUNT; return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr, int) { *v = TDE_EMU_SM_COUNT; return cudaSuccess; }
static inline cudaError_t cudaMalloc(void** p, size_t n) { *p = std::aligned_alloc(256, (n + 255) & ~(size_t)255); return *p ? cudaSuccess : cudaErrorInvalidValue; }
static inline cudaError_t cudaFree(void* p) { std::free(p); return cudaSuccess; }
static inline cudaError_t cudaMallocHost(void** p, size_t n) { return cudaMalloc(p, n); }
static inline cudaError_t cudaFreeHost(void* p) { std::free(p); return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaMemset(void* p, int v, size_t n) { std::memset(p, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { std::memset(p, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { std::memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { std::memmove(d, s, n); return cudaSuccess; }
enum cudaMemoryType { cudaMemoryTypeUnregistered, cudaMemoryTypeHost, cudaMemoryTypeDevice, cudaMemoryTypeManaged };
struct cudaPointerAttributes { cudaMemoryType type = cudaMemoryTypeDevice; int device = 0; };
static inline cudaError_t cudaPointerGetAttributes(cudaPointerAttributes* a, const void*) { *a = cudaPointerAttributes(); return cudaSuccess; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline const char* cudaGetErrorString(cudaError_t) { return "cuda_emu"; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
enum cudaStreamCaptureStatus { cudaStreamCaptureStatusNone, cudaStreamCaptureStatusActive };
static inline cudaError_t cudaStreamIsCapturing(cudaStream_t, cudaStreamCaptureStatus* s) { *s = cudaStreamCaptureStatusNone; return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = nullptr; return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = (void*)1; return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
template <class K> static inline cudaError_t cudaFuncSetAttribute(K, cudaFuncAttribute, int) { return cudaSuccess; }
template <class K> static inline cudaError_t cudaFuncGetAttributes(cudaFuncAttributes* a, K) { *a = cudaFuncAttributes(); return cudaSuccess; }
template <class K> static inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int* n, K, int, size_t) { *n = 2; return cudaSuccess; }

#define TDE_LAUNCH(g, b, s, st, ...) emu::make_launcher(dim3(g), dim3(b), (size_t)(s), __VA_ARGS__)
