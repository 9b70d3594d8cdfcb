// cuda_emul.h — TEST INFRASTRUCTURE: lets g++ compile a .cu file of the product as plain C++ so that kernels whose threads do not
// cooperate (one thread, or one leading lane, per work item) can be EXECUTED on the CPU, single-source, against the oracle.
//
// Under g++ cuda_runtime.h already turns __device__/__global__/__host__ into nothing and provides float4/int2/dim3 and the make_*
// helpers. What is added here: the built-in index variables, the IEEE-rounded arithmetic intrinsics (the host build uses
// -O2 -ffp-contract=off -fno-fast-math without -mfma, so a*b+c is never fused, as with nvcc --fmad=false), atomics, and two ways to run a kernel:
//   * thread by thread (the caller loops over blockIdx / threadIdx and calls the kernel function): enough for kernels whose threads do
//     not cooperate; the warp/block collectives abort if reached;
//   * emul_launch(grid, block, fn): every thread of a block is a fiber on the calling host thread, blocks run one after the other; __syncthreads / __syncwarp
//     are barriers, the warp collectives (__shfl_*_sync, __ballot_sync, __any/__all_sync, __reduce_*_sync) exchange through a per-warp
//     buffer, `__shared__` variables are function-local statics (one block at a time, so one copy is what a block sees). Threads that
//     return drop out of the barriers, as exited threads do on the device. Lanes named in a collective's mask must all reach it.
#pragma once
#define MDG_HOST_EMULATION 1
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <stdio.h>
#include <sys/mman.h>
#include <ucontext.h>
#include <functional>
#include <memory>
#include <mutex>
#include <vector>

#undef __shared__
#define __shared__ static

#undef __noinline__
#define __noinline__ __attribute__((noinline))
#undef __launch_bounds__
#define __launch_bounds__(...)

static thread_local uint3 threadIdx, blockIdx;
static thread_local dim3 blockDim, gridDim;

using std::max; using std::min;

static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fsqrt_rn(float a) { return sqrtf(a); }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline unsigned __float_as_uint(float a) { unsigned u; memcpy(&u, &a, 4); return u; }
static inline float __uint_as_float(unsigned u) { float a; memcpy(&a, &u, 4); return a; }
static inline int __float_as_int(float a) { int u; memcpy(&u, &a, 4); return u; }
static inline float __int_as_float(int u) { float a; memcpy(&a, &u, 4); return a; }
static inline unsigned long long __float2ull_rn(float a) { return (unsigned long long)rintf(a); }
static inline int __float2int_rn(float a) { return (int)rintf(a); }
static inline int __float2int_rz(float a) { return (int)a; }
static inline int __float2int_rd(float a) { return (int)floorf(a); }

// dynamic shared memory (extern __shared__) and 32-bit shared-window addresses: one arena, offsets into it are the "addresses"
alignas(16) static unsigned char emul_dyn_smem[228 * 1024];
template <typename T> static inline size_t __cvta_generic_to_shared(const T* p) { return (size_t)((const unsigned char*)p - emul_dyn_smem); }

// ---------------------------------------------------------------------------------------------- block / warp cooperation (emul_launch)
// Every thread of a block is a FIBER (ucontext) on the host thread that called emul_launch; blocks run one after the other. A barrier is
// "count, and while the generation has not advanced, switch to the next unfinished fiber of the block" — no OS threads, no futexes (the
// first version used one host thread per device thread and std::barrier; on the 8-core CI container the CPU suite spent 25 of its 32
// CPU-minutes in the kernel). Scheduling is round-robin in thread order and deterministic. Threads that return drop out of the barriers.
struct EmulBar { int expected = 0, count = 0; unsigned gen = 0; };
struct EmulWarp { EmulBar bar; unsigned long long slot[32]; };
struct EmulFiber { ucontext_t ctx; uint3 tid; int lane, warp; bool done; };
struct EmulBlock {
    EmulBar bar; std::vector<EmulWarp> warps; std::vector<EmulFiber> fibers; int cur = 0, live = 0; ucontext_t main_ctx;
    const std::function<void()>* fn = nullptr;
};
static thread_local EmulBlock* emul_block = nullptr;
static thread_local int emul_lane = 0, emul_warp = 0;

static inline void emul_enter(const EmulFiber& f) { threadIdx = f.tid; emul_lane = f.lane; emul_warp = f.warp; }   // this translation unit's view of "which thread am I"
static inline void emul_yield() {
    EmulBlock* b = emul_block; const int n = (int)b->fibers.size(), me = b->cur; int nx = me;
    do { nx = nx + 1 == n ? 0 : nx + 1; } while (b->fibers[nx].done && nx != me);
    if (nx == me) return;
    b->cur = nx; swapcontext(&b->fibers[me].ctx, &b->fibers[nx].ctx);
    emul_enter(b->fibers[me]);
}
static inline void emul_bar_wait(EmulBar& x) {
    const unsigned g = x.gen;
    if (++x.count >= x.expected) { x.count = 0; ++x.gen; return; }
    unsigned long long spins = 0;
    while (x.gen == g) { emul_yield(); if (++spins > (1ull << 26)) { fprintf(stderr, "cuda_emul: barrier never completed (divergent collective?)\n"); abort(); } }
}
static inline void emul_bar_drop(EmulBar& x) { if (--x.expected > 0 && x.count >= x.expected) { x.count = 0; ++x.gen; } }

static inline EmulWarp& emul_w() { if (!emul_block) abort(); /* collective reached in thread-by-thread mode */ return emul_block->warps[emul_warp]; }
static inline void __syncthreads() { if (!emul_block) abort(); emul_bar_wait(emul_block->bar); }
static inline void __syncwarp(unsigned = 0xffffffffu) { if (emul_block) emul_bar_wait(emul_w().bar); }   // thread-by-thread mode: lanes run one after the other

// every lane publishes a value, then reads the slot of `src` (two warp barriers: publish | read)
template <typename T> static inline T emul_exchange(T v, int src) {
    static_assert(sizeof(T) <= 8, "shuffle of more than 8 bytes");
    EmulWarp& w = emul_w(); unsigned long long bits = 0; memcpy(&bits, &v, sizeof(T)); w.slot[emul_lane] = bits;
    emul_bar_wait(w.bar);
    T r; const unsigned long long got = w.slot[src & 31]; memcpy(&r, &got, sizeof(T));
    emul_bar_wait(w.bar);
    return r;
}
template <typename T> static inline T __shfl_sync(unsigned, T v, int src, int width = 32) { return emul_exchange(v, (emul_lane & ~(width - 1)) | (src & (width - 1))); }
template <typename T> static inline T __shfl_up_sync(unsigned, T v, unsigned d, int width = 32) { const int s = emul_lane - (int)d; return emul_exchange(v, s >= (emul_lane & ~(width - 1)) ? s : emul_lane); }
template <typename T> static inline T __shfl_down_sync(unsigned, T v, unsigned d, int width = 32) { const int s = emul_lane + (int)d; return emul_exchange(v, s <= (emul_lane | (width - 1)) ? s : emul_lane); }
template <typename T> static inline T __shfl_xor_sync(unsigned, T v, int m, int width = 32) { const int s = emul_lane ^ m; return emul_exchange(v, s <= (emul_lane | (width - 1)) ? s : emul_lane); }
// gather of one value per lane over the lanes of `mask`, folded by f
template <typename F> static inline unsigned emul_fold(unsigned mask, unsigned v, unsigned init, F f) {
    EmulWarp& w = emul_w(); w.slot[emul_lane] = v;
    emul_bar_wait(w.bar);
    unsigned r = init; for (int l = 0; l < 32; ++l) if (mask >> l & 1u) r = f(r, (unsigned)w.slot[l], l);
    emul_bar_wait(w.bar);
    return r;
}
static inline unsigned __ballot_sync(unsigned mask, int pred) { return emul_fold(mask, pred ? 1u : 0u, 0u, [](unsigned r, unsigned v, int l) { return r | (v << l); }); }
static inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
static inline int __all_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) == mask; }
static inline unsigned __reduce_or_sync(unsigned mask, unsigned v) { return emul_fold(mask, v, 0u, [](unsigned r, unsigned x, int) { return r | x; }); }
static inline unsigned __reduce_add_sync(unsigned mask, unsigned v) { return emul_fold(mask, v, 0u, [](unsigned r, unsigned x, int) { return r + x; }); }
static inline unsigned __reduce_max_sync(unsigned mask, unsigned v) { return emul_fold(mask, v, 0u, [](unsigned r, unsigned x, int) { return r > x ? r : x; }); }
static inline unsigned __reduce_min_sync(unsigned mask, unsigned v) { return emul_fold(mask, v, 0xffffffffu, [](unsigned r, unsigned x, int) { return r < x ? r : x; }); }

// fiber stacks of the calling host thread (kept for its lifetime; pages are touched only as deep as a kernel's frames go)
static constexpr size_t EMUL_STACK_BYTES = 512 * 1024;
struct EmulStacks { std::vector<void*> s; ~EmulStacks() { for (void* p : s) munmap(p, EMUL_STACK_BYTES); } };
static inline void* emul_stack(int t) {
    static thread_local EmulStacks st;
    while ((int)st.s.size() <= t) { void* p = mmap(nullptr, EMUL_STACK_BYTES, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0); if (p == MAP_FAILED) abort(); st.s.push_back(p); }
    return st.s[t];
}
static void emul_fiber_main() {
    EmulBlock* b = emul_block; EmulFiber& f = b->fibers[b->cur];
    emul_enter(f);
    (*b->fn)();
    b = emul_block; EmulFiber& me = b->fibers[b->cur];                 // (the fiber may have been resumed any number of times since)
    b->warps[me.warp].slot[me.lane] = 0;                               // an exited lane contributes 0 to later ballots
    me.done = true; --b->live; emul_bar_drop(b->warps[me.warp].bar); emul_bar_drop(b->bar);
    if (b->live == 0) { setcontext(&b->main_ctx); abort(); }
    const int n = (int)b->fibers.size(); int nx = b->cur; do { nx = nx + 1 == n ? 0 : nx + 1; } while (b->fibers[nx].done);
    b->cur = nx; setcontext(&b->fibers[nx].ctx); abort();
}

// run fn() as every thread of every block of the grid; blocks one after the other.
// One launch at a time, process-wide: `__shared__` variables are single static copies, so kernels enqueued by different host threads
// (concurrent callers of one plan, the per-device threads of a multi-device plan) must not overlap here as they may on a device.
inline std::mutex& emul_launch_mutex() { static std::mutex m; return m; }
static inline void emul_launch(dim3 grid, dim3 block, const std::function<void()>& fn) {
    std::lock_guard<std::mutex> emul_guard(emul_launch_mutex());
    const int nthreads = (int)(block.x * block.y * block.z), nwarps = (nthreads + 31) / 32;
    if (!nthreads) return;
    EmulBlock blk; blk.warps.resize(nwarps); blk.fibers.resize(nthreads); blk.fn = &fn;
    gridDim = grid; blockDim = block;
    const uint3 saved_tid = threadIdx, saved_bid = blockIdx;
    for (unsigned bz = 0; bz < grid.z; ++bz) for (unsigned by = 0; by < grid.y; ++by) for (unsigned bx = 0; bx < grid.x; ++bx) {
        blockIdx.x = bx; blockIdx.y = by; blockIdx.z = bz;
        blk.bar = EmulBar{nthreads, 0, 0};
        for (int w = 0; w < nwarps; ++w) { blk.warps[w].bar = EmulBar{std::min(32, nthreads - 32 * w), 0, 0}; memset(blk.warps[w].slot, 0, sizeof(blk.warps[w].slot)); }
        for (int t = 0; t < nthreads; ++t) {
            EmulFiber& f = blk.fibers[t];
            f.tid.x = t % block.x; f.tid.y = (t / block.x) % block.y; f.tid.z = t / (block.x * block.y); f.lane = t & 31; f.warp = t >> 5; f.done = false;
            getcontext(&f.ctx); f.ctx.uc_stack.ss_sp = emul_stack(t); f.ctx.uc_stack.ss_size = EMUL_STACK_BYTES; f.ctx.uc_link = nullptr;
            makecontext(&f.ctx, emul_fiber_main, 0);
        }
        blk.cur = 0; blk.live = nthreads; emul_block = &blk;
        swapcontext(&blk.main_ctx, &blk.fibers[0].ctx);      // returns when the last fiber of the block has finished
        emul_block = nullptr;
    }
    threadIdx = saved_tid; blockIdx = saved_bid; emul_lane = 0; emul_warp = 0;
}

static inline unsigned __byte_perm(unsigned x, unsigned y, unsigned sel) {   // prmt.b32, default mode
    const unsigned long long v = ((unsigned long long)y << 32) | x; unsigned r = 0;
    for (int i = 0; i < 4; ++i) {
        const unsigned n = (sel >> (4 * i)) & 0xfu; unsigned b = (unsigned)(v >> (8 * (n & 7u))) & 0xffu;
        if (n & 8u) b = (b & 0x80u) ? 0xffu : 0x00u;   // msb replication
        r |= b << (8 * i);
    }
    return r;
}
static inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned sh) { sh &= 31u; return sh ? (hi << sh) | (lo >> (32 - sh)) : hi; }
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned sh) { sh &= 31u; return sh ? (lo >> sh) | (hi << (32 - sh)) : lo; }
static inline unsigned __funnelshift_lc(unsigned lo, unsigned hi, unsigned sh) { return sh >= 32u ? lo : __funnelshift_l(lo, hi, sh); }
static inline unsigned __funnelshift_rc(unsigned lo, unsigned hi, unsigned sh) { return sh >= 32u ? hi : __funnelshift_r(lo, hi, sh); }
static inline unsigned __brev(unsigned v) { unsigned r = 0; for (int i = 0; i < 32; ++i) r |= ((v >> i) & 1u) << (31 - i); return r; }
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
static inline unsigned long long __umul64hi(unsigned long long a, unsigned long long b) { return (unsigned long long)(((unsigned __int128)a * b) >> 64); }
static inline int __clzll(long long v) { return v ? __builtin_clzll((unsigned long long)v) : 64; }
static inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
static inline int __ffsll(long long v) { return __builtin_ffsll(v); }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
template <typename T, typename U> static inline T atomicAdd(T* p, U v) { return __atomic_fetch_add(p, (T)v, __ATOMIC_RELAXED); }
static inline float atomicAdd(float* p, float v) { float o = *p, n; do { n = o + v; } while (!__atomic_compare_exchange(p, &o, &n, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)); return o; }
template <typename T> static inline T atomicMax(T* p, T v) { T o = *p; while (o < v && !__atomic_compare_exchange_n(p, &o, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {} return o; }
template <typename T> static inline T atomicMin(T* p, T v) { T o = *p; while (o > v && !__atomic_compare_exchange_n(p, &o, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {} return o; }
template <typename T, typename U> static inline T atomicExch(T* p, U v) { return __atomic_exchange_n(p, (T)v, __ATOMIC_RELAXED); }
template <typename T> static inline T atomicOr(T* p, T v) { return __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }
template <typename T> static inline T atomicCAS(T* p, T cmp, T v) { __atomic_compare_exchange_n(p, &cmp, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED); return cmp; }
