// cuda_emul.h — TEST INFRASTRUCTURE: lets g++ compile a .cu file of the product as plain C++ so that kernels whose threads do not
// cooperate (one thread, or one leading lane, per work item) can be EXECUTED on the CPU, single-source, against the oracle.
//
// Under g++ cuda_runtime.h already turns __device__/__global__/__host__/__shared__ into nothing and provides float4/int2/dim3 and the
// make_* helpers. What is added here: the built-in index variables, the IEEE-rounded arithmetic intrinsics (the host build uses
// -O2 -ffp-contract=off -fno-fast-math without -mfma, so a*b+c is never fused, as with nvcc --fmad=false), and compile-only stubs
// for the warp-cooperative intrinsics (kernels that use them are compiled but must not be called through this header).
#pragma once
#define MDG_HOST_EMULATION 1
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>

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
static inline void __syncwarp(unsigned = 0xffffffffu) {}

// compile-only stubs (abort when executed)
#define MDG_EMUL_STUB { abort(); }
static inline void __syncthreads() MDG_EMUL_STUB
template <typename T> static inline T __shfl_sync(unsigned, T, int, int = 32) MDG_EMUL_STUB
template <typename T> static inline T __shfl_up_sync(unsigned, T, unsigned, int = 32) MDG_EMUL_STUB
template <typename T> static inline T __shfl_down_sync(unsigned, T, unsigned, int = 32) MDG_EMUL_STUB
template <typename T> static inline T __shfl_xor_sync(unsigned, T, int, int = 32) MDG_EMUL_STUB
static inline unsigned __ballot_sync(unsigned, int) MDG_EMUL_STUB
static inline int __any_sync(unsigned, int) MDG_EMUL_STUB
static inline int __all_sync(unsigned, int) MDG_EMUL_STUB
static inline unsigned __reduce_or_sync(unsigned, unsigned) MDG_EMUL_STUB
static inline unsigned __reduce_add_sync(unsigned, unsigned) MDG_EMUL_STUB
static inline unsigned __reduce_max_sync(unsigned, unsigned) MDG_EMUL_STUB
static inline unsigned __reduce_min_sync(unsigned, unsigned) MDG_EMUL_STUB
static inline unsigned __activemask() MDG_EMUL_STUB
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
template <typename T, typename U> static inline T atomicAdd(T* p, U v) { T o = *p; *p = (T)(o + (T)v); return o; }
template <typename T> static inline T atomicMax(T* p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <typename T> static inline T atomicMin(T* p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <typename T> static inline T atomicExch(T* p, T v) { T o = *p; *p = v; return o; }
