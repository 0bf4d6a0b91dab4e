// cuda_shim.h -- lets g++ compile stract_b200/csrc/bm25.cu (+ its .cuh kernels) as plain C++ for the CPU SIMT
// emulator in emu_runtime.cpp.  TEST INFRASTRUCTURE ONLY: tests/test_bm25_emulated.py builds libsb200_emu.so from the
// unmodified kernel sources and runs the BM25 parity tests against it when no GPU is present, so warp-level logic
// (shuffles, ballots, barriers, shared-memory hazards that show up as wrong answers) is exercised before a GPU trip.
// It says nothing about performance and does not model memory consistency weaker than sequential.
//
// Model: every CUDA thread of a block is a ucontext coroutine; threads of a warp run until they all wait at a warp
// barrier (every *_sync intrinsic is one), warps run until they all wait at __syncthreads().  A warp whose live
// lanes wait at different kinds of barrier is reported as a divergence error.
#pragma once
#define SB200_EMU 1
#include <stdint.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <chrono>
#include <type_traits>
#include <functional>

// ---- qualifiers ------------------------------------------------------------------------------
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))
#define __launch_bounds__(...)

// ---- vector types ----------------------------------------------------------------------------
struct __attribute__((aligned(16))) uint4 { uint32_t x, y, z, w; };
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { uint4 v; v.x = x; v.y = y; v.z = z; v.w = w; return v; }
struct __attribute__((aligned(16))) double2 { double x, y; };
static inline double2 make_double2(double x, double y) { double2 v; v.x = x; v.y = y; return v; }
struct EmuDim3 { unsigned x, y, z; };
extern EmuDim3 threadIdx, blockIdx, blockDim, gridDim;

// ---- the engine ------------------------------------------------------------------------------
namespace emu {
void launch(unsigned grid, unsigned block, const std::function<void()>& body);
void warp_barrier();
void block_barrier();
unsigned lane_id();
unsigned alive_mask();                       // live lanes of the calling warp
const uint64_t* allgather(uint64_t v);       // one warp barrier; slot i = value of lane i
unsigned char* dyn_smem();
}
#define SB_DYN_SMEM(name) unsigned char* name = ::emu::dyn_smem()

// ---- runtime API stubs -----------------------------------------------------------------------
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2 };
typedef void* cudaStream_t;
struct EmuEvent { std::chrono::steady_clock::time_point t; };
typedef EmuEvent* cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };
enum { cudaStreamNonBlocking = 1, cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
static inline const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "emulated CUDA error"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaMalloc(void** p, size_t n) { *p = aligned_alloc(256, (n + 255) / 256 * 256 + 256); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
template <class T> static inline cudaError_t cudaMalloc(T** p, size_t n) { return cudaMalloc((void**)p, n); }
static inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMallocAsync(void** p, size_t n, cudaStream_t) { return cudaMalloc(p, n); }
static inline cudaError_t cudaFreeAsync(void* p, cudaStream_t) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { if (n) memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { if (n) memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = nullptr) { if (n) memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = (void*)0x1; return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new EmuEvent(); return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = nullptr) { e->t = std::chrono::steady_clock::now(); return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) { *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count(); return cudaSuccess; }
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }

namespace cub {
struct DeviceScan {
  template <class In, class Out>
  static cudaError_t ExclusiveSum(void* tmp, size_t& bytes, In in, Out out, int64_t n, cudaStream_t = nullptr) {
    if (!tmp) { bytes = 16; return cudaSuccess; }
    typename std::remove_reference<decltype(out[0])>::type acc = 0;
    for (int64_t i = 0; i < n; i++) { const auto v = in[i]; out[i] = acc; acc += v; }
    return cudaSuccess;
  }
};
}

// ---- launch ----------------------------------------------------------------------------------
#define SB_LAUNCH(kernel, grid, block, smem, stream, ...)                                          \
  do {                                                                                             \
    ::emu::launch((unsigned)(grid), (unsigned)(block), [&]() { kernel(__VA_ARGS__); });             \
    ::sb200::g_launches.fetch_add(1, std::memory_order_relaxed);                                   \
  } while (0)

// ---- device intrinsics -----------------------------------------------------------------------
template <class A, class B> static inline typename std::common_type<A, B>::type min(A a, B b) { return b < a ? b : a; }
template <class A, class B> static inline typename std::common_type<A, B>::type max(A a, B b) { return a < b ? b : a; }
template <class T> static inline T __ldg(const T* p) { return *p; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t sh) { return (uint32_t)(((((uint64_t)hi) << 32) | lo) >> (sh & 31u)); }
static inline uint32_t __float_as_uint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline long long __double_as_longlong(double d) { long long u; memcpy(&u, &d, 8); return u; }
static inline double __longlong_as_double(long long u) { double d; memcpy(&d, &u, 8); return d; }
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline float __fdiv_rn(float a, float b) { volatile float r = a / b; return r; }
static inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
static inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
static inline void __threadfence_block() {}
static inline __attribute__((always_inline)) void __syncthreads() { ::emu::block_barrier(); }
static inline __attribute__((always_inline)) void __syncwarp(unsigned = 0xffffffffu) { ::emu::warp_barrier(); }
template <class T> static inline T atomicAdd(T* p, T v) { const T o = *p; *p = o + v; return o; }
static inline unsigned atomicOr(unsigned* p, unsigned v) { const unsigned o = *p; *p = o | v; return o; }

namespace emu {
template <class T> static inline uint64_t pack(T v) { static_assert(sizeof(T) <= 8, "payload"); uint64_t r = 0; memcpy(&r, &v, sizeof(T)); return r; }
template <class T> static inline T unpack(uint64_t r) { T v; memcpy(&v, &r, sizeof(T)); return v; }
}
template <class T> static inline __attribute__((always_inline)) T __shfl_sync(unsigned, T v, int src) { const uint64_t* s = ::emu::allgather(::emu::pack(v)); return ::emu::unpack<T>(s[src & 31]); }
template <class T> static inline __attribute__((always_inline)) T __shfl_up_sync(unsigned, T v, unsigned d) { const unsigned l = ::emu::lane_id(); const uint64_t* s = ::emu::allgather(::emu::pack(v)); return l >= d ? ::emu::unpack<T>(s[l - d]) : v; }
template <class T> static inline __attribute__((always_inline)) T __shfl_down_sync(unsigned, T v, unsigned d) { const unsigned l = ::emu::lane_id(); const uint64_t* s = ::emu::allgather(::emu::pack(v)); return l + d < 32 ? ::emu::unpack<T>(s[l + d]) : v; }
template <class T> static inline __attribute__((always_inline)) T __shfl_xor_sync(unsigned, T v, int m) { const unsigned l = ::emu::lane_id(); const uint64_t* s = ::emu::allgather(::emu::pack(v)); return ::emu::unpack<T>(s[(l ^ (unsigned)m) & 31]); }
static inline __attribute__((always_inline)) unsigned __ballot_sync(unsigned, bool p) {
  const unsigned alive = ::emu::alive_mask();
  const uint64_t* s = ::emu::allgather(p ? 1u : 0u);
  unsigned m = 0;
  for (int i = 0; i < 32; i++) if (((alive >> i) & 1u) && s[i]) m |= 1u << i;
  return m;
}
static inline __attribute__((always_inline)) int __any_sync(unsigned mask, bool p) { return __ballot_sync(mask, p) != 0; }
static inline __attribute__((always_inline)) unsigned __reduce_min_sync(unsigned, unsigned v) {
  const unsigned alive = ::emu::alive_mask();
  const uint64_t* s = ::emu::allgather(v);
  unsigned m = 0xFFFFFFFFu;
  for (int i = 0; i < 32; i++) if ((alive >> i) & 1u) m = (unsigned)s[i] < m ? (unsigned)s[i] : m;
  return m;
}
