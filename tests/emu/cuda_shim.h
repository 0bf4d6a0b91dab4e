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
#include <vector>
#include <algorithm>

// ---- qualifiers ------------------------------------------------------------------------------
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))
#define __launch_bounds__(...)
#define __constant__ static

// ---- vector types ----------------------------------------------------------------------------
struct __attribute__((aligned(16))) uint4 { uint32_t x, y, z, w; };
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { uint4 v; v.x = x; v.y = y; v.z = z; v.w = w; return v; }
struct __attribute__((aligned(16))) double2 { double x, y; };
static inline double2 make_double2(double x, double y) { double2 v; v.x = x; v.y = y; return v; }
struct __attribute__((aligned(16))) ulonglong2 { unsigned long long x, y; };
static inline ulonglong2 make_ulonglong2(unsigned long long x, unsigned long long y) { ulonglong2 v; v.x = x; v.y = y; return v; }
struct EmuDim3 { unsigned x, y, z; };
extern EmuDim3 threadIdx, blockIdx, blockDim, gridDim;

// ---- the engine ------------------------------------------------------------------------------
namespace emu {
void launch(unsigned grid, unsigned block, const std::function<void()>& body);
void warp_barrier(unsigned mask);
void block_barrier();
unsigned lane_id();
unsigned alive_mask();                                        // live lanes of the calling warp
const uint64_t* allgather_begin(unsigned mask, uint64_t v);   // slot i = value of lane i (valid until allgather_end)
void allgather_end(unsigned mask);
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
static inline cudaError_t cudaMemset(void* d, int v, size_t n) { if (n) memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = nullptr) { if (n) memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = (void*)0x1; return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new EmuEvent(); return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = nullptr) { e->t = std::chrono::steady_clock::now(); return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) { *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count(); return cudaSuccess; }
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }
enum { cudaEventDisableTiming = 2, cudaIpcMemLazyEnablePeerAccess = 1, cudaMemPoolAttrReleaseThreshold = 4,
       cudaLimitPersistingL2CacheSize = 6, cudaStreamAttributeAccessPolicyWindow = 1 };
enum cudaMemoryType { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1, cudaMemoryTypeDevice = 2, cudaMemoryTypeManaged = 3 };
enum cudaAccessProperty { cudaAccessPropertyNormal = 0, cudaAccessPropertyStreaming = 1, cudaAccessPropertyPersisting = 2 };
struct cudaPointerAttributes { cudaMemoryType type; int device; void* devicePointer; void* hostPointer; };
static inline cudaError_t cudaPointerGetAttributes(cudaPointerAttributes* a, const void*) { a->type = cudaMemoryTypeDevice; a->device = 0; return cudaSuccess; }  // one address space: every pointer is "device" memory
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { return cudaEventCreate(e); }
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16 };
static inline cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr, int) { *v = 4; return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithPriority(cudaStream_t* s, unsigned, int) { *s = (void*)0x1; return cudaSuccess; }
static inline cudaError_t cudaDeviceGetStreamPriorityRange(int* lo, int* hi) { *lo = 0; *hi = -1; return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
static inline cudaError_t cudaMallocHost(void** p, size_t n) { *p = aligned_alloc(64, (n + 63) / 64 * 64); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
static inline cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
template <class T> static inline cudaError_t cudaMemcpyToSymbol(T& sym, const void* src, size_t n) { memcpy((void*)&sym, src, n); return cudaSuccess; }
struct cudaIpcMemHandle_t { char reserved[64]; };
static inline cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t*, void*) { return 1; }
static const cudaError_t cudaErrorPeerAccessAlreadyEnabled = 704;
static inline cudaError_t cudaDeviceCanAccessPeer(int* can, int, int) { *can = 1; return cudaSuccess; }
static inline cudaError_t cudaDeviceEnablePeerAccess(int, unsigned) { return cudaSuccess; }
static inline cudaError_t cudaIpcOpenMemHandle(void**, cudaIpcMemHandle_t, unsigned) { return 1; }
static inline cudaError_t cudaIpcCloseMemHandle(void*) { return cudaSuccess; }
typedef void* cudaMemPool_t;
static inline cudaError_t cudaDeviceGetDefaultMemPool(cudaMemPool_t* p, int) { *p = nullptr; return cudaSuccess; }
static inline cudaError_t cudaMemPoolSetAttribute(cudaMemPool_t, int, void*) { return cudaSuccess; }
static inline cudaError_t cudaMemPoolTrimTo(cudaMemPool_t, size_t) { return cudaSuccess; }
struct cudaDeviceProp { int persistingL2CacheMaxSize; int accessPolicyMaxWindowSize; };
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) { p->persistingL2CacheMaxSize = 0; p->accessPolicyMaxWindowSize = 0; return cudaSuccess; }
static inline cudaError_t cudaDeviceSetLimit(int, size_t) { return cudaSuccess; }
static inline cudaError_t cudaCtxResetPersistingL2Cache() { return cudaSuccess; }
struct cudaAccessPolicyWindow { void* base_ptr; size_t num_bytes; float hitRatio; cudaAccessProperty hitProp, missProp; };
union cudaStreamAttrValue { cudaAccessPolicyWindow accessPolicyWindow; int pad[16]; };
static inline cudaError_t cudaStreamSetAttribute(cudaStream_t, int, const cudaStreamAttrValue*) { return cudaSuccess; }

namespace cub {
template <class T> struct DoubleBuffer {
  T* d_buffers[2]; int selector;
  DoubleBuffer(T* a, T* b) { d_buffers[0] = a; d_buffers[1] = b; selector = 0; }
  T* Current() { return d_buffers[selector]; }
  T* Alternate() { return d_buffers[selector ^ 1]; }
};
struct DeviceRadixSort {
  // stable LSD radix sort semantics on key bits [b0, b1): result lands in the alternate buffers, selector flips
  template <class K, class V>
  static cudaError_t sort_impl(void* tmp, size_t& bytes, DoubleBuffer<K>& k, DoubleBuffer<V>* v, int64_t n, int b0, int b1, bool desc) {
    if (!tmp) { bytes = 16; return cudaSuccess; }
    std::vector<int64_t> idx((size_t)n);
    for (int64_t i = 0; i < n; i++) idx[(size_t)i] = i;
    const int w = b1 - b0;
    const uint64_t mask = w >= 64 ? ~0ull : ((1ull << w) - 1ull);
    K* src = k.Current();
    auto key = [&](int64_t i) { return ((uint64_t)src[i] >> b0) & mask; };
    if (desc) std::stable_sort(idx.begin(), idx.end(), [&](int64_t a, int64_t b) { return key(a) > key(b); });
    else std::stable_sort(idx.begin(), idx.end(), [&](int64_t a, int64_t b) { return key(a) < key(b); });
    K* dk = k.Alternate();
    for (int64_t i = 0; i < n; i++) dk[i] = src[idx[(size_t)i]];
    k.selector ^= 1;
    if (v) { V* sv = v->Current(); V* dv = v->Alternate(); for (int64_t i = 0; i < n; i++) dv[i] = sv[idx[(size_t)i]]; v->selector ^= 1; }
    return cudaSuccess;
  }
  template <class K, class V, class N> static cudaError_t SortPairs(void* t, size_t& b, DoubleBuffer<K>& k, DoubleBuffer<V>& v, N n, int b0, int b1, cudaStream_t = nullptr) { return sort_impl<K, V>(t, b, k, &v, (int64_t)n, b0, b1, false); }
  template <class K, class V, class N> static cudaError_t SortPairsDescending(void* t, size_t& b, DoubleBuffer<K>& k, DoubleBuffer<V>& v, N n, int b0, int b1, cudaStream_t = nullptr) { return sort_impl<K, V>(t, b, k, &v, (int64_t)n, b0, b1, true); }
  template <class K, class N> static cudaError_t SortKeys(void* t, size_t& b, DoubleBuffer<K>& k, N n, int b0, int b1, cudaStream_t = nullptr) { return sort_impl<K, char>(t, b, k, (DoubleBuffer<char>*)nullptr, (int64_t)n, b0, b1, false); }
};
struct DeviceSelect {
  template <class In, class Fl, class Out, class Cnt, class N>
  static cudaError_t Flagged(void* tmp, size_t& bytes, In in, Fl flags, Out out, Cnt count, N n, cudaStream_t = nullptr) {
    if (!tmp) { bytes = 16; return cudaSuccess; }
    int64_t c = 0;
    for (int64_t i = 0; i < (int64_t)n; i++) if (flags[i]) out[c++] = in[i];
    *count = c;
    return cudaSuccess;
  }
};
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
static inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
static inline int __ffsll(long long v) { return __builtin_ffsll(v); }
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
static inline double __dsub_rn(double a, double b) { volatile double r = a - b; return r; }
static inline double __ddiv_rn(double a, double b) { volatile double r = a / b; return r; }
static inline double __dsqrt_rn(double a) { volatile double r = __builtin_sqrt(a); return r; }
static inline double __ull2double_rn(unsigned long long v) { return (double)v; }
static inline unsigned long long __double2ull_rz(double d) { return d <= 0.0 ? 0ull : (d >= 18446744073709551615.0 ? ~0ull : (unsigned long long)d); }
static inline int __clzll(long long v) { return v == 0 ? 64 : __builtin_clzll((unsigned long long)v); }
static inline unsigned atomicCAS(unsigned* p, unsigned cmp, unsigned val) { const unsigned o = *p; if (o == cmp) *p = val; return o; }
static inline void __threadfence_block() {}
static inline void __threadfence() {}
static inline void __threadfence_system() {}
static inline __attribute__((always_inline)) void __syncthreads() { ::emu::block_barrier(); }
static inline __attribute__((always_inline)) void __syncwarp(unsigned mask = 0xffffffffu) { ::emu::warp_barrier(mask); }
template <class T> static inline T atomicAdd(T* p, T v) { const T o = *p; *p = o + v; return o; }
static inline unsigned atomicOr(unsigned* p, unsigned v) { const unsigned o = *p; *p = o | v; return o; }
static inline unsigned long long atomicOr(unsigned long long* p, unsigned long long v) { const unsigned long long o = *p; *p = o | v; return o; }

namespace emu {
template <class T> static inline uint64_t pack(T v) { static_assert(sizeof(T) <= 8, "payload"); uint64_t r = 0; memcpy(&r, &v, sizeof(T)); return r; }
template <class T> static inline T unpack(uint64_t r) { T v; memcpy(&v, &r, sizeof(T)); return v; }
}
#define EMU_INL static inline __attribute__((always_inline))
// source lane of a shuffle inside a segment of `width` lanes; out-of-segment sources return the caller's own value
template <class T> EMU_INL T emu_shfl(unsigned mask, T v, unsigned src_lane, bool valid) {
  const unsigned l = ::emu::lane_id();
  const uint64_t* s = ::emu::allgather_begin(mask, ::emu::pack(v));
  const T r = valid ? ::emu::unpack<T>(s[src_lane & 31u]) : v;
  ::emu::allgather_end(mask);
  (void)l;
  return r;
}
template <class T> EMU_INL T __shfl_sync(unsigned m, T v, int src, int width = 32) { const unsigned l = ::emu::lane_id(), w = (unsigned)width; return emu_shfl(m, v, (l & ~(w - 1u)) | ((unsigned)src & (w - 1u)), true); }
template <class T> EMU_INL T __shfl_up_sync(unsigned m, T v, unsigned d, int width = 32) { const unsigned l = ::emu::lane_id(), w = (unsigned)width; const bool ok = (l & (w - 1u)) >= d; return emu_shfl(m, v, l - d, ok); }
template <class T> EMU_INL T __shfl_down_sync(unsigned m, T v, unsigned d, int width = 32) { const unsigned l = ::emu::lane_id(), w = (unsigned)width; const bool ok = (l & (w - 1u)) + d < w; return emu_shfl(m, v, l + d, ok); }
template <class T> EMU_INL T __shfl_xor_sync(unsigned m, T v, int x, int width = 32) { const unsigned l = ::emu::lane_id(), w = (unsigned)width; const unsigned t = l ^ (unsigned)x; const bool ok = (t & ~(w - 1u)) == (l & ~(w - 1u)); return emu_shfl(m, v, t, ok); }
EMU_INL unsigned __ballot_sync(unsigned mask, bool p) {
  const unsigned members = mask & ::emu::alive_mask();
  const uint64_t* s = ::emu::allgather_begin(mask, p ? 1u : 0u);
  unsigned m = 0;
  for (int i = 0; i < 32; i++) if (((members >> i) & 1u) && s[i]) m |= 1u << i;
  ::emu::allgather_end(mask);
  return m;
}
EMU_INL int __any_sync(unsigned mask, bool p) { return __ballot_sync(mask, p) != 0; }
EMU_INL int __all_sync(unsigned mask, bool p) { return __ballot_sync(mask, !p) == 0; }
EMU_INL unsigned __reduce_min_sync(unsigned mask, unsigned v) {
  const unsigned members = mask & ::emu::alive_mask();
  const uint64_t* s = ::emu::allgather_begin(mask, v);
  unsigned m = 0xFFFFFFFFu;
  for (int i = 0; i < 32; i++) if ((members >> i) & 1u) m = (unsigned)s[i] < m ? (unsigned)s[i] : m;
  ::emu::allgather_end(mask);
  return m;
}
