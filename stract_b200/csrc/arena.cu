// arena.cu -- CUDA backend of the opt-in slab arena (arena.h) + its diagnostics in the C ABI
#include "common.cuh"
#include "arena.h"

#include <stdlib.h>
#include <iterator>
#include <random>

namespace sb200 {

static void* cuda_slab_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaMalloc(&p, bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return p;
}
static void cuda_slab_free(void* p) { cudaFree(p); }
static void cuda_stream_sync(void* s) { if (cudaStreamSynchronize((cudaStream_t)s) != cudaSuccess) cudaGetLastError(); }

static Arena* g_arena[64] = {nullptr};
static std::mutex g_arena_mu;

bool arena_enabled() {
  static const bool on = env_flag("SB200_ARENA", true);  // default since round 2 (SB200_ARENA=0: the driver's stream-ordered pool)
  return on;
}
static Arena* arena_of(int dev, bool create) {
  if (dev < 0 || dev >= 64) return nullptr;
  std::lock_guard<std::mutex> lk(g_arena_mu);
  if (!g_arena[dev] && create) {
    size_t min_slab = (size_t)2 << 30;
    if (const char* e = getenv("SB200_ARENA_SLAB_MB")) { const long mb = atol(e); if (mb > 0) min_slab = (size_t)mb << 20; }
    g_arena[dev] = new (std::nothrow) Arena(ArenaBackend{cuda_slab_alloc, cuda_slab_free, cuda_stream_sync}, min_slab);
  }
  return g_arena[dev];
}

void* arena_alloc(size_t bytes, cudaStream_t stream, int* dev_out) {
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
  Arena* a = arena_of(dev, true);
  if (!a) return nullptr;
  void* p = a->alloc(bytes, (void*)stream);
  if (p && dev_out) *dev_out = dev;
  return p;
}
void arena_free(void* p, cudaStream_t stream, int dev) {
  Arena* a = arena_of(dev, false);
  if (a) a->free(p, (void*)stream);
}
void arena_retire_stream(int dev, cudaStream_t stream) {
  Arena* a = arena_of(dev, false);
  if (a) a->retire_stream((void*)stream);
}

}  // namespace sb200

using namespace sb200;

extern "C" {

int sb200_arena_stats(int device, uint64_t* reserved, uint64_t* in_use, uint64_t* peak, uint64_t* slabs) {
  Arena* a = arena_of(device, false);
  if (reserved) *reserved = a ? a->reserved() : 0;
  if (in_use) *in_use = a ? a->in_use() : 0;
  if (peak) *peak = a ? a->peak() : 0;
  if (slabs) *slabs = a ? a->n_slabs() : 0;
  return SB200_OK;
}

int sb200_arena_trim(int device) {
  Arena* a = arena_of(device, false);
  if (!a) return SB200_OK;
  int prev = -1;
  SB_CUDA(cudaGetDevice(&prev));
  SB_CUDA(cudaSetDevice(device));
  a->trim();
  SB_CUDA(cudaSetDevice(prev));
  return SB200_OK;
}

// Give cached device memory back to the driver: the stream-ordered pool the staging pipeline keeps warm (release
// threshold = max) and the slab arena's empty slabs.  For processes that are done with a large graph.
int sb200_release_cached_memory(int device) {
  int prev = -1;
  SB_CUDA(cudaGetDevice(&prev));
  SB_CUDA(cudaSetDevice(device));
  SB_CUDA(cudaDeviceSynchronize());
  cudaMemPool_t pool;
  SB_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
  SB_CUDA(cudaMemPoolTrimTo(pool, 0));
  Arena* a = arena_of(device, false);
  if (a) a->trim();
  SB_CUDA(cudaSetDevice(prev));
  return SB200_OK;
}

// Randomised self-test of the allocator logic over host memory (no GPU): returns 0 when every invariant held.
static int g_synced = 0;
int sb200_arena_selftest(uint64_t seed, uint32_t ops) {
  g_synced = 0;
  ArenaBackend be{[](size_t n) -> void* { return malloc(n); }, [](void* p) { free(p); }, [](void*) { g_synced++; }};
  Arena a(be, (size_t)1 << 20);
  std::mt19937_64 rng(seed);
  struct Rec { char* p; size_t n; unsigned char fill; };
  std::vector<Rec> live;
  void* streams[3] = {(void*)0x10, (void*)0x20, (void*)0x30};
  for (uint32_t i = 0; i < ops; i++) {
    const bool do_alloc = live.empty() || (rng() % 100) < 55;
    void* st = streams[rng() % 3];
    if (do_alloc) {
      size_t n = (rng() % 8 == 0) ? (size_t)(rng() % (3u << 20)) : (size_t)(rng() % 40000);
      char* p = (char*)a.alloc(n, st);
      if (!p) return 1;
      if ((uintptr_t)p % Arena::ALIGN) return 2;
      const unsigned char fill = (unsigned char)(rng() & 0xff);
      memset(p, fill, n);
      live.push_back(Rec{p, n, fill});
    } else {
      const size_t k = rng() % live.size();
      Rec r = live[k]; live[k] = live.back(); live.pop_back();
      for (size_t j = 0; j < r.n; j += 97) if ((unsigned char)r.p[j] != r.fill) return 3;  // somebody else wrote into it
      if (!a.free(r.p, st)) return 4;
    }
    if (i % 64 == 0 && !a.check()) return 5;
    if (i % 1000 == 999) a.retire_stream(streams[rng() % 3]);
  }
  for (auto& r : live) {
    for (size_t j = 0; j < r.n; j += 97) if ((unsigned char)r.p[j] != r.fill) return 6;
    if (!a.free(r.p, streams[0])) return 7;
  }
  if (a.free((void*)0x1234, nullptr)) return 8;  // foreign pointer must be rejected
  if (a.in_use() != 0 || a.n_live() != 0 || !a.check()) return 9;
  for (void* s : streams) a.retire_stream(s);
  if (!a.check() || a.n_free_blocks() != a.n_slabs()) return 10;  // everything merged back: one block per slab
  const size_t slabs_before = a.n_slabs(), reserved_before = a.reserved();
  // steady state: replaying an identical allocation sequence must not grow the arena
  for (int rep = 0; rep < 3; rep++) {
    std::mt19937_64 r2(seed ^ 0x9e3779b97f4a7c15ull);
    std::vector<void*> ps;
    for (int i = 0; i < 200; i++) { void* p = a.alloc((size_t)(r2() % 300000), streams[0]); if (!p) return 11; ps.push_back(p); }
    for (size_t i = 0; i < ps.size(); i += 2) a.free(ps[i], streams[0]);
    for (size_t i = 1; i < ps.size(); i += 2) a.free(ps[i], streams[0]);
    if (rep == 0) continue;  // the first replay may add slabs; later ones must not
    static size_t grown_slabs, grown_reserved;
    if (rep == 1) { grown_slabs = a.n_slabs(); grown_reserved = a.reserved(); }
    if (rep == 2 && (a.n_slabs() != grown_slabs || a.reserved() != grown_reserved)) return 12;
  }
  (void)slabs_before; (void)reserved_before;
  a.trim();
  if (a.n_slabs() != 0 || a.reserved() != 0) return 13;
  return 0;
}

}  // extern "C"
