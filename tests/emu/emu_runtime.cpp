// emu_runtime.cpp -- coroutine engine of the CPU SIMT emulator (see cuda_shim.h) + the few process-wide symbols
// the BM25 translation units expect from the rest of the library.  TEST INFRASTRUCTURE ONLY.
#include <ucontext.h>
#include <dlfcn.h>
#include <stdarg.h>
#include <atomic>
#include <vector>

#include "../../stract_b200/csrc/common.cuh"

EmuDim3 threadIdx = {0, 0, 0}, blockIdx = {0, 0, 0}, blockDim = {1, 1, 1}, gridDim = {1, 1, 1};

namespace emu {

enum { RUN = 0, AT_WARP = 1, AT_BLOCK = 2 };
struct Lane { ucontext_t ctx; char* stack = nullptr; bool done = true; int wait = RUN; unsigned tid = 0; int parity = 0; void* site = nullptr; unsigned long nbar = 0; int kind = 0; };
static const size_t STACK_BYTES = 512 * 1024;
static std::vector<Lane> lanes;
static ucontext_t sched_ctx;
static Lane* cur = nullptr;
static const std::function<void()>* cur_body = nullptr;
static uint64_t slots[64][2][32];                       // [warp][parity][lane]
static unsigned alive[64];
static unsigned char dyn[256 * 1024] __attribute__((aligned(128)));

unsigned char* dyn_smem() { return dyn; }
unsigned lane_id() { return cur->tid & 31u; }
unsigned alive_mask() { return alive[cur->tid >> 5]; }

static void yield(int kind) { cur->wait = kind; cur->nbar++; swapcontext(&cur->ctx, &sched_ctx); }
void warp_barrier() { cur->site = __builtin_return_address(0); cur->kind = 0; yield(AT_WARP); }
void block_barrier() { yield(AT_BLOCK); }
const uint64_t* allgather(uint64_t v) {
  Lane* me = cur;
  const unsigned w = me->tid >> 5, l = me->tid & 31u;
  const int p = me->parity;
  slots[w][p][l] = v;
  me->site = __builtin_return_address(0); me->kind = 1;
  me->parity ^= 1;       // the next collective uses the other buffer: a fast lane cannot overwrite what a slow one still reads
  yield(AT_WARP);
  return slots[w][p];
}

static void lane_main() {
  (*cur_body)();
  cur->done = true;
  alive[cur->tid >> 5] &= ~(1u << (cur->tid & 31u));
  // returning switches to uc_link == sched_ctx
}

static void resume(Lane& L, unsigned block, unsigned nthreads, unsigned grid) {
  cur = &L;
  threadIdx.x = L.tid; threadIdx.y = threadIdx.z = 0;
  blockIdx.x = block; blockIdx.y = blockIdx.z = 0;
  blockDim.x = nthreads; blockDim.y = blockDim.z = 1;
  gridDim.x = grid; gridDim.y = gridDim.z = 1;
  swapcontext(&sched_ctx, &L.ctx);
}

void launch(unsigned grid, unsigned block, const std::function<void()>& body) {
  if (block == 0 || block > 2048) { fprintf(stderr, "emu: bad block size %u\n", block); abort(); }
  if (lanes.size() < block) lanes.resize(block);
  for (unsigned t = 0; t < block; t++) if (!lanes[t].stack) lanes[t].stack = (char*)aligned_alloc(64, STACK_BYTES);
  cur_body = &body;
  const unsigned nwarps = (block + 31) / 32;
  for (unsigned b = 0; b < grid; b++) {
    for (unsigned w = 0; w < nwarps; w++) alive[w] = 0;
    for (unsigned t = 0; t < block; t++) {
      Lane& L = lanes[t];
      L.done = false; L.wait = RUN; L.tid = t; L.parity = 0; L.nbar = 0;
      getcontext(&L.ctx);
      L.ctx.uc_stack.ss_sp = L.stack; L.ctx.uc_stack.ss_size = STACK_BYTES; L.ctx.uc_link = &sched_ctx;
      makecontext(&L.ctx, (void (*)())lane_main, 0);
      alive[t >> 5] |= 1u << (t & 31u);
    }
    for (;;) {
      unsigned live_total = 0;
      for (unsigned w = 0; w < nwarps; w++) {
        const unsigned t0 = w * 32, t1 = std::min(block, t0 + 32);
        for (;;) {
          bool ran = false;
          for (unsigned t = t0; t < t1; t++) if (!lanes[t].done && lanes[t].wait == RUN) { resume(lanes[t], b, block, grid); ran = true; }
          unsigned live = 0, at_warp = 0, at_block = 0;
          for (unsigned t = t0; t < t1; t++) if (!lanes[t].done) { live++; at_warp += lanes[t].wait == AT_WARP; at_block += lanes[t].wait == AT_BLOCK; }
          if (live && at_warp == live) {   // release the warp barrier
            {   // a __syncwarp() meeting a shuffle/ballot is never intended: the collective would exchange garbage
              int k0 = -1; bool mix = false;
              for (unsigned t = t0; t < t1; t++) if (!lanes[t].done) { if (k0 < 0) k0 = lanes[t].kind; else if (lanes[t].kind != k0) mix = true; }
              if (mix) {
                Dl_info di; memset(&di, 0, sizeof(di)); dladdr(lanes[t0].site, &di);
                fprintf(stderr, "emu: block %u warp %u: __syncwarp() meets a shuffle/ballot:", b, w);
                int prevk = -1;
                for (unsigned t = t0; t < t1; t++) if (!lanes[t].done && lanes[t].kind != prevk) { prevk = lanes[t].kind; fprintf(stderr, " lane %u.. %s n=%lu @0x%lx", t & 31u, prevk ? "collective" : "syncwarp", lanes[t].nbar, (unsigned long)((char*)lanes[t].site - (char*)di.dli_fbase)); }
                fprintf(stderr, "\n");
                abort();
              }
            }
            if (getenv("SB200_EMU_COUNTS")) {   // uniform control flow => every live lane has passed the same number of barriers
              unsigned long c0 = ~0ul; bool skew = false;
              for (unsigned t = t0; t < t1; t++) if (!lanes[t].done) { if (c0 == ~0ul) c0 = lanes[t].nbar; else if (lanes[t].nbar != c0) skew = true; }
              static int rep = 0;
              if (skew && rep++ < 4) {
                Dl_info di; memset(&di, 0, sizeof(di)); dladdr(lanes[t0].site, &di);
                fprintf(stderr, "emu: block %u warp %u: barrier COUNT skew:", b, w);
                unsigned long prevc = ~0ul;
                for (unsigned t = t0; t < t1; t++) if (!lanes[t].done && lanes[t].nbar != prevc) { prevc = lanes[t].nbar; fprintf(stderr, " lane %u.. n=%lu @0x%lx", t & 31u, prevc, (unsigned long)((char*)lanes[t].site - (char*)di.dli_fbase)); }
                fprintf(stderr, "\n");
              }
            }
            if (getenv("SB200_EMU_SITES")) {   // lanes meeting at DIFFERENT barrier call sites: legal for __syncwarp, fatal for collectives
              void* s0 = nullptr; bool mixed = false;
              for (unsigned t = t0; t < t1; t++) if (!lanes[t].done) { if (!s0) s0 = lanes[t].site; else if (lanes[t].site != s0) mixed = true; }
              static int reported = 0;
              if (mixed && reported++ < 8) {   // offsets are relative to the library: addr2line -e libsb200_emu.so <offset>
                Dl_info di; memset(&di, 0, sizeof(di));
                dladdr(s0, &di);
                fprintf(stderr, "emu: block %u warp %u: lanes meet at different barrier sites:", b, w);
                void* prev = nullptr;
                for (unsigned t = t0; t < t1; t++) if (!lanes[t].done && lanes[t].site != prev) {
                  prev = lanes[t].site;
                  fprintf(stderr, " lane %u.. @0x%lx", t & 31u, (unsigned long)((char*)prev - (char*)di.dli_fbase));
                }
                fprintf(stderr, "\n");
              }
            }
            for (unsigned t = t0; t < t1; t++) lanes[t].wait = RUN;
            continue;
          }
          if (at_warp && at_block) {
            fprintf(stderr, "emu: divergence error in block %u warp %u: %u lanes wait at a warp barrier, %u at __syncthreads()\n", b, w, at_warp, at_block);
            abort();
          }
          if (at_warp) {  // some lanes exited while the others wait for them at a warp collective: they would hang on a GPU too
            fprintf(stderr, "emu: block %u warp %u: %u of %u live lanes wait at a warp barrier that the rest never reaches\n", b, w, at_warp, live);
            abort();
          }
          (void)ran;
          break;
        }
        for (unsigned t = t0; t < t1; t++) live_total += !lanes[t].done;
      }
      if (!live_total) break;
      for (unsigned t = 0; t < block; t++) if (!lanes[t].done) lanes[t].wait = RUN;   // everybody is at __syncthreads(): release
    }
  }
  cur = nullptr; cur_body = nullptr;
}

}  // namespace emu

// ---- process-wide symbols normally provided by api_graph.cu / graph_stage.cu -------------------------------
namespace sb200 {
static thread_local char t_err[1024] = "";
std::atomic<uint64_t> g_launches{0};
thread_local cudaStream_t t_pool_stream = nullptr;
void set_error(const char* fmt, ...) { va_list ap; va_start(ap, fmt); vsnprintf(t_err, sizeof(t_err), fmt, ap); va_end(ap); }
bool arena_enabled() { return false; }
void* arena_alloc(size_t, cudaStream_t, int*) { return nullptr; }
void arena_free(void*, cudaStream_t, int) {}
void arena_retire_stream(int, cudaStream_t) {}
bool is_device_ptr(const void*) { return false; }
int copy_in(void* dst, const void* src, size_t bytes, cudaStream_t) { if (bytes) memmove(dst, src, bytes); return SB200_OK; }
}  // namespace sb200

extern "C" {
__attribute__((visibility("default"))) const char* sb200_last_error(void) { return sb200::t_err; }
__attribute__((visibility("default"))) const char* sb200_version(void) { return "stract_b200 CPU SIMT emulation (tests only)"; }
__attribute__((visibility("default"))) uint64_t sb200_kernel_launch_count(void) { return sb200::g_launches.load(); }
}
