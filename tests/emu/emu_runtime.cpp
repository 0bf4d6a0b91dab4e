// emu_runtime.cpp -- coroutine engine of the CPU SIMT emulator (see cuda_shim.h).  TEST INFRASTRUCTURE ONLY.
//
// Scheduling: threads of a block are ucontext coroutines.  A warp's lanes run one after the other until each waits
// at a barrier.  Warp-level barriers carry the member mask of the intrinsic (__syncwarp(mask), __shfl_sync(mask,..)):
// a group is released when every live lane named in the mask waits with the same mask, so sub-warp collectives
// (quads with their own trip counts) work like on the hardware.  When no lane of the block can run and no group is
// complete, the remaining lanes are at __syncthreads() -- or the kernel would hang on a GPU too, which is reported.
#include <ucontext.h>
#include <dlfcn.h>
#include <vector>

EmuDim3 threadIdx = {0, 0, 0}, blockIdx = {0, 0, 0}, blockDim = {1, 1, 1}, gridDim = {1, 1, 1};

namespace emu {

enum { RUN = 0, AT_WARP = 1, AT_BLOCK = 2 };
struct Lane {
  ucontext_t ctx; char* stack = nullptr; bool done = true; int wait = RUN; unsigned tid = 0;
  unsigned wmask = 0;   // member mask of the warp-level barrier this lane waits at
  int kind = 0;         // 0 = __syncwarp, 1 = first half of a collective (values written), 2 = second half (values read)
  void* site = nullptr;
};
static const size_t STACK_BYTES = 512 * 1024;
static std::vector<Lane> lanes;
static ucontext_t sched_ctx;
static Lane* cur = nullptr;
static const std::function<void()>* cur_body = nullptr;
static uint64_t slots[64][32];                          // [warp][lane]
static unsigned alive[64];
static unsigned char dyn[256 * 1024] __attribute__((aligned(128)));

unsigned char* dyn_smem() { return dyn; }
unsigned lane_id() { return cur->tid & 31u; }
unsigned alive_mask() { return alive[cur->tid >> 5]; }

static void wait_warp(unsigned mask, int kind, void* site) {
  cur->wait = AT_WARP; cur->wmask = mask; cur->kind = kind; cur->site = site;
  swapcontext(&cur->ctx, &sched_ctx);
}
void warp_barrier(unsigned mask) { wait_warp(mask, 0, __builtin_return_address(0)); }
void block_barrier() { cur->wait = AT_BLOCK; swapcontext(&cur->ctx, &sched_ctx); }
// write, meet, read, meet: the second meeting keeps a fast lane from overwriting a slot a slow one still reads
const uint64_t* allgather_begin(unsigned mask, uint64_t v) {
  Lane* me = cur;
  slots[me->tid >> 5][me->tid & 31u] = v;
  wait_warp(mask, 1, __builtin_return_address(0));
  return slots[me->tid >> 5];
}
void allgather_end(unsigned mask) { wait_warp(mask, 2, __builtin_return_address(0)); }

static void lane_main() {
  (*cur_body)();
  cur->done = true;
  alive[cur->tid >> 5] &= ~(1u << (cur->tid & 31u));
}

static void resume(Lane& L, unsigned block, unsigned nthreads, unsigned grid) {
  cur = &L;
  threadIdx.x = L.tid; threadIdx.y = threadIdx.z = 0;
  blockIdx.x = block; blockIdx.y = blockIdx.z = 0;
  blockDim.x = nthreads; blockDim.y = blockDim.z = 1;
  gridDim.x = grid; gridDim.y = gridDim.z = 1;
  swapcontext(&sched_ctx, &L.ctx);
}

static void die(const char* what, unsigned b, unsigned w, unsigned t0, unsigned t1) {
  Dl_info di; memset(&di, 0, sizeof(di));
  fprintf(stderr, "emu: block %u warp %u: %s\n", b, w, what);
  for (unsigned t = t0; t < t1; t++) if (!lanes[t].done) {
    dladdr(lanes[t].site, &di);
    fprintf(stderr, "  lane %2u: %s mask %08x kind %d @0x%lx\n", t & 31u, lanes[t].wait == AT_WARP ? "warp " : (lanes[t].wait == AT_BLOCK ? "block" : "run  "),
            lanes[t].wmask, lanes[t].kind, (unsigned long)((char*)lanes[t].site - (char*)di.dli_fbase));
  }
  abort();
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
      L.done = false; L.wait = RUN; L.tid = t; L.wmask = 0; L.kind = 0; L.site = nullptr;
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
          for (unsigned t = t0; t < t1; t++) if (!lanes[t].done && lanes[t].wait == RUN) resume(lanes[t], b, block, grid);
          // release every complete group of this warp
          bool released = false;
          unsigned handled = 0;
          for (unsigned t = t0; t < t1; t++) {
            if (lanes[t].done || lanes[t].wait != AT_WARP || ((handled >> (t & 31u)) & 1u)) continue;
            const unsigned M = lanes[t].wmask & alive[w];
            bool complete = true, kinds_ok = true;
            for (unsigned u = t0; u < t1; u++) if ((M >> (u & 31u)) & 1u) {
              if (lanes[u].wait != AT_WARP || (lanes[u].wmask & alive[w]) != M) { complete = false; break; }
              if (lanes[u].kind != lanes[t].kind) kinds_ok = false;
            }
            if (!complete) { handled |= 1u << (t & 31u); continue; }   // other groups of this warp may still be complete
            handled |= M;
            if (!kinds_ok) die("a __syncwarp() meets a shuffle/ballot (or two halves of different collectives meet)", b, w, t0, t1);
            for (unsigned u = t0; u < t1; u++) if ((M >> (u & 31u)) & 1u) lanes[u].wait = RUN;
            released = true;
          }
          if (released) continue;
          unsigned live = 0, at_warp = 0;
          for (unsigned t = t0; t < t1; t++) if (!lanes[t].done) { live++; at_warp += lanes[t].wait == AT_WARP; }
          if (at_warp) die("lanes wait at a warp-level barrier whose other members never arrive (divergent collective / barrier)", b, w, t0, t1);
          live_total += live;
          break;
        }
      }
      if (!live_total) break;
      for (unsigned t = 0; t < block; t++) if (!lanes[t].done) lanes[t].wait = RUN;   // everybody is at __syncthreads(): release
    }
  }
  cur = nullptr; cur_body = nullptr;
}

}  // namespace emu
