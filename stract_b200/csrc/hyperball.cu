// hyperball.cu -- the HyperBall iteration on the device (hot path 1).
//
// Reference semantics (crates/core/src/webgraph/centrality/harmonic.rs):
//   update_all_counters  :116-157   for every kept edge u->v with u in the changed set:
//                                   if any(old[u][i] > new[v][i]) { new[v].merge(old[u]); mark v }
//   update_changed_counters :75-114 same, driven from the exact changed set through forward links
//   update_centralities  :159-176   score[v] += (size(new[v]) -sat size(old[v])) as f64 / (t+1)  (Kahan)
//   Counters::step       :210-212   old = new.clone()
// All branches compute the synchronous update new[v] = max(old[v], max_{u->v, u changed} old[u]);
// merging an unchanged u is a no-op (it was merged the iteration after it last changed), which is
// why the reference may use a Bloom filter for the changed set and why the dense kernel below may
// skip the frontier test altogether.
//
// Device formulation
//   * registers live in two N x 64 B arrays (ping-pong).  Row v of the "new" array is rewritten only
//     if v changed now or changed in the previous iteration (it is then two iterations stale);
//     every other row is already equal in both arrays, so `old = new.clone()` costs nothing.
//   * pull kernels (destination-major CSR, one writer per row, no atomics on registers):
//       k_pull_quad : rows with in-degree <= 32, 4 lanes per row (16 B of the 64-B counter each)
//       k_pull_warp : longer rows cut into 1024-edge work items, one warp per item: the 32 source
//                     indices of a batch are loaded coalesced, each quad of lanes gathers one 64-B
//                     counter (4 x LDG.128), byte-wise max with __vmaxu4, xor-shuffle reduce over
//                     the 8 quads; rows spanning several items go through a partial buffer + k_pull_merge
//     DENSE variant gathers every in-neighbour; FRONTIER variant tests the changed bitmap first.
//   * push kernel (source-major CSR) for small frontiers: half-warp per out-edge, 32-bit CAS max.
//   * k_finalize: per changed node, HyperLogLog<64>::size() in the reference's exact f64 operation
//     order (sequential sum over registers 0..63, no FMA), saturating difference against the cached
//     size(old), KahanSum update; nodes that changed in the previous iteration but not now get the
//     reference's "+= 0.0" (it is idempotent after one application, so skipping the other N-1 zero
//     adds is bit-exact).
// Roofline: HBM.  Algorithmic bytes per iteration: 68 B x E_active + 132 B x N_written + 40 B x N_changed.
#include <chrono>
#include "graph.cuh"
#include "../../include/sb200_hll_tables.h"

#ifndef SB200_EMU
#include <cub/cub.cuh>
#endif
#include <algorithm>
#include <cmath>
#include <cstdlib>

namespace sb200 {

struct HllTables {
  double raw[SB200_HLL_P5_LEN];
  double bias[SB200_HLL_P5_LEN];
  double lc[65];  // lc[v] = 64*ln(64/v), host libm (matches the reference's f64::ln), lc[0] unused
};
__constant__ HllTables c_tab;
static bool g_tab_loaded[64] = {false};

int load_tables(int device) {
  if (device >= 0 && device < 64 && g_tab_loaded[device]) return SB200_OK;
  static HllTables h;
  for (int i = 0; i < SB200_HLL_P5_LEN; i++) { h.raw[i] = SB200_HLL_RAW_P5[i]; h.bias[i] = SB200_HLL_BIAS_P5[i]; }
  h.lc[0] = 0.0;
  for (int v = 1; v <= 64; v++) h.lc[v] = 64.0 * std::log(64.0 / (double)v);  // hyperloglog.rs:4472-4476
  SB_CUDA(cudaMemcpyToSymbol(c_tab, &h, sizeof(h)));
  if (device >= 0 && device < 64) g_tab_loaded[device] = true;
  return SB200_OK;
}

// ---- HyperLogLog<64>::size(), crates/core/src/hyperloglog.rs:4484-4516, bit-exact -----------------
// `tab` points to a shared-memory copy of c_tab (divergent indexing into __constant__ serialises).
__device__ __forceinline__ double pow2_neg(uint32_t k) {  // ONE_OVER_POWER_OF_TWO[k] == 2^-k
  return __longlong_as_double((long long)(1023 - (int)k) << 52);
}
__device__ uint64_t hll64_size(const uint32_t w[16], const HllTables* tab) {
  double sum = 0.0;
  int zeros = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) {
#pragma unroll
    for (int b = 0; b < 4; b++) {
      uint32_t r = (w[i] >> (8 * b)) & 0xFFu;
      sum = __dadd_rn(sum, pow2_neg(r));
      zeros += (r == 0);
    }
  }
  const double z = __ddiv_rn(1.0, sum);
  const double e = __dmul_rn(0.709 * 4096.0, z);  // am() * m.powi(2) * z, left-assoc; 0.709*2^12 is exact
  double e_star = e;
  if (e <= 320.0) {
    // estimate_bias(e, 6): tables [b-1-4] = "precision 5"; Rust >= 1.82 binary_search_by
    const int LEN = SB200_HLL_P5_LEN;
    int size = LEN, base = 0;
    while (size > 1) {
      int half = size >> 1, mid = base + half;
      base = (tab->raw[mid] > e) ? base : mid;
      size -= half;
    }
    int r = base;
    if (!(tab->raw[base] == e)) r = base + (tab->raw[base] < e ? 1 : 0);
    if (r == LEN) r = LEN - 1;
    int il = r, ir = (r < LEN - 1) ? r + 1 : -1;
    double bsum = 0.0;
#pragma unroll 1
    for (int k = 0; k < 6; k++) {
      bool right;
      int idx;
      if (il >= 0 && ir >= 0) {
        double dl = fabs(__dsub_rn(tab->raw[il], e)), dr = fabs(__dsub_rn(tab->raw[ir], e));
        right = dr < dl;
        idx = right ? ir : il;
      } else if (il >= 0) { right = false; idx = il; }
      else { right = true; idx = ir; }
      bsum = __dadd_rn(bsum, tab->bias[idx]);
      if (right) ir = (idx < LEN - 1) ? idx + 1 : -1;
      else il = (idx > 0) ? idx - 1 : -1;
    }
    e_star = __dsub_rn(e, __ddiv_rn(bsum, 6.0));
  }
  const double h = (zeros != 0) ? tab->lc[zeros] : e_star;
  const double pick = (h <= 40.0) ? h : e_star;  // threshold(b=6) = 40
  // Rust `as usize`: saturating, NaN -> 0
  if (!(pick == pick) || pick <= 0.0) return 0ull;
  if (pick >= 18446744073709551616.0) return 0xFFFFFFFFFFFFFFFFull;
  return (uint64_t)__double2ull_rz(pick);
}

__device__ __forceinline__ void kahan_add(double& sum, double& err, double rhs) {  // kahan_sum.rs:46-53
  const double y = __dsub_rn(rhs, err);
  const double t = __dadd_rn(sum, y);
  err = __dsub_rn(__dsub_rn(t, sum), y);
  sum = t;
}

// ---- init: counters seeded with the node's own id (low 64 bits), harmonic.rs:53-73 ------------------
__global__ void k_hb_init(const uint64_t* __restrict__ id_lo, const uint32_t* __restrict__ perm, uint64_t N,
                          uint4* r0, uint4* r1, uint64_t* size_cache, double* ksum, double* kerr) {
  __shared__ HllTables tab;
  for (int i = threadIdx.x; i < (int)(sizeof(HllTables) / 8); i += blockDim.x) ((double*)&tab)[i] = ((const double*)&c_tab)[i];
  __syncthreads();
  uint64_t v = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (v >= N) return;
  const uint64_t hash = id_lo[perm[v]] * 11400714819323198549ull;  // FastHasher hyperloglog.rs:4311-4313
  const uint32_t j = (uint32_t)(hash >> 58);
  const uint64_t wv = hash << 6;
  const uint32_t p = (wv == 0 ? 64u : (uint32_t)__clzll((long long)wv)) + 1u;
  uint32_t w[16];
#pragma unroll
  for (int i = 0; i < 16; i++) w[i] = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) if ((int)(j >> 2) == i) w[i] = p << (8 * (j & 3));
#pragma unroll
  for (int q = 0; q < 4; q++) {
    uint4 x = make_uint4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
    r0[v * 4 + q] = x; r1[v * 4 + q] = x;
  }
  size_cache[v] = hll64_size(w, &tab);
  ksum[v] = 0.0; kerr[v] = 0.0;
}
__global__ void k_bm_fill(uint32_t* bm, uint64_t N, uint64_t words) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= words) return;
  uint64_t first = i * 32;
  uint32_t m = 0xFFFFFFFFu;
  if (first + 32 > N) m = (N > first) ? ((1u << (N - first)) - 1u) : 0u;
  bm[i] = m;
}

// Fused exchange: a produced row is stored into the local `new` array AND, over NVLink peer memory, into every
// other rank's replica of it (one writer per row, so plain 16-B stores; the changed bit goes through a
// system-scope atomic OR).  This replaces the per-iteration all-gather: the transfer of a row overlaps the
// gathers of the rows still being computed, and rows that did not change never cross the link.
// sharded handles: destination rows are owned in interleaved blocks of 32 (block b -> rank b % world)
__device__ __forceinline__ bool owned_row(const PeerOut& o, uint32_t row) {
  return o.world <= 1u || ((row >> 5) % o.world) == o.rank;
}
// Subscriber filter: a peer that has no in-edge from `row` never gathers it, so the row is stored only into the
// replicas of the ranks named in its subscriber mask (at C2 size 69 % / 55 % / 42 % of the (row, peer) pairs at
// 2 / 4 / 8 ranks).  A non-subscriber's copy of the row simply stays at its initial value and is never read.
__device__ __forceinline__ void publish_row(uint4* __restrict__ newr, uint32_t* __restrict__ bm_cur, const PeerOut& peers,
                                            uint32_t row, uint32_t sub, uint4 acc, bool write, bool changed) {
  if (write) newr[(uint64_t)row * 4 + sub] = acc;
  if (changed && sub == 0) atomicOr(bm_cur + (row >> 5), 1u << (row & 31));
  if (!write || peers.n == 0) return;
  const uint32_t want = peers.sub ? __ldg(peers.sub + row) : 0xFFFFFFFFu;
  for (int p = 0; p < peers.n; p++) {
    if ((want >> peers.prank[p]) & 1u) peers.newr[p][(uint64_t)row * 4 + sub] = acc;
    // the changed bit is NOT pushed per row: a 32-row block has one owner, so k_publish_bitmap copies the owner's
    // finished bitmap words to the peers with plain stores (3.5 M remote atomics per peer and iteration measured
    // as the bottleneck of the 8-GPU run)
  }
}

__device__ __forceinline__ bool bm_test(const uint32_t* __restrict__ bm, uint32_t v) {
  return (__ldg(bm + (v >> 5)) >> (v & 31)) & 1u;
}

// ---- pull, short rows: 4 lanes per destination row ----------------------------------------------------
template <bool FRONTIER>
__device__ __forceinline__ void quad_rows(uint64_t row, bool live, uint32_t sub, uint32_t lane,
    const uint32_t* __restrict__ row_ptr, const uint32_t* __restrict__ col, uint32_t col_base,
    const uint4* __restrict__ oldr, uint4* __restrict__ newr,
    const uint32_t* __restrict__ bm_prev, uint32_t* __restrict__ bm_cur, const PeerOut& peers) {
  const uint32_t e0 = live ? row_ptr[row] - col_base : 0u, e1 = live ? row_ptr[row + 1] - col_base : 0u;
  const uint4 own = oldr[row * 4 + sub];
  uint4 acc = own;
  // lane `sub` fetches source index e+sub (one 16-B request per quad per 4 edges, prefetched one step ahead)
  // and the quad shares the four indices by shuffle; an out-of-range or (FRONTIER) unchanged source is
  // redirected to the row itself, which is a no-op under max and hits L1.
  const unsigned qmask = 0xFu << (lane & ~3u);
  const uint32_t self = (uint32_t)row;
  uint32_t nxt = (e0 + sub < e1) ? ld_stream_u32(col + e0 + sub) : self;
  for (uint32_t e = e0; e < e1; e += 4) {
    uint32_t mine = nxt;
    nxt = (e + 4 + sub < e1) ? ld_stream_u32(col + e + 4 + sub) : self;
    if (FRONTIER) mine = bm_test(bm_prev, mine) ? mine : self;
    const uint32_t i0 = __shfl_sync(qmask, mine, 0, 4);
    const uint32_t i1 = __shfl_sync(qmask, mine, 1, 4);
    const uint32_t i2 = __shfl_sync(qmask, mine, 2, 4);
    const uint32_t i3 = __shfl_sync(qmask, mine, 3, 4);
    const uint4 v0 = oldr[(uint64_t)i0 * 4 + sub];
    const uint4 v1 = oldr[(uint64_t)i1 * 4 + sub];
    const uint4 v2 = oldr[(uint64_t)i2 * 4 + sub];
    const uint4 v3 = oldr[(uint64_t)i3 * 4 + sub];
    acc = vmax_u8x16(vmax_u8x16(acc, v0), vmax_u8x16(vmax_u8x16(v1, v2), v3));
  }
  const unsigned ball = __ballot_sync(0xffffffffu, ne_u4(acc, own));
  const bool changed = ((ball >> (lane & ~3u)) & 0xFu) != 0u;
  if (!live) return;
  publish_row(newr, bm_cur, peers, (uint32_t)row, sub, acc, changed || bm_test(bm_prev, (uint32_t)row), changed);
}

template <bool FRONTIER>
__global__ void __launch_bounds__(256, 7) k_pull_quad(uint64_t row_begin, uint64_t row_end,
    const uint32_t* __restrict__ row_ptr, const uint32_t* __restrict__ col, uint32_t col_base,
    const uint4* __restrict__ oldr, uint4* __restrict__ newr,
    const uint32_t* __restrict__ bm_prev, uint32_t* __restrict__ bm_cur, const PeerOut peers) {
  const uint64_t gt = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  const uint32_t sub = threadIdx.x & 3;
  const uint32_t lane = threadIdx.x & 31;
  uint64_t row = row_begin + (gt >> 2);
  bool live = row < row_end;
  if (!live) row = row_end - 1;  // keep the warp converged; results of dead quads are discarded
  live = live && owned_row(peers, (uint32_t)row);
  if (__ballot_sync(0xffffffffu, live) == 0u) return;  // none of this warp's rows belongs to this rank
  quad_rows<FRONTIER>(row, live, sub, lane, row_ptr, col, col_base, oldr, newr, bm_prev, bm_cur, peers);
}

// Sharded handles with the fused exchange: the short rows are most of the rows, so this kernel carries most of the
// stores into the peers' replicas -- it is bound by NVLink, not by HBM, while k_pull_warp is the opposite.  This variant
// is a small persistent grid (a few CTAs per SM) that walks the OWNED 32-row blocks only (4 warps per block), so that it can
// sit next to k_pull_warp on a second, higher-priority stream: link traffic of the short rows under the gathers of the long ones.
template <bool FRONTIER>
__global__ void __launch_bounds__(256, 4) k_pull_quad_owned(uint64_t row_begin, uint64_t row_end, uint64_t first_block, uint64_t n_tasks,
    const uint32_t* __restrict__ row_ptr, const uint32_t* __restrict__ col, uint32_t col_base,
    const uint4* __restrict__ oldr, uint4* __restrict__ newr,
    const uint32_t* __restrict__ bm_prev, uint32_t* __restrict__ bm_cur, const PeerOut peers) {
  const uint32_t sub = threadIdx.x & 3, lane = threadIdx.x & 31;
  const uint64_t warps = (uint64_t)gridDim.x * (blockDim.x >> 5);
  for (uint64_t task = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; task < n_tasks; task += warps) {
    const uint64_t block = first_block + (task >> 2) * peers.world;
    uint64_t row = block * 32 + (task & 3) * 8 + (lane >> 2);
    const bool live = row >= row_begin && row < row_end;
    if (!live) row = row_begin;
    if (__ballot_sync(0xffffffffu, live) == 0u) continue;
    quad_rows<FRONTIER>(row, live, sub, lane, row_ptr, col, col_base, oldr, newr, bm_prev, bm_cur, peers);
  }
}

// Experiment (SB200_QUAD2=1, single-rank handles): two adjacent rows per quad, their index loads and gathers interleaved --
// twice the loads in flight per lane for the short-row class, which ncu shows waiting on the long scoreboard (69 % of the
// DRAM peak at 87 % occupancy).  40 registers => 6 CTAs/SM instead of 7.
template <bool FRONTIER>
__global__ void __launch_bounds__(256, 4) k_pull_quad2(uint64_t row_begin, uint64_t row_end,
    const uint32_t* __restrict__ row_ptr, const uint32_t* __restrict__ col,
    const uint4* __restrict__ oldr, uint4* __restrict__ newr,
    const uint32_t* __restrict__ bm_prev, uint32_t* __restrict__ bm_cur) {
  const uint64_t gt = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  const uint32_t sub = threadIdx.x & 3, lane = threadIdx.x & 31;
  const uint64_t pair = gt >> 2;
  uint64_t rowA = row_begin + 2 * pair, rowB = rowA + 1;
  const bool liveA = rowA < row_end, liveB = rowB < row_end;
  if (!liveA) rowA = row_end - 1;
  if (!liveB) rowB = row_end - 1;
  if (__ballot_sync(0xffffffffu, liveA) == 0u) return;
  const uint32_t a0 = liveA ? row_ptr[rowA] : 0u, a1 = liveA ? row_ptr[rowA + 1] : 0u;
  const uint32_t b0 = liveB ? row_ptr[rowB] : 0u, b1 = liveB ? row_ptr[rowB + 1] : 0u;
  uint4 accA = oldr[rowA * 4 + sub], accB = oldr[rowB * 4 + sub];
  const unsigned qmask = 0xFu << (lane & ~3u);
  const uint32_t selfA = (uint32_t)rowA, selfB = (uint32_t)rowB;
  const uint32_t steps = max((a1 - a0 + 3u) >> 2, (b1 - b0 + 3u) >> 2);
  uint32_t nA = (a0 + sub < a1) ? ld_stream_u32(col + a0 + sub) : selfA;
  uint32_t nB = (b0 + sub < b1) ? ld_stream_u32(col + b0 + sub) : selfB;
  for (uint32_t i = 0; i < steps; i++) {
    uint32_t mA = nA, mB = nB;
    const uint32_t ea = a0 + 4 * (i + 1) + sub, eb = b0 + 4 * (i + 1) + sub;
    nA = (ea < a1) ? ld_stream_u32(col + ea) : selfA;
    nB = (eb < b1) ? ld_stream_u32(col + eb) : selfB;
    if (FRONTIER) { mA = bm_test(bm_prev, mA) ? mA : selfA; mB = bm_test(bm_prev, mB) ? mB : selfB; }
    const uint32_t iA0 = __shfl_sync(qmask, mA, 0, 4), iA1 = __shfl_sync(qmask, mA, 1, 4), iA2 = __shfl_sync(qmask, mA, 2, 4), iA3 = __shfl_sync(qmask, mA, 3, 4);
    const uint32_t iB0 = __shfl_sync(qmask, mB, 0, 4), iB1 = __shfl_sync(qmask, mB, 1, 4), iB2 = __shfl_sync(qmask, mB, 2, 4), iB3 = __shfl_sync(qmask, mB, 3, 4);
    const uint4 vA0 = oldr[(uint64_t)iA0 * 4 + sub], vA1 = oldr[(uint64_t)iA1 * 4 + sub], vA2 = oldr[(uint64_t)iA2 * 4 + sub], vA3 = oldr[(uint64_t)iA3 * 4 + sub];
    const uint4 vB0 = oldr[(uint64_t)iB0 * 4 + sub], vB1 = oldr[(uint64_t)iB1 * 4 + sub], vB2 = oldr[(uint64_t)iB2 * 4 + sub], vB3 = oldr[(uint64_t)iB3 * 4 + sub];
    accA = vmax_u8x16(vmax_u8x16(accA, vA0), vmax_u8x16(vmax_u8x16(vA1, vA2), vA3));
    accB = vmax_u8x16(vmax_u8x16(accB, vB0), vmax_u8x16(vmax_u8x16(vB1, vB2), vB3));
  }
  const uint4 ownA = oldr[rowA * 4 + sub], ownB = oldr[rowB * 4 + sub];   // re-read (L1/L2 hit) instead of holding 8 registers across the loop
  const unsigned ballA = __ballot_sync(0xffffffffu, ne_u4(accA, ownA)), ballB = __ballot_sync(0xffffffffu, ne_u4(accB, ownB));
  const bool chA = ((ballA >> (lane & ~3u)) & 0xFu) != 0u, chB = ((ballB >> (lane & ~3u)) & 0xFu) != 0u;
  if (liveA && (chA || bm_test(bm_prev, selfA))) newr[rowA * 4 + sub] = accA;
  if (liveA && chA && sub == 0) atomicOr(bm_cur + (selfA >> 5), 1u << (selfA & 31));
  if (liveB && (chB || bm_test(bm_prev, selfB))) newr[rowB * 4 + sub] = accB;
  if (liveB && chB && sub == 0) atomicOr(bm_cur + (selfB >> 5), 1u << (selfB & 31));
}

// ---- pull, long rows: one warp per <=CHUNK_EDGES work item ---------------------------------------------
template <bool FRONTIER, bool LISTED>
// 8 CTAs/SM (<= 32 registers): at full scale the gathers are DRAM-latency bound and the kernel's speed tracks the
// number of resident warps (36 registers = 7 CTAs measured 11 % slower than 32 registers = 8 CTAs)
__global__ void __launch_bounds__(256, 8) k_pull_warp(uint64_t n_items, uint64_t first_multi_free_item,
    const uint32_t* __restrict__ item_list, const uint32_t* __restrict__ item_row, const uint32_t* __restrict__ item_start, uint32_t warp_row_begin,
    const uint32_t* __restrict__ row_ptr, const uint32_t* __restrict__ col, uint32_t col_base,
    const uint4* __restrict__ oldr, uint4* __restrict__ newr, uint4* __restrict__ partial,
    const uint32_t* __restrict__ bm_prev, uint32_t* __restrict__ bm_cur, const PeerOut peers) {
  uint64_t item = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
  if (item >= n_items) return;  // whole warp exits together
  if (LISTED) item = item_list[item];   // sharded: n_items counts the owned items, listed ascending
  const uint32_t lane = threadIdx.x & 31, sub = lane & 3, q = lane >> 2;
  const uint32_t row = item_row[item];
  if (!owned_row(peers, row)) return;  // sharded: another rank owns this row (warp-uniform)
  const uint32_t chunk = (uint32_t)item - item_start[row - warp_row_begin];
  const uint32_t rs = row_ptr[row] - col_base, re = row_ptr[row + 1] - col_base;
  const uint32_t e0 = rs + chunk * (uint32_t)CHUNK_EDGES;
  const uint32_t e1 = min(e0 + (uint32_t)CHUNK_EDGES, re);
  uint4 acc = make_uint4(0, 0, 0, 0);
  // out-of-range lanes and (FRONTIER) unchanged sources are redirected to the row itself: merging one's own row
  // is a no-op under max and hits L1, so the gather loop is branch-free
  uint32_t nxt = (e0 + lane < e1) ? ld_stream_u32(col + e0 + lane) : row;
  for (uint32_t base = e0; base < e1; base += 32) {
    uint32_t mine = nxt;
    nxt = (base + 32 + lane < e1) ? ld_stream_u32(col + base + 32 + lane) : row;
    if (FRONTIER) mine = bm_test(bm_prev, mine) ? mine : row;
    const uint32_t i0 = __shfl_sync(0xffffffffu, mine, q);
    const uint32_t i1 = __shfl_sync(0xffffffffu, mine, q + 8);
    const uint32_t i2 = __shfl_sync(0xffffffffu, mine, q + 16);
    const uint32_t i3 = __shfl_sync(0xffffffffu, mine, q + 24);
    const uint4 v0 = oldr[(uint64_t)i0 * 4 + sub];
    const uint4 v1 = oldr[(uint64_t)i1 * 4 + sub];
    const uint4 v2 = oldr[(uint64_t)i2 * 4 + sub];
    const uint4 v3 = oldr[(uint64_t)i3 * 4 + sub];
    acc = vmax_u8x16(vmax_u8x16(acc, v0), vmax_u8x16(vmax_u8x16(v1, v2), v3));
  }
#pragma unroll
  for (int off = 4; off < 32; off <<= 1) {
    uint4 o;
    o.x = __shfl_xor_sync(0xffffffffu, acc.x, off); o.y = __shfl_xor_sync(0xffffffffu, acc.y, off);
    o.z = __shfl_xor_sync(0xffffffffu, acc.z, off); o.w = __shfl_xor_sync(0xffffffffu, acc.w, off);
    acc = vmax_u8x16(acc, o);
  }
  if (item < first_multi_free_item) {  // this row spans several items: park the partial maximum
    if (q == 0) partial[item * 4 + sub] = acc;
    return;
  }
  const uint4 own = oldr[(uint64_t)row * 4 + sub];
  acc = vmax_u8x16(acc, own);
  const unsigned ball = __ballot_sync(0xffffffffu, ne_u4(acc, own));
  const bool changed = (ball & 0xFu) != 0u;
  if (q == 0) publish_row(newr, bm_cur, peers, row, sub, acc, changed || bm_test(bm_prev, row), changed);
}

// rows spanning several work items: one warp reduces the parked partials
__global__ void __launch_bounds__(256) k_pull_merge(uint64_t n_rows, const uint32_t* __restrict__ item_start,
    uint32_t warp_row_begin, const uint4* __restrict__ partial, const uint4* __restrict__ oldr,
    uint4* __restrict__ newr, const uint32_t* __restrict__ bm_prev, uint32_t* __restrict__ bm_cur, const PeerOut peers) {
  const uint64_t r = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
  if (r >= n_rows) return;
  const uint32_t lane = threadIdx.x & 31, sub = lane & 3, q = lane >> 2;
  const uint32_t row = warp_row_begin + (uint32_t)r;
  if (!owned_row(peers, row)) return;
  const uint32_t i0 = item_start[r], i1 = item_start[r + 1];
  uint4 acc = make_uint4(0, 0, 0, 0);
  for (uint32_t it = i0 + q; it < i1; it += 8) acc = vmax_u8x16(acc, partial[(uint64_t)it * 4 + sub]);
#pragma unroll
  for (int off = 4; off < 32; off <<= 1) {
    uint4 o;
    o.x = __shfl_xor_sync(0xffffffffu, acc.x, off); o.y = __shfl_xor_sync(0xffffffffu, acc.y, off);
    o.z = __shfl_xor_sync(0xffffffffu, acc.z, off); o.w = __shfl_xor_sync(0xffffffffu, acc.w, off);
    acc = vmax_u8x16(acc, o);
  }
  const uint4 own = oldr[(uint64_t)row * 4 + sub];
  acc = vmax_u8x16(acc, own);
  const unsigned ball = __ballot_sync(0xffffffffu, ne_u4(acc, own));
  const bool changed = (ball & 0xFu) != 0u;
  if (q == 0) publish_row(newr, bm_cur, peers, row, sub, acc, changed || bm_test(bm_prev, row), changed);
}

// ---- push from a small frontier (update_changed_counters, harmonic.rs:75-114) --------------------------
// slot_sum (sharded handles): total of the out-degrees, i.e. the number of push slots of this rank
__global__ void k_frontier_compact(const uint32_t* __restrict__ bm, uint64_t words, const uint32_t* __restrict__ fwd_ptr,
                                   uint32_t* list, uint32_t* outdeg, unsigned long long* counter,
                                   unsigned long long* slot_sum) {
  uint64_t w = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (w >= words) return;
  uint32_t m = bm[w];
  if (!m) return;
  unsigned long long pos = atomicAdd(counter, (unsigned long long)__popc(m));
  unsigned long long sum = 0;
  while (m) {
    int b = __ffs(m) - 1; m &= m - 1;
    uint32_t v = (uint32_t)(w * 32 + b);
    const uint32_t d = fwd_ptr[v + 1] - fwd_ptr[v];
    list[pos] = v; outdeg[pos] = d; pos++; sum += d;
  }
  if (slot_sum && sum) atomicAdd(slot_sum, sum);
}
// refresh the two-iteration-old rows of the frontier nodes; a sharded handle refreshes the rows it owns (the others
// arrive from their owners)
__global__ void k_copy_stale(const uint32_t* __restrict__ list, uint64_t n, const uint4* __restrict__ oldr,
                             uint4* __restrict__ newr, uint32_t world, uint32_t rank) {
  uint64_t gt = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  uint64_t i = gt >> 2;
  if (i >= n) return;
  uint64_t v = list[i];
  if (world > 1 && ((v >> 5) % world) != rank) return;
  newr[v * 4 + (gt & 3)] = oldr[v * 4 + (gt & 3)];
}
// fused exchange after a push: every owned row that was refreshed or changed in this iteration goes to the peers
// (the frontier is small by construction: one quad per OWNED bitmap word, i.e. per 32-row block, walks the set bits)
__global__ void k_publish_rows(const uint32_t* __restrict__ bm_prev, const uint32_t* __restrict__ bm_cur, uint64_t n_rows,
                               const uint4* __restrict__ newr, const PeerOut peers) {
  const uint64_t gt = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  const uint64_t w = (gt >> 2) * peers.world + peers.rank;     // this rank's k-th bitmap word
  if (w * 32 >= n_rows) return;
  uint32_t m = __ldg(bm_prev + w) | __ldg(bm_cur + w);
  while (m) {
    const uint64_t row = w * 32 + (uint64_t)(__ffs(m) - 1);
    m &= m - 1;
    if (row >= n_rows) break;
    const uint4 v = newr[row * 4 + (gt & 3)];
    const uint32_t want = peers.sub ? __ldg(peers.sub + row) : 0xFFFFFFFFu;
    for (int p = 0; p < peers.n; p++) if ((want >> peers.prank[p]) & 1u) peers.newr[p][row * 4 + (gt & 3)] = v;
  }
}
__global__ void __launch_bounds__(256) k_push(const uint32_t* __restrict__ list, const uint32_t* __restrict__ off,
    uint32_t n_front, uint64_t n_slots, const uint32_t* __restrict__ fwd_ptr, const uint32_t* __restrict__ fwd_dst,
    const uint32_t* __restrict__ old32, uint32_t* new32, uint32_t* __restrict__ bm_cur) {
  const uint32_t lane = threadIdx.x & 31, l = lane & 15;
  const uint64_t hw = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 4;
  const uint64_t nhw = ((uint64_t)gridDim.x * blockDim.x) >> 4;
  for (uint64_t s = hw; s < ((n_slots + 1) & ~1ull); s += nhw) {  // both halves of a warp iterate together
    bool changed = false;
    uint32_t v = 0;
    if (s < n_slots) {
      uint32_t lo = 0, hi = n_front;  // last j with off[j] <= s
      while (lo + 1 < hi) { uint32_t mid = (lo + hi) >> 1; if (off[mid] <= s) lo = mid; else hi = mid; }
      const uint32_t u = list[lo];
      v = fwd_dst[fwd_ptr[u] + (uint32_t)(s - off[lo])];
      const uint32_t mine = old32[(uint64_t)u * 16 + l];
      uint32_t* addr = new32 + (uint64_t)v * 16 + l;
      uint32_t cur = *addr;
      uint32_t m = bmax4_7bit(cur, mine);
      while (m != cur) {
        const uint32_t prev = atomicCAS(addr, cur, m);
        if (prev == cur) { changed = true; break; }
        cur = prev; m = bmax4_7bit(cur, mine);
      }
    }
    const unsigned ball = __ballot_sync(0xffffffffu, changed);
    const unsigned half = (lane < 16) ? (ball & 0xFFFFu) : (ball >> 16);
    if (half && l == 0) atomicOr(bm_cur + (v >> 5), 1u << (v & 31));
  }
}

// ---- centrality update for the nodes touched this iteration -------------------------------------------
__global__ void __launch_bounds__(256) k_finalize(uint64_t row_begin, uint64_t row_end,
    const uint4* __restrict__ newr, const uint32_t* __restrict__ bm_prev, const uint32_t* __restrict__ bm_cur,
    uint64_t* __restrict__ size_cache, double* __restrict__ ksum, double* __restrict__ kerr,
    const uint32_t* __restrict__ fwd_ptr, double t_plus_1, unsigned long long* counters, const uint32_t own_world,
    const uint32_t own_rank) {
  __shared__ HllTables tab;
  __shared__ unsigned long long s_cnt[2];
  for (int i = threadIdx.x; i < (int)(sizeof(HllTables) / 8); i += blockDim.x) ((double*)&tab)[i] = ((const double*)&c_tab)[i];
  if (threadIdx.x < 2) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  // persistent grid: the tables are staged once per CTA, each warp then walks whole 32-row blocks -- only the blocks this
  // rank owns, and only those with a bit set in either bitmap word (late iterations touch a few thousand rows)
  const uint32_t lane = threadIdx.x & 31;
  const uint64_t world = own_world > 1 ? own_world : 1;
  const uint64_t warps = (uint64_t)gridDim.x * (blockDim.x >> 5), w0 = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const uint64_t b0 = row_begin >> 5, b1 = (row_end + 31) >> 5;
  const uint64_t first = b0 + ((uint64_t)own_rank + world - b0 % world) % world;
  unsigned long long my_changed = 0, my_out = 0;
  for (uint64_t b = first + w0 * world; b < b1; b += warps * world) {
    const uint32_t wc = __ldg(bm_cur + b), wp = __ldg(bm_prev + b);
    if ((wc | wp) == 0u) continue;
    const uint64_t v = b * 32 + lane;
    const bool bc = (wc >> lane) & 1u, bp = (wp >> lane) & 1u;
    if (v >= row_begin && v < row_end && (bc || bp)) {
      double s = ksum[v], e = kerr[v];
      if (bc) {
        uint32_t w[16];
#pragma unroll
        for (int q = 0; q < 4; q++) {
          const uint4 x = newr[v * 4 + q];
          w[4 * q] = x.x; w[4 * q + 1] = x.y; w[4 * q + 2] = x.z; w[4 * q + 3] = x.w;
        }
        const uint64_t sn = hll64_size(w, &tab);
        const uint64_t so = size_cache[v];
        const uint64_t d = sn >= so ? sn - so : 0ull;  // checked_sub().unwrap_or_default()
        kahan_add(s, e, __ddiv_rn(__ull2double_rn(d), t_plus_1));
        size_cache[v] = sn;
        my_changed += 1;
        if (fwd_ptr) my_out += fwd_ptr[v + 1] - fwd_ptr[v];
      } else {
        kahan_add(s, e, 0.0);  // the reference adds 0/(t+1) to every unchanged node; once is enough
      }
      ksum[v] = s; kerr[v] = e;
    }
  }
  // block reduce the two counters
  for (int off = 16; off; off >>= 1) {
    my_changed += __shfl_down_sync(0xffffffffu, my_changed, off);
    my_out += __shfl_down_sync(0xffffffffu, my_out, off);
  }
  if ((threadIdx.x & 31) == 0 && (my_changed | my_out)) { atomicAdd(&s_cnt[0], my_changed); atomicAdd(&s_cnt[1], my_out); }
  __syncthreads();
  if (threadIdx.x == 0 && (s_cnt[0] | s_cnt[1])) { atomicAdd(counters + 0, s_cnt[0]); atomicAdd(counters + 1, s_cnt[1]); }
}

// ---- result / debug gathers ------------------------------------------------------------------------------
__global__ void k_result_flags(const uint32_t* __restrict__ inv, const double* __restrict__ ksum, uint64_t N,
                               uint64_t row_begin, uint64_t row_end, double norm, uint32_t* flag, double* val,
                               const uint32_t own_world, const uint32_t own_rank) {
  uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (r >= N) return;
  const uint32_t v = inv[r];
  double c = ksum[v];
  const bool keep = (c > 0.0) && v >= row_begin && v < row_end &&  // normalize_centralities harmonic.rs:178-195
                    (own_world <= 1 || ((v >> 5) % own_world) == own_rank);
  c = __ddiv_rn(c, norm);
  if (isinf(c) || isnan(c)) c = 0.0;
  flag[r] = keep ? 1u : 0u;
  val[r] = c;
}
__global__ void k_result_scatter(const uint32_t* __restrict__ flag, const uint32_t* __restrict__ pos,
                                 const double* __restrict__ val, const uint64_t* __restrict__ id_lo,
                                 const uint64_t* __restrict__ id_hi, uint64_t N, uint64_t cap, uint64_t* out_lo,
                                 uint64_t* out_hi, double* out_c) {
  uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (r >= N || !flag[r]) return;
  uint32_t p = pos[r];
  if (p >= cap) return;
  out_lo[p] = id_lo[r]; out_hi[p] = id_hi[r]; out_c[p] = val[r];
}
__global__ void k_gather_regs(const uint32_t* __restrict__ inv, const uint4* __restrict__ regs, uint64_t first,
                              uint64_t count, uint4* out) {
  uint64_t gt = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  uint64_t i = gt >> 2;
  if (i >= count) return;
  out[i * 4 + (gt & 3)] = regs[(uint64_t)inv[first + i] * 4 + (gt & 3)];
}
__global__ void k_gather_f64x2(const uint32_t* __restrict__ inv, const double* __restrict__ a, const double* __restrict__ b,
                               uint64_t first, uint64_t count, double* oa, double* ob) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= count) return;
  uint32_t v = inv[first + i];
  oa[i] = a[v]; ob[i] = b[v];
}

static double env_f(const char* name, double dflt) {
  const char* s = getenv(name);
  return s ? atof(s) : dflt;
}

int hb_alloc_state(sb200_graph* g) {
  g->dense_frac = env_f("SB200_DENSE_FRAC", g->dense_frac);
  g->push_div = env_f("SB200_PUSH_DIV", g->push_div);
  g->force_mode = (int)env_f("SB200_FORCE_MODE", (double)g->force_mode);
  const uint64_t N = g->N;
  const uint64_t words = (N + 31) / 32;
  SB_TRY(load_tables(g->device));
  SB_TRY(g->regs[0].alloc(std::max<uint64_t>(N, 1) * 64));
  SB_TRY(g->regs[1].alloc(std::max<uint64_t>(N, 1) * 64));
  SB_TRY(g->bm[0].alloc(words + 1)); SB_TRY(g->bm[1].alloc(words + 1));
  SB_TRY(g->size_cache.alloc(std::max<uint64_t>(N, 1)));
  SB_TRY(g->kahan_sum.alloc(std::max<uint64_t>(N, 1))); SB_TRY(g->kahan_err.alloc(std::max<uint64_t>(N, 1)));
  SB_TRY(g->counters.alloc(8));
  if (g->world > 1) {   // sync page of the device-side barrier (plain cudaMalloc: exported through CUDA IPC)
    SB_TRY(g->sync_page.alloc(SYNC_SLOTS));
    SB_CUDA(cudaMemset(g->sync_page.p, 0, SYNC_SLOTS * sizeof(unsigned long long)));
    g->publish_all = env_flag("SB200_PUBLISH_ALL", false);
  }
  if (!g->h_counters) SB_CUDA(cudaMallocHost((void**)&g->h_counters, 8 * sizeof(unsigned long long)));
  // Rows are ordered by in-degree, so the head of the register array holds the hubs -- on a power-law graph also the
  // most-gathered sources.  The first SB200_L2_PERSIST_MB (default 32; 0 = off) MB of the array being READ are pinned
  // as persisting L2 lines for the pull kernels (stream access-policy window, re-pointed at the `old` array every
  // iteration), so the streaming col/row traffic cannot evict them.
  const double mb = env_f("SB200_L2_PERSIST_MB", 32.0);  // measured at C2: 0 -> 57.5, 32 -> 54.1, 64 -> 56.2, 96 -> 60.0 ms/step
  g->l2_window_bytes = 0;
  if (mb > 0) {
    cudaDeviceProp prop;
    SB_CUDA(cudaGetDeviceProperties(&prop, g->device));
    uint64_t want = (uint64_t)(mb * 1048576.0);
    want = std::min<uint64_t>(want, (uint64_t)std::max(prop.persistingL2CacheMaxSize, 0));
    want = std::min<uint64_t>(want, (uint64_t)std::max(prop.accessPolicyMaxWindowSize, 0));
    want = std::min<uint64_t>(want, N * 64);
    if (want) { SB_CUDA(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want)); g->l2_window_bytes = want; }
  }
  return SB200_OK;
}

int hb_reset(sb200_graph* g) {
  cudaStream_t s = g->stream;
  const uint64_t N = g->N, words = (N + 31) / 32;
  g->cur = 0; g->bcur = 0; g->t = 0; g->has_changes = N > 0; g->exchange_pending = false;
  g->n_changed_prev = N; g->frontier_edges_prev = g->E_kept;
  if (N == 0) return SB200_OK;
  SB_LAUNCH(k_hb_init, div_up(N, 256), 256, 0, s, g->id_lo.p, g->perm.p, N, (uint4*)g->regs[0].p, (uint4*)g->regs[1].p,
            g->size_cache.p, g->kahan_sum.p, g->kahan_err.p);
  SB_CHECK_LAUNCH();
  SB_LAUNCH(k_bm_fill, div_up(words, 256), 256, 0, s, g->bm[0].p, N, words);  // changed_nodes filled, harmonic.rs:223-225
  SB_CHECK_LAUNCH();
  SB_CUDA(cudaMemsetAsync(g->bm[1].p, 0, (words + 1) * 4, s));
  SB_CUDA(cudaStreamSynchronize(s));
  return SB200_OK;
}

#define PROF_BEGIN(g, fam) do { if ((g)->profiling) { SB_CUDA(cudaEventRecord((g)->prof_ev[fam][0], (g)->stream)); } } while (0)
#define PROF_END(g, fam, bytes) do { if ((g)->profiling) { SB_CUDA(cudaEventRecord((g)->prof_ev[fam][1], (g)->stream)); \
    (g)->prof_used[fam] = true; (g)->prof_step_bytes[fam] = (double)(bytes); } } while (0)

// publish targets of the iteration being launched: the peers' copies of the `new` register array / `cur` bitmap
static PeerOut make_peer_out(const sb200_graph* g, bool with_targets) {
  PeerOut po;
  memset(&po, 0, sizeof(po));
  po.world = (uint32_t)g->world; po.rank = (uint32_t)g->rank;
  if (with_targets && g->p2p) {
    po.n = g->n_peers;
    po.sub = (g->publish_all || !g->sub_mask.p) ? nullptr : g->sub_mask.p;
    for (int p = 0; p < g->n_peers; p++) {
      po.newr[p] = (uint4*)g->peer_regs[g->cur ^ 1][p]; po.bmc[p] = (uint32_t*)g->peer_bm[g->bcur ^ 1][p];
      po.prank[p] = (uint8_t)g->peer_rank[p];
    }
  }
  return po;
}

template <bool FRONTIER>
static int launch_pull(sb200_graph* g, const uint4* oldr, uint4* newr, const uint32_t* bmp, uint32_t* bmc) {
  cudaStream_t s = g->stream;
  const PeerOut po = make_peer_out(g, true);
  const int FW = FRONTIER ? sb200_graph::F_PULL_WARP_FRONT : sb200_graph::F_PULL_WARP_DENSE;
  const int FQ = FRONTIER ? sb200_graph::F_PULL_QUAD_FRONT : sb200_graph::F_PULL_QUAD_DENSE;
  const double per_edge = FRONTIER ? 4.0 : 68.0;  // col index (+ the 64-B gather when every source is read)
  // fused exchange: short rows on the side stream, next to the long-row kernel (see k_pull_quad_owned)
  if (g->opt_side_ctas < 0) {
    // measured on C2 (profiles/r02_trip9_*gpu.log, r02_trip10_8gpu_sweep.log): unicast peer stores gain from the side stream at
    // 2 and 4 ranks (35.0 -> 33.1, 22.1 -> 17.5 ms) and lose at 8 (27.0 -> 32.6 ms); one NVSwitch multicast target gains (19.5 -> 16.9)
    const bool multicast_target = g->n_peers < g->world - 1;
    g->opt_side_ctas = (int)env_f("SB200_QUAD_SIDE_CTAS", (multicast_target || g->world <= 4) ? 2.0 : 0.0);
  }
  const int side_ctas = g->opt_side_ctas;
  const bool side_quad = side_ctas > 0 && g->world > 1 && g->p2p && g->n_peers > 0 && g->n_items && g->quad_row_end > g->quad_row_begin;
  if (side_quad) {
    if (!g->side_stream) {
      int lo = 0, hi = 0;
      SB_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
      SB_CUDA(cudaStreamCreateWithPriority(&g->side_stream, cudaStreamNonBlocking, hi));
      SB_CUDA(cudaEventCreateWithFlags(&g->ev_fork, cudaEventDisableTiming));
      SB_CUDA(cudaEventCreateWithFlags(&g->ev_join, cudaEventDisableTiming));
      SB_CUDA(cudaEventCreate(&g->side_prof[0])); SB_CUDA(cudaEventCreate(&g->side_prof[1]));
    }
    if (!g->sm_count) SB_CUDA(cudaDeviceGetAttribute(&g->sm_count, cudaDevAttrMultiProcessorCount, g->device));
    const int sm_count = g->sm_count;
    const uint64_t world = (uint64_t)g->world, b0 = g->quad_row_begin >> 5, b1 = (g->quad_row_end - 1) >> 5;
    const uint64_t first = b0 + ((uint64_t)g->rank + world - b0 % world) % world;
    const uint64_t n_tasks = first <= b1 ? ((b1 - first) / world + 1) * 4 : 0;
    SB_CUDA(cudaEventRecord(g->ev_fork, s));
    SB_CUDA(cudaStreamWaitEvent(g->side_stream, g->ev_fork, 0));
    if (g->l2_window_bytes) {
      cudaStreamAttrValue a;
      memset(&a, 0, sizeof(a));
      a.accessPolicyWindow.base_ptr = (void*)oldr; a.accessPolicyWindow.num_bytes = g->l2_window_bytes;
      a.accessPolicyWindow.hitRatio = 1.0f; a.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
      a.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
      SB_CUDA(cudaStreamSetAttribute(g->side_stream, cudaStreamAttributeAccessPolicyWindow, &a));
    }
    if (g->profiling) SB_CUDA(cudaEventRecord(g->side_prof[0], g->side_stream));
    if (getenv("SB200_DEBUG_SIDE")) fprintf(stderr, "[sb200] side quad: rank %d tasks %llu\n", g->rank, (unsigned long long)n_tasks);
    if (n_tasks) {
      const unsigned grid = (unsigned)std::min<uint64_t>(div_up(n_tasks, 8), (uint64_t)sm_count * (uint64_t)side_ctas);
      SB_LAUNCH(k_pull_quad_owned<FRONTIER>, grid, 256, 0, g->side_stream, g->quad_row_begin, g->quad_row_end, first, n_tasks,
                g->row_ptr.p, g->col.p, g->col_base, oldr, newr, bmp, bmc, po);
      SB_CHECK_LAUNCH();
    }
    if (g->profiling) {
      SB_CUDA(cudaEventRecord(g->side_prof[1], g->side_stream));
      g->side_prof_used = true; g->side_prof_family = FQ;
      g->prof_step_bytes[FQ] = g->own_frac * (per_edge * (double)g->E_quad + 68.0 * (double)(g->quad_row_end - g->quad_row_begin));
    }
  }
  if (g->n_items) {
    PROF_BEGIN(g, FW);
    if (g->opt_owned_list < 0) g->opt_owned_list = env_flag("SB200_OWNED_ITEMS", true) ? 1 : 0;
    const bool listed = g->opt_owned_list > 0 && g->world > 1 && g->owned_items.p;
    const uint64_t n_launch = listed ? g->n_owned_items : g->n_items;
    auto kern = listed ? k_pull_warp<FRONTIER, true> : k_pull_warp<FRONTIER, false>;
    if (n_launch)
    SB_LAUNCH(kern, div_up(n_launch * 32, 256), 256, 0, s, n_launch, g->n_multi_items,
              listed ? g->owned_items.p : (const uint32_t*)nullptr, g->item_row.p, g->item_start.p, (uint32_t)g->warp_row_begin, g->row_ptr.p, g->col.p, g->col_base, oldr, newr,
              g->partial.p, bmp, bmc, po);
    SB_CHECK_LAUNCH();
    PROF_END(g, FW, g->own_frac * (per_edge * (double)g->E_warp + 68.0 * (double)(g->warp_row_end - g->warp_row_begin - g->n_multi_rows)));
  }
  if (g->n_multi_rows) {
    PROF_BEGIN(g, sb200_graph::F_PULL_MERGE);
    SB_LAUNCH(k_pull_merge, div_up(g->n_multi_rows * 32, 256), 256, 0, s, g->n_multi_rows, g->item_start.p,
              (uint32_t)g->warp_row_begin, g->partial.p, oldr, newr, bmp, bmc, po);
    SB_CHECK_LAUNCH();
    PROF_END(g, sb200_graph::F_PULL_MERGE, 64.0 * (double)g->n_multi_items + 68.0 * (double)g->n_multi_rows);
  }
  const uint64_t nq = g->quad_row_end - g->quad_row_begin;
  if (side_quad) { SB_CUDA(cudaEventRecord(g->ev_join, g->side_stream)); SB_CUDA(cudaStreamWaitEvent(s, g->ev_join, 0)); }
  else if (nq) {
    PROF_BEGIN(g, FQ);
    static const bool quad2 = env_flag("SB200_QUAD2", false);
    if (quad2 && g->world == 1 && g->col_base == 0)
      SB_LAUNCH(k_pull_quad2<FRONTIER>, div_up(((nq + 1) / 2) * 4, 256), 256, 0, s, g->quad_row_begin, g->quad_row_end, g->row_ptr.p,
                g->col.p, oldr, newr, bmp, bmc);
    else
    SB_LAUNCH(k_pull_quad<FRONTIER>, div_up(nq * 4, 256), 256, 0, s, g->quad_row_begin, g->quad_row_end, g->row_ptr.p,
              g->col.p, g->col_base, oldr, newr, bmp, bmc, po);
    SB_CHECK_LAUNCH();
    PROF_END(g, FQ, g->own_frac * (per_edge * (double)g->E_quad + 68.0 * (double)nq));
  }
  return SB200_OK;
}

static int run_push(sb200_graph* g, const uint4* oldr, uint4* newr, const uint32_t* bmp, uint32_t* bmc) {
  cudaStream_t s = g->stream;
  const uint64_t N = g->N, words = (N + 31) / 32;
  const uint64_t nf = g->n_changed_prev;
  if (nf == 0) return SB200_OK;  // empty frontier: a converged state is a fixed point
  if (g->frontier_list.n < nf + 1) { SB_TRY(g->frontier_list.alloc(nf + 1 + (nf >> 2))); }
  if (g->frontier_off.n < 2 * (nf + 1)) { SB_TRY(g->frontier_off.alloc(2 * (nf + 1) + (nf >> 1))); }
  uint32_t* outdeg = g->frontier_off.p + (g->frontier_off.n / 2);
  SB_CUDA(cudaMemsetAsync(g->counters.p + 2, 0, 2 * sizeof(unsigned long long), s));
  SB_LAUNCH(k_frontier_compact, div_up(words, 256), 256, 0, s, bmp, words, g->fwd_ptr.p, g->frontier_list.p, outdeg,
            g->counters.p + 2, g->world > 1 ? g->counters.p + 3 : (unsigned long long*)nullptr);
  SB_CHECK_LAUNCH();
  uint64_t slots = g->frontier_edges_prev;
  if (g->world > 1) {   // this rank's share of the frontier's out-edges is not known from the previous step
    unsigned long long h = 0;
    SB_CUDA(cudaMemcpyAsync(&h, g->counters.p + 3, sizeof(h), cudaMemcpyDeviceToHost, s));
    SB_CUDA(cudaStreamSynchronize(s));
    slots = h;
  }
  size_t need = 0;
  SB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, need, outdeg, g->frontier_off.p, (int64_t)nf, s));
  if (g->cub_tmp.n < need) SB_TRY(g->cub_tmp.alloc(need + 256));
  SB_CUDA(cub::DeviceScan::ExclusiveSum(g->cub_tmp.p, need, outdeg, g->frontier_off.p, (int64_t)nf, s));
  g_launches.fetch_add(2, std::memory_order_relaxed);
  SB_LAUNCH(k_copy_stale, div_up(nf * 4, 256), 256, 0, s, g->frontier_list.p, nf, oldr, newr, (uint32_t)g->world, (uint32_t)g->rank);
  SB_CHECK_LAUNCH();
  if (slots) {
    unsigned grid = (unsigned)std::min<uint64_t>(div_up(slots * 16, 256), 148u * 16u);
    PROF_BEGIN(g, sb200_graph::F_PUSH);
    SB_LAUNCH(k_push, grid, 256, 0, s, g->frontier_list.p, g->frontier_off.p, (uint32_t)nf, slots, g->fwd_ptr.p,
              g->fwd_dst.p, (const uint32_t*)oldr, (uint32_t*)newr, bmc);
    SB_CHECK_LAUNCH();
    PROF_END(g, sb200_graph::F_PUSH, 132.0 * (double)slots);
  }
  if (g->p2p && g->n_peers > 0) {
    const PeerOut po = make_peer_out(g, true);
    SB_LAUNCH(k_publish_rows, div_up(div_up(words, (uint64_t)g->world) * 4, 256), 256, 0, s, bmp, bmc, N, (const uint4*)newr, po);
    SB_CHECK_LAUNCH();
  }
  return SB200_OK;
}

// fused exchange: each rank owns whole 32-row blocks, i.e. whole bitmap words; publish the owned words to the peers
__global__ void k_publish_bitmap(const uint32_t* __restrict__ bmc, uint64_t words, const PeerOut peers) {
  const uint64_t w = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (w >= words || (w % peers.world) != peers.rank) return;
  const uint32_t v = bmc[w];
  for (int p = 0; p < peers.n; p++) peers.bmc[p][w] = v;
}

// out-edges of the nodes in a changed bitmap (needed once, right after the lazy source-major CSR build)
__global__ void k_frontier_out_edges(const uint32_t* __restrict__ bm, uint64_t words, const uint32_t* __restrict__ fwd_ptr,
                                     unsigned long long* counter) {
  const uint64_t w = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  unsigned long long acc = 0;
  if (w < words) {
    uint32_t m = bm[w];
    while (m) { const int b = __ffs(m) - 1; m &= m - 1; const uint64_t v = w * 32 + b; acc += fwd_ptr[v + 1] - fwd_ptr[v]; }
  }
  for (int o = 16; o; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0 && acc) atomicAdd(counter, acc);
}

// ---- device-side barrier + changed-count sum across the ranks of one box ------------------------------------------------
// One warp.  Lane p stores (epoch, count) into slot [epoch & 1][my rank] of peer p's sync page (release, system scope,
// after a system fence that orders this rank's earlier peer stores -- the rows and bitmap words of the iteration --
// before the flag); lane r then spins (acquire) on slot [epoch & 1][r] of the LOCAL page until rank r's flag of this
// epoch has arrived, and the warp sums the counts.  Two parities: a rank can be at most one barrier ahead of the
// slowest one, so the slot of epoch e is rewritten (epoch e + 2) only after everybody has left barrier e + 1, i.e. has
// long read it.  A watchdog (globaltimer) turns a missing peer into an error instead of a hung GPU.
__device__ __forceinline__ void st_release_sys_u64(unsigned long long* p, unsigned long long v) {
#ifndef SB200_EMU
  asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
#else
  *(volatile unsigned long long*)p = v;
#endif
}
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long* p) {
#ifndef SB200_EMU
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
#else
  return *(const volatile unsigned long long*)p;
#endif
}
__device__ __forceinline__ unsigned long long global_ns() {
#ifndef SB200_EMU
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
#else
  return 0ull;
#endif
}
constexpr unsigned long long SYNC_COUNT_MASK = (1ull << 40) - 1;
__global__ void __launch_bounds__(32) k_barrier_count(const SyncView sv, unsigned long long epoch, const unsigned long long* __restrict__ count_in,
                                                      unsigned long long* out /* [0] total, [1] timed-out ranks */, unsigned long long timeout_ns) {
  const int lane = threadIdx.x;
  const unsigned long long e24 = epoch & 0xFFFFFFull;
  const unsigned long long cnt = count_in ? (*count_in & SYNC_COUNT_MASK) : 0ull;
  const unsigned long long val = (e24 << 40) | cnt;
  const uint32_t slot = (uint32_t)(epoch & 1ull) * 64u + sv.rank;
  __threadfence_system();
  if (lane < sv.n) st_release_sys_u64(sv.peer[lane] + slot, val);
  if (lane == sv.n) st_release_sys_u64(sv.local + slot, val);
  unsigned long long got = 0;
  bool ok = true;
  if (lane < (int)sv.world) {
    const unsigned long long* p = sv.local + (uint32_t)(epoch & 1ull) * 64u + lane;
    const unsigned long long t0 = global_ns();
    for (;;) {
      got = ld_acquire_sys_u64(p);
      if ((got >> 40) == e24) break;
      if (global_ns() - t0 > timeout_ns) { ok = false; got = 0; break; }
#ifdef SB200_EMU
      ok = false; got = 0; break;  // the emulator runs one rank at a time: a peer's flag cannot arrive while we wait
#endif
    }
  }
  unsigned long long sum = (lane < (int)sv.world && ok) ? (got & SYNC_COUNT_MASK) : 0ull;
  for (int o = 16; o; o >>= 1) sum += __shfl_down_sync(0xffffffffu, sum, o);
  const unsigned bad = __ballot_sync(0xffffffffu, !ok);
  if (lane == 0) { out[0] = sum; out[1] = (unsigned long long)__popc(bad); }
}

static SyncView make_sync_view(const sb200_graph* g) {
  SyncView sv;
  memset(&sv, 0, sizeof(sv));
  sv.local = g->sync_page.p; sv.n = g->n_peers; sv.world = (uint32_t)g->world; sv.rank = (uint32_t)g->rank;
  for (int p = 0; p < g->n_peers; p++) sv.peer[p] = (unsigned long long*)g->peer_sync[p];
  return sv;
}
// stand-alone barrier (before the first step of a run: every replica must be initialised before a peer may write into it)
int hb_barrier(sb200_graph* g) {
  cudaStream_t s = g->stream;
  if (!g->sync_page.p || g->n_peers != g->world - 1) SB_FAIL(SB200_ESTATE, "device barrier needs the sync pages of all %d peers", g->world - 1);
  for (int p = 0; p < g->n_peers; p++) if (!g->peer_sync[p]) SB_FAIL(SB200_ESTATE, "peer %d exported no sync page", p);
  g->sync_epoch++;
  SB_LAUNCH(k_barrier_count, 1, 32, 0, s, make_sync_view(g), (unsigned long long)g->sync_epoch, (const unsigned long long*)nullptr,
            g->counters.p + 5, 20000000000ull);
  SB_CHECK_LAUNCH();
  SB_CUDA(cudaMemcpyAsync(g->h_counters + 5, g->counters.p + 5, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
  SB_CUDA(cudaStreamSynchronize(s));
  if (g->h_counters[6]) SB_FAIL(SB200_ESTATE, "inter-rank barrier timed out: %llu rank(s) did not arrive", (unsigned long long)g->h_counters[6]);
  return SB200_OK;
}

// One iteration = launch (everything queued on the handle's stream, nothing waited for) + finish (wait, read the
// counters, flip the ping-pong).  hb_step does both; the group / sharded run loops launch all ranks before they wait.
// with_barrier: the changed counts are summed across the ranks on the device (k_barrier_count) and the step finishes
// with the global count already applied (no sb200_hyperball_exchange_done needed).
int hb_step_launch(sb200_graph* g, bool with_barrier) {
  NvtxRange nvtx("sb200 hyperball iteration (launch)");
  cudaStream_t s = g->stream;
  const uint64_t N = g->N, words = (N + 31) / 32;
  if (g->step_in_flight) SB_FAIL(SB200_ESTATE, "the previous step has not been finished");
  if (g->exchange_pending) SB_FAIL(SB200_ESTATE, "sb200_hyperball_exchange_done() must be called between steps of a sharded handle");
  g->step_with_barrier = with_barrier; g->step_mode = 0;
  if (N == 0) { g->step_in_flight = true; return SB200_OK; }
  const uint4* oldr = (const uint4*)g->regs[g->cur].p;
  uint4* newr = (uint4*)g->regs[g->cur ^ 1].p;
  const uint32_t* bmp = g->bm[g->bcur].p;
  uint32_t* bmc = g->bm[g->bcur ^ 1].p;
  const double dense_frac = g->dense_frac, push_div = g->push_div;
  const int force_mode = g->force_mode;
  int mode;
  const double E = (double)std::max<uint64_t>(g->E_kept, 1);
  if (g->world == 1) {
    // the source-major CSR is built lazily: when a reused handle (or a pinned push policy) first meets a frontier
    // of < N/64 nodes.  A handle that computes one centrality and is dropped never pays for it.
    if (!g->has_fwd && ((g->reuse > 0 && g->t > 0 && (double)g->n_changed_prev * 64.0 <= (double)N) || force_mode == 2)) {
      SB_TRY(build_fwd_csr(g));
      SB_CUDA(cudaMemsetAsync(g->counters.p + 4, 0, sizeof(unsigned long long), s));
      SB_LAUNCH(k_frontier_out_edges, div_up(words, 256), 256, 0, s, bmp, words, g->fwd_ptr.p, g->counters.p + 4);
      SB_CHECK_LAUNCH();
      unsigned long long fe0 = 0;
      SB_CUDA(cudaMemcpyAsync(&fe0, g->counters.p + 4, sizeof(fe0), cudaMemcpyDeviceToHost, s));
      SB_CUDA(cudaStreamSynchronize(s));
      g->frontier_edges_prev = fe0;
    }
    if (g->has_fwd) {
      const double fe = (double)g->frontier_edges_prev;
      if (fe >= dense_frac * E) mode = 0;
      else if (fe * push_div <= E) mode = 2;
      else mode = 1;
    } else {
      mode = ((double)g->n_changed_prev >= 0.75 * (double)N) ? 0 : 1;
    }
  } else {
    // sharded handles: the same lazy rule for the source-major CSR (of the owned rows); every rank sees the same
    // global changed count, so all ranks switch together (SB200_SHARDED_PUSH=0 keeps them on the pull kernels)
    // Thresholds from the 2-GPU runs on C2 (profiles/r02_trip3_2gpu.log): the iteration after 16.6 M of 27.8 M nodes changed
    // costs 5.0 ms as a dense pull and ~3.3 ms frontier-filtered (same 0.75 N rule as a single rank without forward CSR);
    // the one after 0.49 M changed costs 3.5 ms as a frontier pull over all local edges but < 1 ms as a push.
    static const bool auto_push = env_flag("SB200_SHARDED_PUSH", true);
    const bool tiny = (double)g->n_changed_prev * 16.0 <= (double)N;
    if (!g->has_fwd && ((auto_push && g->reuse > 0 && g->t > 0 && tiny) || force_mode == 2)) SB_TRY(build_fwd_csr(g));
    if (g->has_fwd && tiny) mode = 2;
    else mode = ((double)g->n_changed_prev >= 0.75 * (double)N) ? 0 : 1;
  }
  if (force_mode >= 0 && (force_mode < 2 || g->has_fwd)) mode = force_mode;
  g->step_mode = mode;
  SB_CUDA(cudaEventRecord(g->ev0, s));
  // peers write their changed bits straight into this rank's bitmap, so with the fused exchange it is cleared
  // at the END of the previous step (before the inter-step barrier), never at the start of this one
  if (!g->p2p) SB_CUDA(cudaMemsetAsync(bmc, 0, (words + 1) * 4, s));
  SB_CUDA(cudaMemsetAsync(g->counters.p, 0, 8 * sizeof(unsigned long long), s));
  if (g->l2_window_bytes && mode != 2) {
    cudaStreamAttrValue a;
    memset(&a, 0, sizeof(a));
    a.accessPolicyWindow.base_ptr = (void*)oldr; a.accessPolicyWindow.num_bytes = g->l2_window_bytes;
    a.accessPolicyWindow.hitRatio = 1.0f; a.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    a.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    SB_CUDA(cudaStreamSetAttribute(s, cudaStreamAttributeAccessPolicyWindow, &a));
  }
  if (mode == 0) SB_TRY(launch_pull<false>(g, oldr, newr, bmp, bmc));
  else if (mode == 1) SB_TRY(launch_pull<true>(g, oldr, newr, bmp, bmc));
  else SB_TRY(run_push(g, oldr, newr, bmp, bmc));
  if (g->p2p && g->n_peers > 0) {
    const PeerOut po = make_peer_out(g, true);
    SB_LAUNCH(k_publish_bitmap, div_up(words, 256), 256, 0, s, bmc, words, po);
    SB_CHECK_LAUNCH();
  }
  const uint64_t nrows = g->row_end - g->row_begin;
  if (nrows) {
    PROF_BEGIN(g, sb200_graph::F_FINALIZE);
    if (!g->sm_count) SB_CUDA(cudaDeviceGetAttribute(&g->sm_count, cudaDevAttrMultiProcessorCount, g->device));
    const unsigned fin_grid = (unsigned)std::min<uint64_t>(div_up(div_up(nrows, (uint64_t)std::max(g->world, 1)), 256) + 1, (uint64_t)g->sm_count * 8);
    SB_LAUNCH(k_finalize, fin_grid, 256, 0, s, g->row_begin, g->row_end, newr, bmp, bmc, g->size_cache.p,
              g->kahan_sum.p, g->kahan_err.p, g->has_fwd ? g->fwd_ptr.p : (const uint32_t*)nullptr, (double)(g->t + 1),
              g->counters.p, (uint32_t)g->world, (uint32_t)g->rank);
    SB_CHECK_LAUNCH();
    PROF_END(g, sb200_graph::F_FINALIZE, 0.25 * (double)nrows);  // 2 bitmap bits/row; + 112 B per changed row below
  }
  if (g->p2p) SB_CUDA(cudaMemsetAsync((void*)bmp, 0, (words + 1) * 4, s));  // next step's `cur` bitmap
  if (with_barrier) {
    if (!g->sync_page.p || g->n_peers != g->world - 1) SB_FAIL(SB200_ESTATE, "device barrier needs the sync pages of all %d peers", g->world - 1);
    g->sync_epoch++;
    SB_LAUNCH(k_barrier_count, 1, 32, 0, s, make_sync_view(g), (unsigned long long)g->sync_epoch, (const unsigned long long*)g->counters.p,
              g->counters.p + 5, 20000000000ull);
    SB_CHECK_LAUNCH();
  }
  SB_CUDA(cudaMemcpyAsync(g->h_counters, g->counters.p, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
  SB_CUDA(cudaEventRecord(g->ev1, s));
  g->step_in_flight = true;
  return SB200_OK;
}

int hb_step_finish(sb200_graph* g, sb200_iter_stats* st) {
  NvtxRange nvtx("sb200 hyperball iteration (wait)");
  cudaStream_t s = g->stream;
  if (!g->step_in_flight) SB_FAIL(SB200_ESTATE, "no step in flight");
  g->step_in_flight = false;
  if (g->N == 0) { g->has_changes = false; if (st) memset(st, 0, sizeof(*st)); return SB200_OK; }
  SB_CUDA(cudaStreamSynchronize(s));
  float ms = 0; cudaEventElapsedTime(&ms, g->ev0, g->ev1);
  if (g->profiling) {
    g->prof_step_bytes[sb200_graph::F_FINALIZE] += 112.0 * (double)g->h_counters[0];
    for (int f = 0; f < sb200_graph::F_COUNT; f++) if (g->prof_used[f]) {
      float kms = 0; cudaEventElapsedTime(&kms, g->prof_ev[f][0], g->prof_ev[f][1]);
      g->prof_launches[f]++; g->prof_ms[f] += kms; g->prof_bytes[f] += g->prof_step_bytes[f];
      g->prof_used[f] = false;
    }
    if (g->side_prof_used) {
      const int f = g->side_prof_family;
      float kms = 0; cudaEventElapsedTime(&kms, g->side_prof[0], g->side_prof[1]);
      g->prof_launches[f]++; g->prof_ms[f] += kms; g->prof_bytes[f] += g->prof_step_bytes[f];
      g->side_prof_used = false;
    }
  }
  if (st) {
    st->t = g->t; st->mode = (uint32_t)g->step_mode; st->n_changed = g->h_counters[0];
    st->edges_active = g->frontier_edges_prev; st->ms = ms;
  }
  g->cur ^= 1; g->bcur ^= 1; g->t += 1;
  if (g->world == 1) {
    g->n_changed_prev = g->h_counters[0];
    g->frontier_edges_prev = g->h_counters[1];
    g->has_changes = g->h_counters[0] != 0;
  } else if (g->step_with_barrier) {
    if (g->h_counters[6]) SB_FAIL(SB200_ESTATE, "inter-rank barrier timed out in iteration %u: %llu rank(s) did not arrive", g->t - 1, (unsigned long long)g->h_counters[6]);
    g->n_changed_prev = g->h_counters[5];   // the global count, summed on the device
    g->has_changes = g->h_counters[5] != 0;
  } else {
    g->n_changed_prev = g->h_counters[0];
    g->exchange_pending = true;
  }
  return SB200_OK;
}

int hb_step(sb200_graph* g, sb200_iter_stats* st) {
  SB_TRY(hb_step_launch(g, false));
  return hb_step_finish(g, st);
}

int hb_result(sb200_graph* g, uint64_t* id_lo, uint64_t* id_hi, double* cent, uint64_t cap, uint64_t* len) {
  NvtxRange nvtx("sb200 hyperball result");
  cudaStream_t s = g->stream;
  const uint64_t N = g->N;
  if (N == 0) { *len = 0; return SB200_OK; }
  PoolScope scope(getenv("SB200_NO_POOL") ? nullptr : s);  // temporaries from the stream-ordered pool, freed on `s`
  const auto tp0 = std::chrono::steady_clock::now();
  DevBuf<uint32_t> flag, pos; DevBuf<double> val;
  SB_TRY(flag.alloc(N + 1)); SB_TRY(pos.alloc(N + 1)); SB_TRY(val.alloc(N));
  SB_CUDA(cudaMemsetAsync(flag.p + N, 0, 4, s));
  const double norm = (double)(N - 1);
  SB_LAUNCH(k_result_flags, div_up(N, 256), 256, 0, s, g->inv.p, g->kahan_sum.p, N, g->row_begin, g->row_end, norm, flag.p, val.p,
            (uint32_t)g->world, (uint32_t)g->rank);
  SB_CHECK_LAUNCH();
  size_t need = 0;
  SB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, need, flag.p, pos.p, (int64_t)(N + 1), s));
  if (g->cub_tmp.n < need) SB_TRY(g->cub_tmp.alloc(need + 256));
  SB_CUDA(cub::DeviceScan::ExclusiveSum(g->cub_tmp.p, need, flag.p, pos.p, (int64_t)(N + 1), s));
  g_launches.fetch_add(2, std::memory_order_relaxed);
  uint32_t total = 0;
  SB_CUDA(cudaMemcpyAsync(&total, pos.p + N, 4, cudaMemcpyDeviceToHost, s));
  SB_CUDA(cudaStreamSynchronize(s));
  const bool timing = getenv("SB200_RESULT_TIMING") != nullptr;
  const auto tp1 = std::chrono::steady_clock::now();
  if (timing) fprintf(stderr, "[sb200 result] flags+scan %.2f ms\n", std::chrono::duration<double, std::milli>(tp1 - tp0).count());
  *len = total;
  if (!cent) return SB200_OK;
  const uint64_t k = std::min<uint64_t>(total, cap);
  if (k == 0) return SB200_OK;
  DevBuf<uint64_t> olo, ohi; DevBuf<double> oc;
  SB_TRY(olo.alloc(k)); SB_TRY(ohi.alloc(k)); SB_TRY(oc.alloc(k));
  SB_LAUNCH(k_result_scatter, div_up(N, 256), 256, 0, s, flag.p, pos.p, val.p, g->id_lo.p, g->id_hi.p, N, k, olo.p, ohi.p, oc.p);
  SB_CHECK_LAUNCH();
  SB_CUDA(cudaMemcpyAsync(id_lo, olo.p, k * 8, cudaMemcpyDefault, s));
  SB_CUDA(cudaMemcpyAsync(id_hi, ohi.p, k * 8, cudaMemcpyDefault, s));
  SB_CUDA(cudaMemcpyAsync(cent, oc.p, k * 8, cudaMemcpyDefault, s));
  SB_CUDA(cudaStreamSynchronize(s));
  if (timing) fprintf(stderr, "[sb200 result] scatter+d2h of %llu rows %.2f ms\n", (unsigned long long)k,
                      std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tp1).count());
  return SB200_OK;
}

// ---- rank assignment (store_harmonic's second pass, crates/core/src/webgraph/centrality/mod.rs:88-108) -----------
// key = ~order-preserving bits of the centrality (ascending key == descending f64 total order); the stable radix sort
// keeps the input order among equal keys, so feeding the nodes in ascending (descending) id order yields the
// (centrality desc, id asc) order of the rank store, or top_nodes' (centrality, id) descending order (mod.rs:17-37)
__global__ void k_rank_keys(const uint32_t* __restrict__ flag, const uint32_t* __restrict__ pos, const double* __restrict__ val,
                            uint64_t N, uint32_t total, int ties_desc, uint64_t* keys, uint32_t* ranks) {
  const uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (r >= N || !flag[r]) return;
  const uint32_t p = ties_desc ? total - 1u - pos[r] : pos[r];
  const uint64_t b = (uint64_t)__double_as_longlong(val[r]);
  keys[p] = ~((b >> 63) ? ~b : (b | 0x8000000000000000ull));
  ranks[p] = (uint32_t)r;
}
__global__ void k_rank_gather(const uint32_t* __restrict__ order, const double* __restrict__ val, const uint64_t* __restrict__ id_lo,
                              const uint64_t* __restrict__ id_hi, uint64_t k, uint64_t* out_lo, uint64_t* out_hi, double* out_c) {
  const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= k) return;
  const uint32_t r = order[i];
  out_lo[i] = id_lo[r]; out_hi[i] = id_hi[r]; out_c[i] = val[r];
}

int hb_ranked(sb200_graph* g, int ties_desc, uint64_t* id_lo, uint64_t* id_hi, double* cent, uint64_t cap, uint64_t* len) {
  cudaStream_t s = g->stream;
  const uint64_t N = g->N;
  if (g->world != 1) SB_FAIL(SB200_ESTATE, "ranking needs every node: gather the sharded results first");
  if (N == 0) { *len = 0; return SB200_OK; }
  PoolScope scope(getenv("SB200_NO_POOL") ? nullptr : s);
  DevBuf<uint32_t> flag, pos; DevBuf<double> val;
  SB_TRY(flag.alloc(N + 1)); SB_TRY(pos.alloc(N + 1)); SB_TRY(val.alloc(N));
  SB_CUDA(cudaMemsetAsync(flag.p + N, 0, 4, s));
  SB_LAUNCH(k_result_flags, div_up(N, 256), 256, 0, s, g->inv.p, g->kahan_sum.p, N, g->row_begin, g->row_end, (double)(N - 1), flag.p,
            val.p, 1u, 0u);
  SB_CHECK_LAUNCH();
  size_t need = 0;
  SB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, need, flag.p, pos.p, (int64_t)(N + 1), s));
  if (g->cub_tmp.n < need) SB_TRY(g->cub_tmp.alloc(need + 256));
  SB_CUDA(cub::DeviceScan::ExclusiveSum(g->cub_tmp.p, need, flag.p, pos.p, (int64_t)(N + 1), s));
  g_launches.fetch_add(2, std::memory_order_relaxed);
  uint32_t total = 0;
  SB_CUDA(cudaMemcpyAsync(&total, pos.p + N, 4, cudaMemcpyDeviceToHost, s));
  SB_CUDA(cudaStreamSynchronize(s));
  *len = total;
  const uint64_t k = std::min<uint64_t>(total, cap);
  if (!cent || k == 0) return SB200_OK;
  DevBuf<uint64_t> ka, kb; DevBuf<uint32_t> va, vb;
  SB_TRY(ka.alloc(total)); SB_TRY(kb.alloc(total)); SB_TRY(va.alloc(total)); SB_TRY(vb.alloc(total));
  SB_LAUNCH(k_rank_keys, div_up(N, 256), 256, 0, s, flag.p, pos.p, val.p, N, total, ties_desc, ka.p, va.p);
  SB_CHECK_LAUNCH();
  cub::DoubleBuffer<uint64_t> dk(ka.p, kb.p);
  cub::DoubleBuffer<uint32_t> dv(va.p, vb.p);
  need = 0;
  SB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, need, dk, dv, (int64_t)total, 0, 64, s));
  if (g->cub_tmp.n < need) SB_TRY(g->cub_tmp.alloc(need + 256));
  SB_CUDA(cub::DeviceRadixSort::SortPairs(g->cub_tmp.p, need, dk, dv, (int64_t)total, 0, 64, s));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  DevBuf<uint64_t> olo, ohi; DevBuf<double> oc;
  SB_TRY(olo.alloc(k)); SB_TRY(ohi.alloc(k)); SB_TRY(oc.alloc(k));
  SB_LAUNCH(k_rank_gather, div_up(k, 256), 256, 0, s, dv.Current(), val.p, g->id_lo.p, g->id_hi.p, k, olo.p, ohi.p, oc.p);
  SB_CHECK_LAUNCH();
  SB_CUDA(cudaMemcpyAsync(id_lo, olo.p, k * 8, cudaMemcpyDefault, s));
  SB_CUDA(cudaMemcpyAsync(id_hi, ohi.p, k * 8, cudaMemcpyDefault, s));
  SB_CUDA(cudaMemcpyAsync(cent, oc.p, k * 8, cudaMemcpyDefault, s));
  SB_CUDA(cudaStreamSynchronize(s));
  return SB200_OK;
}

int hb_registers(sb200_graph* g, uint64_t first, uint64_t count, uint8_t* out) {
  if (first + count > g->N) SB_FAIL(SB200_EINVAL, "register range [%llu,+%llu) outside %llu nodes", (unsigned long long)first, (unsigned long long)count, (unsigned long long)g->N);
  if (!count) return SB200_OK;
  DevBuf<uint4> tmp; SB_TRY(tmp.alloc(count * 4));
  SB_LAUNCH(k_gather_regs, div_up(count * 4, 256), 256, 0, g->stream, g->inv.p, (const uint4*)g->regs[g->cur].p, first, count, tmp.p);
  SB_CHECK_LAUNCH();
  SB_CUDA(cudaMemcpyAsync(out, tmp.p, count * 64, cudaMemcpyDefault, g->stream));
  SB_CUDA(cudaStreamSynchronize(g->stream));
  return SB200_OK;
}
int hb_kahan(sb200_graph* g, uint64_t first, uint64_t count, double* sum, double* err) {
  if (first + count > g->N) SB_FAIL(SB200_EINVAL, "range outside the node set");
  if (!count) return SB200_OK;
  DevBuf<double> a, b; SB_TRY(a.alloc(count)); SB_TRY(b.alloc(count));
  SB_LAUNCH(k_gather_f64x2, div_up(count, 256), 256, 0, g->stream, g->inv.p, g->kahan_sum.p, g->kahan_err.p, first, count, a.p, b.p);
  SB_CHECK_LAUNCH();
  SB_CUDA(cudaMemcpyAsync(sum, a.p, count * 8, cudaMemcpyDefault, g->stream));
  SB_CUDA(cudaMemcpyAsync(err, b.p, count * 8, cudaMemcpyDefault, g->stream));
  SB_CUDA(cudaStreamSynchronize(g->stream));
  return SB200_OK;
}

}  // namespace sb200
