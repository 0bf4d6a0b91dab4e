// graph_sim.cu -- inbound similarity on the resident CSR (third item).
//
// Reference:
//   bitvec_similarity::BitVec::{new, sim}           crates/core/src/ranking/bitvec_similarity.rs:130-185
//     the sorted, de-duplicated ids of a node's in-neighbours + a 16 x 64-bit "VeryJankyBloomFilter" over their LOW 64 bits
//     (hash = id * 11400714819323198549; word = hash % 16, bit = hash % 64) + sqrt(len);
//     sim(a, b) = 0 if either is empty, 0 if popcount(bloom_a & bloom_b) / max(ones_a, ones_b) < 0.25 (the pre-filter, with its
//     false negatives, is part of the result), else |a ∩ b| / (sqrt_len_a * sqrt_len_b)
//   inbound_similarity::Scorer::{new, calculate_score}   crates/core/src/ranking/inbound_similarity.rs:71-119
//     s = |disliked| + (sum_liked sim - sum_disliked sim), / max(|liked|, 1) if normalized, max(0); a liked / disliked node
//     compared with itself scores `self_score`.
//
// A node's in-neighbour set is its CSR row (unique sources of the kept links) plus itself if it links to itself (the
// staging drops self-links from the CSR -- they are no-ops for HyperBall -- but remembers them in `self_bm`).
// Device formulation: the <= 64 liked / disliked nodes of a pass mark their in-neighbours in one u64 word per node
// (bit j = "in-neighbour of target j"); a candidate's |row ∩ row_j| for all j at once is then a walk over ITS row
// counting bit j of the marks -- no sorted merge, no second row read.
#include "graph.cuh"

#include <algorithm>
#include <vector>

namespace sb200 {
// u128 id -> rank by binary search over the ascending (hi, lo) arrays; 0xFFFFFFFF when the id is not a node
__global__ void k_sim_id_to_rank(const uint64_t* __restrict__ q_lo, const uint64_t* __restrict__ q_hi, uint32_t n, const uint64_t* __restrict__ id_lo,
                                 const uint64_t* __restrict__ id_hi, uint64_t N, uint32_t* out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t lo = q_lo[i], hi = q_hi[i];
  uint64_t a = 0, b = N;
  while (a < b) {
    const uint64_t m = (a + b) >> 1;
    if (id_hi[m] < hi || (id_hi[m] == hi && id_lo[m] < lo)) a = m + 1; else b = m;
  }
  out[i] = (a < N && id_hi[a] == hi && id_lo[a] == lo) ? (uint32_t)a : 0xFFFFFFFFu;
}

struct SimNode { unsigned long long bloom[16]; uint32_t ones, len, internal, rank; };   // rank 0xFFFFFFFF: not a node of the graph

// one warp per listed node: bloom filter, length, internal row; targets also mark their in-neighbours with bit `bit0 + i`
__global__ void __launch_bounds__(128) k_sim_prepare(const uint32_t* __restrict__ ranks, uint32_t n, const uint32_t* __restrict__ inv,
    const uint32_t* __restrict__ perm, const uint64_t* __restrict__ id_lo, const uint32_t* __restrict__ row_ptr, const uint32_t* __restrict__ col,
    const uint32_t* __restrict__ self_bm, SimNode* out, unsigned long long* marks, int mark_bits) {
  __shared__ unsigned long long s_bloom[4][16];
  const uint32_t w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t i = blockIdx.x * 4 + w;
  if (i >= n) return;
  if (lane < 16) s_bloom[w][lane] = 0;
  __syncwarp();
  const uint32_t r = ranks[i];
  uint32_t len = 0, v = 0;
  if (r != 0xFFFFFFFFu) {
    v = inv[r];
    const uint32_t e0 = row_ptr[v], e1 = row_ptr[v + 1];
    const bool self = (self_bm[r >> 5] >> (r & 31u)) & 1u;
    len = e1 - e0 + (self ? 1u : 0u);
    const unsigned long long bit = mark_bits ? (1ull << i) : 0ull;
    for (uint32_t e = e0 + lane; e < e1 + (self ? 1u : 0u); e += 32) {
      const uint32_t u = e < e1 ? col[e] : v;
      const unsigned long long h = id_lo[perm[u]] * 11400714819323198549ull;   // insert_u128 keeps the low 64 bits
      atomicOr(&s_bloom[w][h % 16ull], 1ull << (h % 64ull));
      if (mark_bits) atomicOr(marks + u, bit);
    }
  }
  __syncwarp();
  uint32_t ones = lane < 16 ? (uint32_t)__popcll(s_bloom[w][lane]) : 0u;
  for (int o = 16; o; o >>= 1) ones += __shfl_xor_sync(0xffffffffu, ones, o);
  if (lane < 16) out[i].bloom[lane] = s_bloom[w][lane];
  if (lane == 0) { out[i].ones = ones; out[i].len = len; out[i].internal = v; out[i].rank = r; }
}

// one warp per candidate: sims[c][j0 + j] for the targets of this pass
__global__ void __launch_bounds__(128) k_sim_score(const SimNode* __restrict__ cand, uint32_t n_cand, const SimNode* __restrict__ targ, uint32_t n_targ,
    const uint64_t* __restrict__ c_lo, const uint64_t* __restrict__ c_hi, const uint64_t* __restrict__ t_lo, const uint64_t* __restrict__ t_hi,
    const uint32_t* __restrict__ row_ptr, const uint32_t* __restrict__ col, const uint32_t* __restrict__ self_bm,
    const unsigned long long* __restrict__ marks, double self_score, double* sims, uint32_t stride, uint32_t j0) {
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (c >= n_cand) return;
  const SimNode& C = cand[c];
  // |in(c) ∩ in(t_j)| for every j: count bit j over the marks of c's in-neighbours (lane j keeps the count of target j and j + 32)
  uint32_t cnt_lo = 0, cnt_hi = 0;
  if (C.rank != 0xFFFFFFFFu && C.len) {
    const uint32_t v = C.internal, e0 = row_ptr[v], e1 = row_ptr[v + 1];
    const bool self = (self_bm[C.rank >> 5] >> (C.rank & 31u)) & 1u;
    const uint32_t total = e1 - e0 + (self ? 1u : 0u);
    for (uint32_t base = 0; base < total; base += 32) {
      const uint32_t k = base + lane;
      unsigned long long m = 0;
      if (k < total) m = marks[k < e1 - e0 ? col[e0 + k] : v];
      for (uint32_t j = 0; j < n_targ; j++) {
        const uint32_t b = __popc(__ballot_sync(0xffffffffu, (m >> j) & 1ull));
        if ((j & 31u) == lane) { if (j < 32) cnt_lo += b; else cnt_hi += b; }
      }
    }
  }
  for (uint32_t j = lane; j < n_targ; j += 32) {
    const SimNode& T = targ[j];
    double sim;
    if (c_lo[c] == t_lo[j] && c_hi[c] == t_hi[j]) sim = self_score;          // NodeScorer::sim: the node itself
    else if (C.len == 0 || T.len == 0) sim = 0.0;
    else {
      uint32_t inter = 0;
      for (int w = 0; w < 16; w++) inter += (uint32_t)__popcll(C.bloom[w] & T.bloom[w]);
      const uint32_t mx = max(C.ones, T.ones);
      if (__ddiv_rn((double)inter, (double)mx) < 0.25) sim = 0.0;
      else {
        const double isz = (double)(j < 32 ? cnt_lo : cnt_hi);
        sim = __ddiv_rn(isz, __dmul_rn(__dsqrt_rn((double)T.len), __dsqrt_rn((double)C.len)));   // self = the liked node's BitVec
      }
    }
    sims[(size_t)c * stride + j0 + j] = sim;
  }
}

// calculate_score: sequential f64 sums in list order
__global__ void k_sim_total(const double* __restrict__ sims, uint32_t n_cand, uint32_t n_liked, uint32_t n_disliked, int normalized, double* out) {
  const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_cand) return;
  const double* s = sims + (size_t)c * (n_liked + n_disliked);
  double l = 0.0, d = 0.0;
  for (uint32_t j = 0; j < n_liked; j++) l = __dadd_rn(l, s[j]);
  for (uint32_t j = 0; j < n_disliked; j++) d = __dadd_rn(d, s[n_liked + j]);
  double v = __dadd_rn((double)n_disliked, __dsub_rn(l, d));
  if (normalized) v = __ddiv_rn(v, (double)max(n_liked, 1u));
  out[c] = fmax(v, 0.0);
}
}  // namespace sb200
using namespace sb200;

extern "C" int sb200_inbound_similarity(sb200_graph* g, const uint64_t* liked_lo, const uint64_t* liked_hi, uint32_t n_liked,
                                        const uint64_t* disliked_lo, const uint64_t* disliked_hi, uint32_t n_disliked,
                                        const uint64_t* cand_lo, const uint64_t* cand_hi, uint32_t n_cand, int normalized,
                                        double self_score, double* scores) {
  if (!g) SB_FAIL(SB200_EINVAL, "NULL graph handle");
  SB_CUDA(cudaSetDevice(g->device));
  if (g->world != 1) SB_FAIL(SB200_ESTATE, "inbound similarity runs on single-rank handles");
  if ((n_liked && (!liked_lo || !liked_hi)) || (n_disliked && (!disliked_lo || !disliked_hi)) || (n_cand && (!cand_lo || !cand_hi || !scores)))
    SB_FAIL(SB200_EINVAL, "NULL argument");
  if (!n_cand) return SB200_OK;
  cudaStream_t s = g->stream;
  const uint64_t N = g->N;
  const uint32_t NT = n_liked + n_disliked;
  PoolScope scope(s);
  // ids -> device, ranks
  std::vector<uint64_t> tl(NT), th(NT);
  for (uint32_t i = 0; i < n_liked; i++) { tl[i] = liked_lo[i]; th[i] = liked_hi[i]; }
  for (uint32_t i = 0; i < n_disliked; i++) { tl[n_liked + i] = disliked_lo[i]; th[n_liked + i] = disliked_hi[i]; }
  DevBuf<uint64_t> d_tl, d_th, d_cl, d_ch; DevBuf<uint32_t> r_t, r_c;
  SB_TRY(d_tl.alloc(std::max<uint32_t>(NT, 1))); SB_TRY(d_th.alloc(std::max<uint32_t>(NT, 1))); SB_TRY(r_t.alloc(std::max<uint32_t>(NT, 1)));
  SB_TRY(d_cl.alloc(n_cand)); SB_TRY(d_ch.alloc(n_cand)); SB_TRY(r_c.alloc(n_cand));
  if (NT) { SB_CUDA(cudaMemcpyAsync(d_tl.p, tl.data(), NT * 8, cudaMemcpyHostToDevice, s)); SB_CUDA(cudaMemcpyAsync(d_th.p, th.data(), NT * 8, cudaMemcpyHostToDevice, s)); }
  SB_CUDA(cudaMemcpyAsync(d_cl.p, cand_lo, (size_t)n_cand * 8, cudaMemcpyDefault, s));
  SB_CUDA(cudaMemcpyAsync(d_ch.p, cand_hi, (size_t)n_cand * 8, cudaMemcpyDefault, s));
  if (N) {
    if (NT) { SB_LAUNCH(k_sim_id_to_rank, div_up(NT, 256), 256, 0, s, d_tl.p, d_th.p, NT, g->id_lo.p, g->id_hi.p, N, r_t.p); SB_CHECK_LAUNCH(); }
    SB_LAUNCH(k_sim_id_to_rank, div_up(n_cand, 256), 256, 0, s, d_cl.p, d_ch.p, n_cand, g->id_lo.p, g->id_hi.p, N, r_c.p); SB_CHECK_LAUNCH();
  } else {
    if (NT) SB_CUDA(cudaMemsetAsync(r_t.p, 0xFF, NT * 4, s));
    SB_CUDA(cudaMemsetAsync(r_c.p, 0xFF, (size_t)n_cand * 4, s));
  }
  DevBuf<SimNode> cn, tn; DevBuf<unsigned long long> marks; DevBuf<double> sims, out;
  SB_TRY(cn.alloc(n_cand)); SB_TRY(tn.alloc(std::max<uint32_t>(std::min<uint32_t>(NT, 64), 1)));
  SB_TRY(marks.alloc(std::max<uint64_t>(N, 1))); SB_TRY(sims.alloc((size_t)n_cand * std::max<uint32_t>(NT, 1))); SB_TRY(out.alloc(n_cand));
  SB_LAUNCH(k_sim_prepare, div_up(n_cand, 4), 128, 0, s, r_c.p, n_cand, g->inv.p, g->perm.p, g->id_lo.p, g->row_ptr.p, g->col.p, g->self_bm.p, cn.p,
            (unsigned long long*)nullptr, 0);
  SB_CHECK_LAUNCH();
  for (uint32_t j0 = 0; j0 < NT; j0 += 64) {   // <= 64 liked / disliked nodes per pass: one mark bit each
    const uint32_t nj = std::min<uint32_t>(64, NT - j0);
    SB_CUDA(cudaMemsetAsync(marks.p, 0, std::max<uint64_t>(N, 1) * 8, s));
    SB_LAUNCH(k_sim_prepare, div_up(nj, 4), 128, 0, s, r_t.p + j0, nj, g->inv.p, g->perm.p, g->id_lo.p, g->row_ptr.p, g->col.p, g->self_bm.p, tn.p, marks.p, 1);
    SB_CHECK_LAUNCH();
    SB_LAUNCH(k_sim_score, div_up((uint64_t)n_cand * 32, 128), 128, 0, s, cn.p, n_cand, tn.p, nj, d_cl.p, d_ch.p, d_tl.p + j0, d_th.p + j0, g->row_ptr.p,
              g->col.p, g->self_bm.p, marks.p, self_score, sims.p, NT, j0);
    SB_CHECK_LAUNCH();
  }
  SB_LAUNCH(k_sim_total, div_up(n_cand, 256), 256, 0, s, sims.p, n_cand, n_liked, n_disliked, normalized, out.p);
  SB_CHECK_LAUNCH();
  SB_CUDA(cudaMemcpyAsync(scores, out.p, (size_t)n_cand * 8, cudaMemcpyDefault, s));
  SB_CUDA(cudaStreamSynchronize(s));
  return SB200_OK;
}
