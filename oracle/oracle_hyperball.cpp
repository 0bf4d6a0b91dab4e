// oracle/oracle_hyperball.cpp -- CPU restatement of the reference's harmonic-centrality path.
// TEST INFRASTRUCTURE ONLY (see oracle_common.h).  Follows, by file:line in /root/reference:
//   crates/core/src/hyperloglog.rs:4302-4547      HyperLogLog<N, FastHasher>  (add/size/merge)
//   crates/core/src/kahan_sum.rs:30-54            KahanSum
//   crates/bloom/src/lib.rs:38-130                U64BloomFilter
//   crates/core/src/webgraph/centrality/harmonic.rs:34-287   HyperBall loop
//   crates/core/src/webgraph/store.rs:297-357     host_edges()/host_nodes() (dedup, first wins)
//
// Parity status.  Pinned by reference KATs: KahanSum (kahan_sum.rs:86-125), bloom
// (bloom/src/lib.rs:194-245), HLL<128> property tests (hyperloglog.rs:4553-4599), harmonic
// orderings (harmonic.rs:358-578).  "Parity unpinned" for the numeric value of
// HyperLogLog<64>::size(): the reference holds no numeric golden for N=64, and size() runs
// std's binary search over an empirical table that is NOT monotone (precision-5 raw table has
// inversions at indices 127/128 and 130/131), so the result depends on the std implementation;
// we follow Rust >= 1.82 (`slice::binary_search_by`, branch-free variant).
//
// Two implementations of the same math:
//   orc_hb_faithful_*  mirrors the reference's data structures (ordered maps keyed by u128,
//                      one heap vector per counter, deep clone per iteration, bloom frontier,
//                      per-scan hash-set dedup, 2N size() calls per iteration) -- this is the
//                      "reference CPU path" that gets timed, single-threaded like the reference.
//   orc_hb_dense_*     flat arrays over dense node ranks, steppable, optional OpenMP --
//                      the parity checker for per-iteration registers and the "optimised CPU"
//                      baseline.
#include "oracle_common.h"
#include "../include/sb200_hll_tables.h"
#include "oracle_dense.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <set>
#include <unordered_map>
#include <unordered_set>
#include <vector>
#include <atomic>
#include <thread>

// minimal parallel-for (this image's g++ wrapper cannot find libgomp.spec, so no OpenMP)
template <class F> static void orc_parallel_for(int64_t n, int threads, int64_t grain, F f) {
  if (threads <= 1 || n <= grain) { f(0, n, 0); return; }
  std::atomic<int64_t> next(0);
  std::vector<std::thread> pool;
  for (int t = 0; t < threads; t++)
    pool.emplace_back([&, t]() {
      for (;;) { int64_t b = next.fetch_add(grain); if (b >= n) break; f(b, std::min(n, b + grain), t); }
    });
  for (auto& th : pool) th.join();
}

// ---------------------------------------------------------------- HyperLogLog ---------------
static const uint64_t LARGE_PRIME = 11400714819323198549ull;  // hyperloglog.rs:4312, bloom lib.rs:36

static inline int hll_b(int n) {  // hyperloglog.rs:4381-4383  (N as f64).log2() as usize
  return (int)std::log2((double)n);
}

ORC_API void orc_hll_add(uint8_t* regs, int n, uint64_t item) {  // hyperloglog.rs:4385-4396
  const int b = hll_b(n);
  const uint64_t hash = item * LARGE_PRIME;  // FastHasher, wrapping
  const uint64_t j = hash >> (64 - b);
  const uint64_t w = hash << b;
  const unsigned p = (w == 0 ? 64u : (unsigned)__builtin_clzll(w)) + 1u;
  if ((uint8_t)p > regs[j]) regs[j] = (uint8_t)p;
}

ORC_API void orc_hll_add_range(uint8_t* regs, int n, uint64_t first, uint64_t count) {
  for (uint64_t i = 0; i < count; i++) orc_hll_add(regs, n, first + i);
}

ORC_API void orc_hll_merge(uint8_t* dst, const uint8_t* src, int n) {  // :4531-4535
  for (int i = 0; i < n; i++) dst[i] = std::max(dst[i], src[i]);
}

// Rust >= 1.82 core::slice::binary_search_by on `v.total_cmp(&e)`, returning the index the
// reference then uses (Ok(i)|Err(i) -> i, Err(len) -> len-1).  hyperloglog.rs:4411-4416.
static inline size_t rust_bsearch_total(const double* a, size_t len, double e) {
  size_t size = len, base = 0;
  while (size > 1) {
    size_t half = size / 2, mid = base + half;
    // cmp = a[mid].total_cmp(e); keep base when Greater
    base = (a[mid] > e) ? base : mid;
    size -= half;
  }
  size_t r = base;
  if (!(a[base] == e)) r = base + (a[base] < e ? 1 : 0);
  if (r == len) r = len - 1;
  return r;
}

static double hll_estimate_bias(double e, const double* raw, const double* bias, size_t len) {
  // hyperloglog.rs:4408-4469
  const int K = 6;
  long idx_left = (long)rust_bsearch_total(raw, len, e);
  long idx_right = (idx_left < (long)len - 1) ? idx_left + 1 : -1;
  long neighbors[K];
  for (int k = 0; k < K; k++) {
    bool right_instead_left;
    long idx;
    if (idx_left >= 0 && idx_right >= 0) {
      double dl = std::fabs(raw[idx_left] - e), dr = std::fabs(raw[idx_right] - e);
      if (dr < dl) { right_instead_left = true; idx = idx_right; }
      else { right_instead_left = false; idx = idx_left; }
    } else if (idx_left >= 0) { right_instead_left = false; idx = idx_left; }
    else { right_instead_left = true; idx = idx_right; }
    neighbors[k] = idx;
    if (right_instead_left) idx_right = (idx < (long)len - 1) ? idx + 1 : -1;
    else idx_left = (idx > 0) ? idx - 1 : -1;
  }
  double s = 0.0;
  for (int k = 0; k < K; k++) s += bias[neighbors[k]];
  return s / (double)K;
}

struct Pow2Tab { double v[256]; Pow2Tab() { for (int k = 0; k < 256; k++) v[k] = std::ldexp(1.0, -k); } };
static const Pow2Tab POW2NEG;  // ONE_OVER_POWER_OF_TWO hyperloglog.rs:4043 (entries are exactly 2^-k)

ORC_API uint64_t orc_hll_size(const uint8_t* regs, int n) {  // hyperloglog.rs:4484-4516
  const double m = (double)n;
  const int b = hll_b(n);
  double am;  // :4366-4378
  if (n >= 128) am = 0.7213 / (1. + (1.079 / m));
  else if (n >= 64) am = 0.709;
  else if (n >= 32) am = 0.697;
  else am = 0.673;
  double sum = 0.0;
  size_t v = 0;
  for (int i = 0; i < n; i++) {
    sum += POW2NEG.v[regs[i]];
    v += (regs[i] == 0);
  }
  const double z = 1.0 / sum;
  const double e = am * (m * m) * z;
  double e_star = e;
  if (e <= 5. * m) {
    // table index b-1-4: N=64 -> "precision 5", N=128 -> "precision 6"
    const double *raw, *bias; size_t len;
    if (b == 6) { raw = SB200_HLL_RAW_P5; bias = SB200_HLL_BIAS_P5; len = SB200_HLL_P5_LEN; }
    else if (b == 7) { raw = SB200_HLL_RAW_P6; bias = SB200_HLL_BIAS_P6; len = SB200_HLL_P6_LEN; }
    else return UINT64_MAX;  // other sizes are not on the hot path
    e_star = e - hll_estimate_bias(e, raw, bias, len);
  }
  double h = (v != 0) ? m * std::log(m / (double)v) : e_star;  // linear_counting :4472-4476
  static const double THRESH[] = {10, 20, 40, 80, 220, 400, 900, 1800, 3100, 6500,
                                  11500, 20000, 50000, 120000, 350000};  // :28-44
  const double thr = THRESH[b - 4];
  return (h <= thr) ? orc_f64_as_u64(h) : orc_f64_as_u64(e_star);
}

// the 65-entry linear-counting LUT the GPU library receives from its host side is
// m*ln(m/v); exposed so tests can pin the library's LUT against libm here.
ORC_API double orc_hll64_linear_counting(uint32_t v) { return 64.0 * std::log(64.0 / (double)v); }

// ---------------------------------------------------------------- KahanSum ------------------
ORC_API void orc_kahan_add(double* sum, double* err, double rhs) {  // kahan_sum.rs:46-53
  volatile double y = rhs - *err;
  volatile double t = *sum + y;
  volatile double e = (t - *sum) - y;
  *err = e;
  *sum = t;
}

// ---------------------------------------------------------------- U64BloomFilter ------------
struct Bloom {
  std::vector<uint64_t> w;
  uint64_t nbits = 0;
  static uint64_t num_bits(uint64_t items, double fp) {  // bloom lib.rs:40-42
    const double ln2 = std::log(2.0);
    return (uint64_t)std::ceil(((double)items) * std::log(fp) / (-8.0 * (ln2 * ln2)));
  }
  Bloom() {}
  Bloom(uint64_t items, double fp) { nbits = num_bits(items, fp); w.assign((nbits + 63) / 64, 0); }
  void insert(uint64_t item) { uint64_t h = (item * LARGE_PRIME) % nbits; w[h >> 6] |= 1ull << (h & 63); }
  bool contains(uint64_t item) const { uint64_t h = (item * LARGE_PRIME) % nbits; return (w[h >> 6] >> (h & 63)) & 1; }
  uint64_t count_ones() const { uint64_t c = 0; for (uint64_t x : w) c += __builtin_popcountll(x); return c; }
  uint64_t estimate_card() const {  // lib.rs:108-123 ; note `.ln() as i64` binds first
    uint64_t ones = count_ones();
    if (ones == 0 || nbits == 0) return 0;
    if (ones == nbits) return UINT64_MAX;
    double l = std::log(1.0 - ((double)ones) / ((double)nbits));
    int64_t li = (int64_t)l;  // truncation toward zero (|l| small here)
    int64_t r = (-(int64_t)nbits) * li;
    return r < 0 ? 0 : (uint64_t)r;
  }
};

ORC_API uint64_t orc_bloom_num_bits(uint64_t items, double fp) { return Bloom::num_bits(items, fp); }
ORC_API void* orc_bloom_new(uint64_t items, double fp) { return new Bloom(items, fp); }
ORC_API void orc_bloom_free(void* b) { delete (Bloom*)b; }
ORC_API void orc_bloom_insert(void* b, uint64_t x) { ((Bloom*)b)->insert(x); }
ORC_API int orc_bloom_contains(void* b, uint64_t x) { return ((Bloom*)b)->contains(x); }
ORC_API uint64_t orc_bloom_estimate_card(void* b) { return ((Bloom*)b)->estimate_card(); }

// ---------------------------------------------------------------- graph staging -------------
struct U128Hash { size_t operator()(u128 x) const { uint64_t a = (uint64_t)x, b = (uint64_t)(x >> 64); a ^= b * 0x9E3779B97F4A7C15ull; a ^= a >> 29; a *= 0xBF58476D1CE4E5B9ull; a ^= a >> 32; return (size_t)a; } };
struct PairHash { size_t operator()(const std::pair<u128, u128>& p) const { U128Hash h; return h(p.first) * 31 + h(p.second) + 0x1234567; } };

struct EdgeIn { u128 from, to; uint64_t rel; };

static std::vector<EdgeIn> gather_edges(const uint64_t* flo, const uint64_t* fhi, const uint64_t* tlo,
                                        const uint64_t* thi, const uint64_t* rel, uint64_t n) {
  std::vector<EdgeIn> v(n);
  for (uint64_t i = 0; i < n; i++) v[i] = {orc_make_u128(fhi[i], flo[i]), orc_make_u128(thi[i], tlo[i]), rel[i]};
  return v;
}

// ---------------------------------------------------------------- faithful HyperBall --------
typedef std::vector<uint8_t> Hll;

struct Faithful {
  std::vector<EdgeIn> edges;  // raw stream as the Webgraph iterator would yield it
  uint64_t skip_mask;
  std::map<u128, double> result;
  uint64_t num_nodes = 0;
  uint32_t iters = 0;
  // Webgraph::host_edges(): every scan re-deduplicates on (from,to), first wins (store.rs:313)
  template <class F> void for_host_edges(F f) const {
    std::unordered_set<std::pair<u128, u128>, PairHash> seen;
    seen.reserve(edges.size() * 2);
    for (const EdgeIn& e : edges) if (seen.insert({e.from, e.to}).second) f(e);
  }
};

static inline bool any_greater(const Hll& from, const Hll& to) {
  for (size_t i = 0; i < to.size(); i++) if (from[i] > to[i]) return true;
  return false;
}

ORC_API void* orc_hb_faithful_run(const uint64_t* flo, const uint64_t* fhi, const uint64_t* tlo,
                                  const uint64_t* thi, const uint64_t* rel, uint64_t n_edges,
                                  uint64_t skip_mask, uint32_t max_iters) {
  Faithful* g = new Faithful();
  g->edges = gather_edges(flo, fhi, tlo, thi, rel, n_edges);
  g->skip_mask = skip_mask;
  const int NREG = 64;  // HYPERLOGLOG_COUNTERS harmonic.rs:34

  // host_nodes(): all endpoints of all edges (incl. skipped), store.rs:338-357
  std::unordered_set<u128, U128Hash> nodes;
  for (const EdgeIn& e : g->edges) { nodes.insert(e.from); nodes.insert(e.to); }

  // forward links for the small-frontier branch (ForwardlinksQuery, harmonic.rs:86)
  std::unordered_map<u128, std::vector<std::pair<u128, uint64_t>>, U128Hash> fwd;
  g->for_host_edges([&](const EdgeIn& e) { fwd[e.from].push_back({e.to, e.rel}); });

  std::map<u128, Hll> old_c, new_c;           // Counters harmonic.rs:197-213
  std::map<u128, Kahan> cent;
  for (u128 node : nodes) {                     // initialize :53-73
    Hll c(NREG, 0);
    orc_hll_add(c.data(), NREG, (uint64_t)node);  // add_u128 uses the low 64 bits :4398-4400
    old_c.emplace(node, std::move(c));
    cent.emplace(node, Kahan());
  }
  new_c = old_c;
  const uint64_t num_nodes = nodes.size();
  g->num_nodes = num_nodes;
  if (num_nodes == 0) return g;

  Bloom changed(num_nodes, 0.05);               // :221-225
  for (u128 node : nodes) changed.insert((uint64_t)node);
  const uint64_t exact_thr = (uint64_t)std::max(0.0, std::round(std::sqrt((double)num_nodes)));
  const double norm = (double)(num_nodes - 1);
  bool exact_counting = false, has_changes = true;
  uint64_t t = 0;
  std::set<u128> exact_changed;

  while (has_changes && (max_iters == 0 || t < max_iters)) {
    Bloom new_changed(num_nodes, 0.05);
    if (!exact_changed.empty() && exact_changed.size() <= exact_thr) {
      // update_changed_counters :75-114
      std::set<u128> next;
      bool hc = false;
      for (u128 ch : exact_changed) {
        auto it = fwd.find(ch);
        if (it == fwd.end()) continue;
        for (auto& te : it->second) {
          if (te.second & skip_mask) continue;
          auto ct = new_c.find(te.first); auto cf = old_c.find(ch);
          if (ct == new_c.end() || cf == old_c.end()) continue;
          if (any_greater(cf->second, ct->second)) {
            orc_hll_merge(ct->second.data(), cf->second.data(), NREG);
            new_changed.insert((uint64_t)te.first);
            next.insert(te.first);
            hc = true;
          }
        }
      }
      exact_changed.swap(next);
      has_changes = hc;
    } else {
      // update_all_counters :116-157
      const bool track = exact_counting;
      if (track) exact_changed.clear();
      bool hc = false;
      g->for_host_edges([&](const EdgeIn& e) {
        if (e.rel & skip_mask) return;
        if (!changed.contains((uint64_t)e.from)) return;
        auto ct = new_c.find(e.to); auto cf = old_c.find(e.from);
        if (ct == new_c.end() || cf == old_c.end()) return;
        if (any_greater(cf->second, ct->second)) {
          orc_hll_merge(ct->second.data(), cf->second.data(), NREG);
          new_changed.insert((uint64_t)e.to);
          if (track) exact_changed.insert(e.to);
          hc = true;
        }
      });
      has_changes = hc;
    }
    // update_centralities :159-176 -- every node, two size() calls
    for (auto& kv : cent) {
      uint64_t sn = orc_hll_size(new_c[kv.first].data(), NREG);
      uint64_t so = orc_hll_size(old_c[kv.first].data(), NREG);
      uint64_t d = sn >= so ? sn - so : 0;  // checked_sub().unwrap_or_default()
      orc_kahan_add(&kv.second.sum, &kv.second.err, (double)d / (double)(t + 1));
    }
    old_c = new_c;  // Counters::step deep clone :210-212
    changed = std::move(new_changed);
    t += 1;
    if (changed.estimate_card() <= exact_thr) exact_counting = true;
  }
  g->iters = (uint32_t)t;
  // normalize_centralities :178-195
  for (auto& kv : cent) {
    double c = kv.second.sum;
    if (!(c > 0.0)) continue;
    c = c / norm;
    if (!std::isfinite(c)) c = 0.0;
    g->result.emplace(kv.first, c);
  }
  return g;
}

ORC_API uint64_t orc_hb_faithful_num_nodes(void* h) { return ((Faithful*)h)->num_nodes; }
ORC_API uint32_t orc_hb_faithful_iters(void* h) { return ((Faithful*)h)->iters; }
ORC_API uint64_t orc_hb_faithful_len(void* h) { return ((Faithful*)h)->result.size(); }
ORC_API void orc_hb_faithful_result(void* h, uint64_t* id_lo, uint64_t* id_hi, double* c) {
  uint64_t i = 0;
  for (auto& kv : ((Faithful*)h)->result) { id_lo[i] = (uint64_t)kv.first; id_hi[i] = (uint64_t)(kv.first >> 64); c[i] = kv.second; i++; }
}
ORC_API void orc_hb_faithful_free(void* h) { delete (Faithful*)h; }

// ---------------------------------------------------------------- dense HyperBall -----------
// Same math over dense node ranks (rank = position in ascending u128 order), synchronous
// update new[v] = max(old[v], max_{u->v kept, u changed} old[u]); steppable.

ORC_API void* orc_hb_dense_create(const uint64_t* flo, const uint64_t* fhi, const uint64_t* tlo,
                                  const uint64_t* thi, const uint64_t* rel, uint64_t n_edges,
                                  uint64_t skip_mask, int threads) {
  Dense* g = new Dense();
  g->threads = threads < 1 ? 1 : threads;
  std::vector<EdgeIn> edges = gather_edges(flo, fhi, tlo, thi, rel, n_edges);
  g->ids.reserve(edges.size() * 2);
  for (auto& e : edges) { g->ids.push_back(e.from); g->ids.push_back(e.to); }
  std::sort(g->ids.begin(), g->ids.end());
  g->ids.erase(std::unique(g->ids.begin(), g->ids.end()), g->ids.end());
  const uint64_t N = g->ids.size();
  auto rank = [&](u128 x) { return (uint32_t)(std::lower_bound(g->ids.begin(), g->ids.end(), x) - g->ids.begin()); };
  // dedup (from,to), first occurrence decides rel flags (store.rs:313)
  struct E { uint32_t to, from; uint64_t pos; uint8_t skip; };
  std::vector<E> es(edges.size());
  for (uint64_t i = 0; i < edges.size(); i++)
    es[i] = {rank(edges[i].to), rank(edges[i].from), i, (uint8_t)((edges[i].rel & skip_mask) != 0)};
  std::sort(es.begin(), es.end(), [](const E& a, const E& b) {
    if (a.to != b.to) return a.to < b.to;
    if (a.from != b.from) return a.from < b.from;
    return a.pos < b.pos; });
  g->row_ptr.assign(N + 1, 0);
  for (uint64_t i = 0; i < es.size(); i++) {
    if (i > 0 && es[i].to == es[i - 1].to && es[i].from == es[i - 1].from) continue;
    if (es[i].skip) continue;
    g->col.push_back(es[i].from);
    g->row_ptr[es[i].to + 1]++;
  }
  for (uint64_t v = 0; v < N; v++) g->row_ptr[v + 1] += g->row_ptr[v];
  g->old_r.assign(N * 64, 0);
  for (uint64_t v = 0; v < N; v++) orc_hll_add(&g->old_r[v * 64], 64, (uint64_t)g->ids[v]);
  g->new_r = g->old_r;
  g->changed.assign(N, 1);
  g->new_changed.assign(N, 0);
  g->cent.assign(N, Kahan());
  g->size_old.resize(N);
  for (uint64_t v = 0; v < N; v++) g->size_old[v] = orc_hll_size(&g->old_r[v * 64], 64);
  return g;
}

ORC_API uint64_t orc_hb_dense_num_nodes(void* h) { return ((Dense*)h)->ids.size(); }
ORC_API uint64_t orc_hb_dense_num_edges(void* h) { return ((Dense*)h)->col.size(); }
ORC_API uint64_t orc_hb_dense_iters(void* h) { return ((Dense*)h)->t; }

// one synchronous iteration; returns the number of nodes whose registers changed
ORC_API uint64_t orc_hb_dense_step(void* h) {
  Dense* g = (Dense*)h;
  const int64_t N = (int64_t)g->ids.size();
  std::vector<uint64_t> n_changed_t(g->threads, 0);
  const double div = (double)(g->t + 1);
  orc_parallel_for(N, g->threads, 4096, [&](int64_t vb, int64_t ve, int tid) {
  uint64_t n_changed = 0;
  for (int64_t v = vb; v < ve; v++) {
    uint8_t acc[64];
    const uint8_t* ov = &g->old_r[v * 64];
    memcpy(acc, ov, 64);
    for (uint64_t e = g->row_ptr[v]; e < g->row_ptr[v + 1]; e++) {
      uint32_t u = g->col[e];
      if (!g->changed[u]) continue;
      const uint8_t* ou = &g->old_r[(uint64_t)u * 64];
      for (int i = 0; i < 64; i++) acc[i] = acc[i] > ou[i] ? acc[i] : ou[i];
    }
    bool ch = memcmp(acc, ov, 64) != 0;
    memcpy(&g->new_r[v * 64], acc, 64);
    g->new_changed[v] = ch;
    // update_centralities: every node, every iteration (zero adds included)
    uint64_t sn = ch ? orc_hll_size(acc, 64) : g->size_old[v];
    uint64_t so = g->size_old[v];
    uint64_t d = sn >= so ? sn - so : 0;
    orc_kahan_add(&g->cent[v].sum, &g->cent[v].err, (double)d / div);
    g->size_old[v] = sn;
    n_changed += ch;
  }
  n_changed_t[tid] += n_changed;
  });
  uint64_t n_changed = 0;
  for (uint64_t x : n_changed_t) n_changed += x;
  g->old_r.swap(g->new_r);  // old = new (new keeps stale values but is fully rewritten next step)
  g->changed.swap(g->new_changed);
  g->t += 1;
  g->has_changes = n_changed != 0;
  g->n_changed_last = n_changed;
  return n_changed;
}

ORC_API uint32_t orc_hb_dense_run(void* h, uint32_t max_iters) {
  Dense* g = (Dense*)h;
  while (g->has_changes && (max_iters == 0 || g->t < max_iters)) orc_hb_dense_step(h);
  return (uint32_t)g->t;
}

// registers of the current state ("old" after a step == the reference's counters.new), rank order
ORC_API void orc_hb_dense_registers(void* h, uint64_t first, uint64_t count, uint8_t* out) {
  Dense* g = (Dense*)h;
  memcpy(out, &g->old_r[first * 64], count * 64);
}
ORC_API void orc_hb_dense_ids(void* h, uint64_t* lo, uint64_t* hi) {
  Dense* g = (Dense*)h;
  for (size_t i = 0; i < g->ids.size(); i++) { lo[i] = (uint64_t)g->ids[i]; hi[i] = (uint64_t)(g->ids[i] >> 64); }
}
ORC_API void orc_hb_dense_kahan(void* h, double* sum, double* err) {
  Dense* g = (Dense*)h;
  for (size_t i = 0; i < g->cent.size(); i++) { sum[i] = g->cent[i].sum; err[i] = g->cent[i].err; }
}
// normalised result, ascending id, only > 0; returns length (call with null outputs to size)
ORC_API uint64_t orc_hb_dense_result(void* h, uint64_t* id_lo, uint64_t* id_hi, double* c) {
  Dense* g = (Dense*)h;
  const uint64_t N = g->ids.size();
  if (N == 0) return 0;
  const double norm = (double)(N - 1);
  uint64_t k = 0;
  for (uint64_t v = 0; v < N; v++) {
    double x = g->cent[v].sum;
    if (!(x > 0.0)) continue;
    x = x / norm;
    if (!std::isfinite(x)) x = 0.0;
    if (c) { id_lo[k] = (uint64_t)g->ids[v]; id_hi[k] = (uint64_t)(g->ids[v] >> 64); c[k] = x; }
    k++;
  }
  return k;
}
ORC_API void orc_hb_dense_free(void* h) { delete (Dense*)h; }
