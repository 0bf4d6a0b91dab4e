// oracle/oracle_hyperball_mt.cpp -- full-size staging for the dense CPU restatement.
// TEST INFRASTRUCTURE ONLY (see oracle_common.h).
//
// orc_hb_dense_create (oracle_hyperball.cpp) sorts every endpoint and every edge on one thread, which is
// fine up to ~10^8 edges.  bench.py checks the CUDA path against the oracle at BASELINE configs[1] size
// (10^9 edges, 2.8*10^7 nodes), so this file builds the SAME Dense object with all host threads:
//   nodes  = every endpoint of every edge, skipped or not        store.rs:338-357
//   edges  = unique (from,to), the FIRST occurrence's flags decide  store.rs:313 (unique_by)
//   kept   = !(rel & skip_mask)                                    harmonic.rs:36-49
// and returns a handle the orc_hb_dense_* functions accept.  tests/test_oracle_path1.py checks it against
// orc_hb_dense_create on random inputs (ids, CSR rows as sets, per-iteration registers).
//
// The second half is a CPU port of the repo's own synthetic edge generator (stract_b200/synth.py ==
// csrc/synth.cu), used by `bench.py --impl reference` to materialise the 10^9-edge stream without a GPU.
#include "oracle_common.h"
#include "oracle_dense.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

extern "C" void orc_hll_add(uint8_t* regs, int n, uint64_t item);
extern "C" uint64_t orc_hll_size(const uint8_t* regs, int n);

namespace {
template <class F> void par_for(int64_t n, int threads, int64_t grain, F f) {
  if (threads <= 1 || n <= grain) { if (n > 0) f(0, n, 0); return; }
  std::atomic<int64_t> next(0);
  std::vector<std::thread> pool;
  for (int t = 0; t < threads; t++)
    pool.emplace_back([&, t]() {
      for (;;) { int64_t b = next.fetch_add(grain); if (b >= n) break; f(b, std::min(n, b + grain), t); }
    });
  for (auto& th : pool) th.join();
}
// one contiguous slice per thread (keeps stream order inside a slice)
template <class F> void par_slices(uint64_t n, int threads, F f) {
  std::vector<std::thread> pool;
  for (int t = 0; t < threads; t++) {
    const uint64_t b = n * (uint64_t)t / threads, e = n * (uint64_t)(t + 1) / threads;
    pool.emplace_back([=]() { f(b, e, t); });
  }
  for (auto& th : pool) th.join();
}

struct LocalSet {  // open addressing over u128; the all-zero key is kept in a flag
  std::vector<u128> slot;
  uint64_t used = 0, mask = 0;
  bool has_zero = false;
  static uint64_t mix(u128 x) {
    uint64_t a = (uint64_t)x ^ ((uint64_t)(x >> 64) * 0x9E3779B97F4A7C15ull);
    a ^= a >> 32; a *= 0xD6E8FEB86659FD93ull; a ^= a >> 29;
    return a;
  }
  LocalSet() { slot.assign(1024, 0); mask = 1023; }
  void grow() {
    std::vector<u128> old; old.swap(slot);
    slot.assign(old.size() * 2, 0); mask = slot.size() - 1; used = 0;
    for (u128 k : old) if (k) put(k);
  }
  void put(u128 k) {
    uint64_t i = mix(k) & mask;
    while (slot[i] != 0) { if (slot[i] == k) return; i = (i + 1) & mask; }
    slot[i] = k; used++;
  }
  void insert(u128 k) {
    if (k == 0) { has_zero = true; return; }
    if ((used + 1) * 2 > slot.size()) grow();
    put(k);
  }
};
constexpr int ID_BUCKET_BITS = 12;  // ids are split on their top bits: ordered, independent buckets
inline uint32_t id_bucket(u128 x) { return (uint32_t)(x >> (128 - ID_BUCKET_BITS)); }
constexpr int IDX_BITS = 16;  // rank lookup: directory over the top 16 bits, then a short binary search
}  // namespace

ORC_API void* orc_hb_dense_create_mt(const uint64_t* flo, const uint64_t* fhi, const uint64_t* tlo,
                                     const uint64_t* thi, const uint64_t* rel, uint64_t n_edges,
                                     uint64_t skip_mask, int threads) {
  Dense* g = new Dense();
  const int T = threads < 1 ? 1 : threads;
  g->threads = T;
  const uint64_t E = n_edges;
  const bool timing = getenv("ORC_TIMING") != nullptr;
  auto t_prev = std::chrono::steady_clock::now();
  auto mark = [&](const char* what) {
    if (!timing) return;
    const auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "[orc stage] %-28s %8.2f s\n", what, std::chrono::duration<double>(now - t_prev).count());
    t_prev = now;
  };

  // ---- 1. host_nodes(): distinct endpoints ------------------------------------------------------
  const int NB = 1 << ID_BUCKET_BITS;
  std::vector<std::vector<u128>> local(T);             // per thread: its distinct endpoints, grouped by bucket
  std::vector<std::vector<uint64_t>> local_off(T);     // [NB+1]
  par_slices(E, T, [&](uint64_t b, uint64_t e, int t) {
    LocalSet s;
    for (uint64_t i = b; i < e; i++) { s.insert(orc_make_u128(fhi[i], flo[i])); s.insert(orc_make_u128(thi[i], tlo[i])); }
    std::vector<uint64_t> cnt(NB + 1, 0);
    if (s.has_zero) cnt[1]++;
    for (u128 k : s.slot) if (k) cnt[id_bucket(k) + 1]++;
    for (int q = 0; q < NB; q++) cnt[q + 1] += cnt[q];
    std::vector<u128> out(cnt[NB]);
    std::vector<uint64_t> pos(cnt.begin(), cnt.end() - 1);
    if (s.has_zero) out[pos[0]++] = 0;
    for (u128 k : s.slot) if (k) out[pos[id_bucket(k)]++] = k;
    local[t].swap(out); local_off[t].swap(cnt);
  });
  mark("1a local endpoint sets");
  std::vector<std::vector<u128>> bucket(NB);
  par_for(NB, T, 8, [&](int64_t qb, int64_t qe, int) {
    for (int64_t q = qb; q < qe; q++) {
      std::vector<u128>& v = bucket[q];
      for (int t = 0; t < T; t++) v.insert(v.end(), local[t].begin() + local_off[t][q], local[t].begin() + local_off[t][q + 1]);
      std::sort(v.begin(), v.end());
      v.erase(std::unique(v.begin(), v.end()), v.end());
    }
  });
  local.clear(); local.shrink_to_fit();
  std::vector<uint64_t> boff(NB + 1, 0);
  for (int q = 0; q < NB; q++) boff[q + 1] = boff[q] + bucket[q].size();
  const uint64_t N = boff[NB];
  g->ids.resize(N);
  par_for(NB, T, 8, [&](int64_t qb, int64_t qe, int) {
    for (int64_t q = qb; q < qe; q++) if (!bucket[q].empty()) memcpy(&g->ids[boff[q]], bucket[q].data(), bucket[q].size() * sizeof(u128));
  });
  bucket.clear(); bucket.shrink_to_fit();
  if (N > 0xFFFFFFFFull) { delete g; return nullptr; }

  mark("1b bucket sort + concat");
  // ---- 2. rank directory -----------------------------------------------------------------------
  const uint64_t NI = 1ull << IDX_BITS;
  std::vector<uint64_t> dir(NI + 1, 0);
  {
    for (uint64_t v = 0; v < N; v++) dir[(uint64_t)(g->ids[v] >> (128 - IDX_BITS)) + 1]++;
    for (uint64_t q = 0; q < NI; q++) dir[q + 1] += dir[q];
  }
  const u128* ids = g->ids.data();
  auto rank = [&](u128 x) -> uint32_t {
    const uint64_t q = (uint64_t)(x >> (128 - IDX_BITS));
    return (uint32_t)(std::lower_bound(ids + dir[q], ids + dir[q + 1], x) - ids);
  };

  // ---- 3. edges -> (to,from) ranks, partitioned by destination range --------------------------------
  struct Ent { uint64_t key, pos_skip; };  // key = to<<32 | from ; pos_skip = stream position << 1 | skipped
  const int TB = 4096;
  const uint64_t rows_per = N ? (N + TB - 1) / TB : 1;
  std::unique_ptr<uint32_t[]> tr(new uint32_t[E ? E : 1]), fr(new uint32_t[E ? E : 1]);  // uninitialised: first touched by the worker threads
  std::vector<std::vector<uint64_t>> cnt(T, std::vector<uint64_t>(TB, 0));
  par_slices(E, T, [&](uint64_t b, uint64_t e, int t) {
    for (uint64_t i = b; i < e; i++) {
      const uint32_t a = rank(orc_make_u128(thi[i], tlo[i])), f = rank(orc_make_u128(fhi[i], flo[i]));
      tr[i] = a; fr[i] = f; cnt[t][a / rows_per]++;
    }
  });
  mark("3a rank lookups");
  std::vector<uint64_t> tb_off(TB + 1, 0);
  {
    uint64_t run = 0;
    for (int q = 0; q < TB; q++) {
      tb_off[q] = run;
      for (int t = 0; t < T; t++) { const uint64_t c = cnt[t][q]; cnt[t][q] = run; run += c; }  // thread-major inside a bucket: stream order
    }
    tb_off[TB] = run;
  }
  std::unique_ptr<Ent[]> ents(new Ent[E ? E : 1]);
  par_slices(E, T, [&](uint64_t b, uint64_t e, int t) {
    std::vector<uint64_t>& p = cnt[t];
    for (uint64_t i = b; i < e; i++) {
      const uint32_t a = tr[i];
      ents[p[a / rows_per]++] = {((uint64_t)a << 32) | fr[i], (i << 1) | (uint64_t)((rel[i] & skip_mask) != 0)};
    }
  });
  tr.reset(); fr.reset();
  mark("3b scatter");

  // ---- 4. per bucket: sort, first-wins dedup, drop skipped, emit CSR pieces ---------------------------
  g->row_ptr.assign(N + 1, 0);
  std::vector<std::vector<uint32_t>> colb(TB);
  par_for(TB, T, 4, [&](int64_t qb, int64_t qe, int) {
    for (int64_t q = qb; q < qe; q++) {
      Ent* b = ents.get() + tb_off[q]; Ent* e = ents.get() + tb_off[q + 1];
      std::sort(b, e, [](const Ent& x, const Ent& y) { return x.key != y.key ? x.key < y.key : x.pos_skip < y.pos_skip; });
      std::vector<uint32_t>& c = colb[q];
      for (Ent* p = b; p < e; p++) {
        if (p > b && p->key == (p - 1)->key) continue;   // a later duplicate of (from,to)
        if (p->pos_skip & 1) continue;                   // the first occurrence is a skipped rel
        c.push_back((uint32_t)p->key);
        g->row_ptr[(p->key >> 32) + 1]++;
      }
    }
  });
  ents.reset();
  mark("4a bucket sort + dedup");
  for (uint64_t v = 0; v < N; v++) g->row_ptr[v + 1] += g->row_ptr[v];
  g->col.resize(g->row_ptr[N]);
  par_for(TB, T, 4, [&](int64_t qb, int64_t qe, int) {
    for (int64_t q = qb; q < qe; q++) {
      const uint64_t first_row = std::min<uint64_t>((uint64_t)q * rows_per, N);
      if (!colb[q].empty()) memcpy(&g->col[g->row_ptr[first_row]], colb[q].data(), colb[q].size() * 4);
    }
  });
  colb.clear(); colb.shrink_to_fit();

  mark("4b csr concat");
  // ---- 5. HyperBall state (harmonic.rs:53-73) ----------------------------------------------------
  g->old_r.assign(N * 64, 0);
  g->size_old.resize(N);
  par_for((int64_t)N, T, 1 << 16, [&](int64_t vb, int64_t ve, int) {
    for (int64_t v = vb; v < ve; v++) {
      orc_hll_add(&g->old_r[v * 64], 64, (uint64_t)g->ids[v]);
      g->size_old[v] = orc_hll_size(&g->old_r[v * 64], 64);
    }
  });
  g->new_r = g->old_r;
  g->changed.assign(N, 1);
  g->new_changed.assign(N, 0);
  g->cent.assign(N, Kahan());
  mark("5 state init");
  return g;
}

// back to the freshly-created state (bench: several timed runs over one staged graph)
ORC_API void orc_hb_dense_reset(void* h) {
  Dense* g = (Dense*)h;
  const uint64_t N = g->ids.size();
  std::fill(g->old_r.begin(), g->old_r.end(), 0);
  par_for((int64_t)N, g->threads, 1 << 16, [&](int64_t vb, int64_t ve, int) {
    for (int64_t v = vb; v < ve; v++) {
      orc_hll_add(&g->old_r[v * 64], 64, (uint64_t)g->ids[v]);
      g->size_old[v] = orc_hll_size(&g->old_r[v * 64], 64);
    }
  });
  g->new_r = g->old_r;
  std::fill(g->changed.begin(), g->changed.end(), 1);
  std::fill(g->new_changed.begin(), g->new_changed.end(), 0);
  std::fill(g->cent.begin(), g->cent.end(), Kahan());
  g->t = 0; g->has_changes = true; g->n_changed_last = 0;
}

// (kept edges that are self-loops: the CUDA path drops them as no-op merges, the oracle keeps them)
ORC_API uint64_t orc_hb_dense_num_self_loops(void* h) {
  Dense* g = (Dense*)h;
  const int64_t N = (int64_t)g->ids.size();
  std::vector<uint64_t> part(g->threads, 0);
  par_for(N, g->threads, 1 << 14, [&](int64_t vb, int64_t ve, int t) {
    uint64_t c = 0;
    for (int64_t v = vb; v < ve; v++) for (uint64_t e = g->row_ptr[v]; e < g->row_ptr[v + 1]; e++) c += g->col[e] == (uint32_t)v;
    part[t] += c;
  });
  uint64_t s = 0; for (uint64_t x : part) s += x;
  return s;
}
// zero-copy views for full-size comparisons (valid until the next step / free)
ORC_API const uint8_t* orc_hb_dense_registers_ptr(void* h) { return ((Dense*)h)->old_r.data(); }

// ---------------------------------------------------------------- synthetic edge stream -----------
// CPU port of stract_b200/csrc/synth.cu (== stract_b200/synth.py): kind 0 uniform, kind 1 R-MAT
// (0.57,0.19,0.19,0.05) folded mod n_nodes; ids = splitmix64(7, 2j / 2j+1); 10 % NOFOLLOW from splitmix64(9, i).
static inline uint64_t splitmix64(uint64_t seed, uint64_t i) {
  uint64_t z = seed + (i + 1) * 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
ORC_API void orc_synth_edges(int kind, uint64_t n_nodes, uint64_t first, uint64_t count, uint64_t seed, int scale,
                             uint64_t* flo, uint64_t* fhi, uint64_t* tlo, uint64_t* thi, uint64_t* rel, int threads) {
  par_for((int64_t)count, threads < 1 ? 1 : threads, 1 << 18, [&](int64_t kb, int64_t ke, int) {
    for (int64_t k = kb; k < ke; k++) {
      const uint64_t i = first + (uint64_t)k;
      uint64_t f = 0, t = 0;
      if (kind == 0) {
        f = splitmix64(seed, 2 * i) % n_nodes;
        t = splitmix64(seed, 2 * i + 1) % n_nodes;
      } else {
        for (int w = 0; w * 4 < scale; w++) {
          const uint64_t r = splitmix64(seed, 7 * i + w);
          for (int q = 0; q < 4 && w * 4 + q < scale; q++) {
            const uint32_t x = (uint32_t)(r >> (16 * q)) & 0xFFFFu;
            const uint64_t fb = x >= 49807u;
            const uint64_t tb = ((x >= 37356u) && (x < 49807u)) || (x >= 62259u);
            f = (f << 1) | fb; t = (t << 1) | tb;
          }
        }
        f %= n_nodes; t %= n_nodes;
      }
      fhi[k] = splitmix64(7, 2 * f); flo[k] = splitmix64(7, 2 * f + 1);
      thi[k] = splitmix64(7, 2 * t); tlo[k] = splitmix64(7, 2 * t + 1);
      rel[k] = (splitmix64(9, i) % 10 == 0) ? (1ull << 8) : 0ull;
    }
  });
}
