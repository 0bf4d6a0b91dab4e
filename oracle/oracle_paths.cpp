// oracle/oracle_paths.cpp -- CPU restatement of the webgraph's shortest-path search and of ApproxHarmonic.
// TEST INFRASTRUCTURE ONLY (see oracle_common.h).  Follows, by file:line in /root/reference:
//   crates/core/src/webgraph/shortest_path.rs:57-105     dijkstra_multi (binary heap, u8 costs, the max_dist cut-off
//                                                        that fires when a node with cost > max_dist is POPPED)
//   crates/core/src/webgraph/centrality/approx_harmonic.rs:40-88   ApproxHarmonic::build for a given sample
// The graph is handed over as the unique (from, to) links the search may follow, with dense node ranks (positions in
// ascending id order), which is what tests derive from the oracle's own staging (orc_hb_dense_*).
// Parity status: the reference holds no numeric golden for either function; ApproxHarmonic's own result is not
// deterministic (random sample; f32 terms accumulated in a DashMap from a rayon pool), so parity is on the distances
// (exact) and, for a fixed sample, on the sums within f32 accumulation error.
#include "oracle_common.h"

#include <cstring>
#include <functional>
#include <queue>
#include <vector>

namespace {
struct Csr { std::vector<uint64_t> ptr; std::vector<uint32_t> adj; };
static Csr build(uint32_t n, const uint32_t* from, const uint32_t* to, uint64_t m, bool reversed) {
  Csr c; c.ptr.assign((size_t)n + 1, 0); c.adj.resize(m);
  for (uint64_t e = 0; e < m; e++) c.ptr[(reversed ? to[e] : from[e]) + 1]++;
  for (uint32_t v = 0; v < n; v++) c.ptr[v + 1] += c.ptr[v];
  std::vector<uint64_t> pos(c.ptr.begin(), c.ptr.end() - 1);
  for (uint64_t e = 0; e < m; e++) { const uint32_t a = reversed ? to[e] : from[e], b = reversed ? from[e] : to[e]; c.adj[pos[a]++] = b; }
  return c;
}
// dijkstra_multi: `dist` must be filled with 255 ("not in the map")
static void dijkstra_multi(const Csr& g, const uint32_t* sources, uint32_t n_sources, int max_dist /* < 0: None */, uint8_t* dist) {
  typedef std::pair<uint8_t, uint32_t> St;
  std::priority_queue<St, std::vector<St>, std::greater<St>> q;   // BinaryHeap of Reverse((cost, node))
  for (uint32_t i = 0; i < n_sources; i++) { q.push({0, sources[i]}); dist[sources[i]] = 0; }
  while (!q.empty()) {
    const St st = q.top(); q.pop();
    const uint8_t cost = st.first; const uint32_t v = st.second;
    if (cost > dist[v]) continue;
    if (max_dist >= 0 && cost > max_dist) return;
    for (uint64_t e = g.ptr[v]; e < g.ptr[v + 1]; e++) {
      const uint32_t w = g.adj[e];
      if ((int)cost + 1 < (int)dist[w]) { dist[w] = (uint8_t)(cost + 1); q.push({(uint8_t)(cost + 1), w}); }
    }
  }
}
}  // namespace

// one search per group: sources with group[i] == s start search s together.  dist_out [n_groups][n], 255 = not reached
ORC_API void orc_graph_distances(uint32_t n, const uint32_t* from, const uint32_t* to, uint64_t m, const uint32_t* sources,
                                 const uint32_t* group, uint32_t n_sources, uint32_t n_groups, int max_dist, int reversed, uint8_t* dist_out) {
  const Csr g = build(n, from, to, m, reversed != 0);
  memset(dist_out, 255, (size_t)n_groups * n);
  for (uint32_t s = 0; s < n_groups; s++) {
    std::vector<uint32_t> src;
    for (uint32_t i = 0; i < n_sources; i++) if (group[i] == s) src.push_back(sources[i]);
    if (!src.empty()) dijkstra_multi(g, src.data(), (uint32_t)src.size(), max_dist, dist_out + (size_t)s * n);
  }
}

// ApproxHarmonic::build for the sample `sources` (ranks), sources taken in the order given; out32 = the reference's f32
// accumulation in that order, out64 = the same f32 terms summed in f64 (what the device path computes); 0 = not reached
ORC_API void orc_approx_harmonic(uint32_t n, const uint32_t* from, const uint32_t* to, uint64_t m, const uint32_t* sources,
                                 uint32_t n_sources, int max_dist, uint64_t num_nodes, float* out32, double* out64) {
  const Csr g = build(n, from, to, m, false);
  const float nn = (float)num_nodes;
  const float norm = nn / ((float)n_sources * (nn - 1.0f));
  std::vector<uint8_t> dist(n);
  for (uint32_t v = 0; v < n; v++) { out32[v] = 0.0f; out64[v] = 0.0; }
  for (uint32_t i = 0; i < n_sources; i++) {
    std::fill(dist.begin(), dist.end(), (uint8_t)255);
    dijkstra_multi(g, sources + i, 1, max_dist, dist.data());
    for (uint32_t v = 0; v < n; v++) {
      if (dist[v] == 255 || dist[v] == 0) continue;
      const float term = (1.0f / (float)dist[v]) * norm;
      out32[v] += term; out64[v] += (double)term;
    }
  }
}
