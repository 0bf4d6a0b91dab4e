/* oracle/oracle_dense.h -- the flat-array HyperBall state shared by oracle_hyperball.cpp (single-thread staging,
 * step loop) and oracle_hyperball_mt.cpp (all-threads staging for full-size runs).  TEST INFRASTRUCTURE ONLY. */
#ifndef SB200_ORACLE_DENSE_H
#define SB200_ORACLE_DENSE_H
#include <vector>
#include "oracle_common.h"

struct Kahan { double sum = 0, err = 0; };

// Same math as the reference over dense node ranks (rank = position in ascending u128 order), synchronous
// update new[v] = max(old[v], max_{u->v kept, u changed} old[u]); steppable.
struct Dense {
  std::vector<u128> ids;            // ascending
  std::vector<uint64_t> row_ptr;    // dst-major CSR over kept, unique, non-skipped edges
  std::vector<uint32_t> col;
  std::vector<uint8_t> old_r, new_r;  // N x 64
  std::vector<uint8_t> changed, new_changed;
  std::vector<Kahan> cent;
  std::vector<uint64_t> size_old;
  uint64_t t = 0;
  bool has_changes = true;
  uint64_t n_changed_last = 0;
  int threads = 1;
};
#endif
