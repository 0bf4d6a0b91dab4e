/* oracle/ -- CPU restatement of the reference's algorithms.  TEST INFRASTRUCTURE ONLY.
 *
 * Nothing in the product path (stract_b200/, include/, libstract_b200.so) may include,
 * link or call anything in this directory.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs use it, and only as the checker or
 * as the timed CPU baseline.
 *
 * Parity status: the reference is 100% Rust and no Rust toolchain exists in this image,
 * so the oracle cannot be diffed against a reference binary.  It is pinned against every
 * known-answer test the reference holds for the two hot paths; see
 * tests/test_oracle_*.py.  Where the reference only has property tests (HLL<64>::size,
 * BitPacker4x byte layout) the header of the relevant file says "parity unpinned".
 */
#ifndef SB200_ORACLE_COMMON_H
#define SB200_ORACLE_COMMON_H
#include <stdint.h>
#include <stddef.h>

typedef unsigned __int128 u128;

#ifdef __cplusplus
#define ORC_API extern "C" __attribute__((visibility("default")))
#else
#define ORC_API __attribute__((visibility("default")))
#endif

static inline u128 orc_make_u128(uint64_t hi, uint64_t lo) { return ((u128)hi << 64) | (u128)lo; }

/* Rust `x as usize` for f64: saturating, NaN -> 0 (reference relies on it in
 * crates/core/src/hyperloglog.rs:4512-4515). */
static inline uint64_t orc_f64_as_u64(double x) {
  if (!(x == x)) return 0;
  if (x <= 0.0) return 0;
  if (x >= 18446744073709551616.0) return UINT64_MAX;
  return (uint64_t)x;
}

#endif
