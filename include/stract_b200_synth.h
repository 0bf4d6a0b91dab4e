/* stract_b200_synth.h -- device-side generators of the synthetic benchmark inputs
 * (SURVEY.md 8d / stract_b200/synth.py define the streams; these produce the same values
 * directly in HBM so the 1B-edge configuration never has to exist in host memory).
 * Benchmark support, not part of the drop-in boundary. */
#ifndef STRACT_B200_SYNTH_H
#define STRACT_B200_SYNTH_H
#include "stract_b200.h"
#ifdef __cplusplus
extern "C" {
#endif
/* kind 0: uniform (config C1), kind 1: R-MAT (0.57,0.19,0.19,0.05) with `scale` levels folded mod
 * n_nodes (configs C2/C3).  Writes edges [first, first+count) of the stream into the five DEVICE
 * arrays (each `count` u64).  Ids: hi = splitmix64(7,2j), lo = splitmix64(7,2j+1); flags: NOFOLLOW
 * for 10 % of the edges (splitmix64(9,i) % 10 == 0). */
SB200_API int sb200_synth_edges(int kind, uint64_t n_nodes, uint64_t first, uint64_t count, uint64_t seed,
                                int scale, int device, uint64_t* from_lo, uint64_t* from_hi,
                                uint64_t* to_lo, uint64_t* to_hi, uint64_t* rel_flags);
#ifdef __cplusplus
}
#endif
#endif
