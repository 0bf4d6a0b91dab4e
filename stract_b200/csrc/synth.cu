// synth.cu -- device-side generators of the synthetic webgraph streams (see stract_b200_synth.h).
#include "common.cuh"
#include "../../include/stract_b200_synth.h"

namespace sb200 {
__device__ __forceinline__ uint64_t splitmix64(uint64_t seed, uint64_t i) {
  uint64_t z = seed + (i + 1) * 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__global__ void k_synth(int kind, uint64_t n_nodes, uint64_t first, uint64_t count, uint64_t seed, int scale,
                        uint64_t* flo, uint64_t* fhi, uint64_t* tlo, uint64_t* thi, uint64_t* rel) {
  const uint64_t k = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (k >= count) return;
  const uint64_t i = first + k;
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
}  // namespace sb200

extern "C" int sb200_synth_edges(int kind, uint64_t n_nodes, uint64_t first, uint64_t count, uint64_t seed, int scale,
                                 int device, uint64_t* from_lo, uint64_t* from_hi, uint64_t* to_lo, uint64_t* to_hi,
                                 uint64_t* rel_flags) {
  using namespace sb200;
  if (n_nodes == 0 || scale < 0 || scale > 28) SB_FAIL(SB200_EINVAL, "bad generator arguments");
  SB_CUDA(cudaSetDevice(device));
  if (count == 0) return SB200_OK;
  SB_LAUNCH(k_synth, div_up(count, 256), 256, 0, 0, kind, n_nodes, first, count, seed, scale, from_lo, from_hi, to_lo,
            to_hi, rel_flags);
  SB_CHECK_LAUNCH();
  SB_CUDA(cudaDeviceSynchronize());
  return SB200_OK;
}
