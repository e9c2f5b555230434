// rmat.h — counter-based Graph500 R-MAT edge generator, identical on host and
// device: (A,B,C,D) = (.57,.19,.19,.05), one independent
// splitmix64 stream per edge index, ids scrambled by a fixed bijection of
// [0, 2^scale), self loops and duplicates kept.  The reference ships no
// generator; this is the synthetic-input definition used by bench.py for both
// the CUDA path and the CPU baseline.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define GL_HD __host__ __device__ __forceinline__
#else
#define GL_HD inline
#endif

namespace gl {

GL_HD uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

// bijection of [0, 2^scale): odd multiply and xor-shift-right are both
// invertible modulo 2^scale.
GL_HD uint64_t rmat_scramble(uint64_t v, int scale, uint64_t seed) {
  const uint64_t mask = (scale >= 64) ? ~0ull : ((1ull << scale) - 1);
  const uint64_t k0 = splitmix64(seed ^ 0x5851F42D4C957F2Dull) | 1ull;
  const uint64_t k1 = splitmix64(seed ^ 0x14057B7EF767814Full) | 1ull;
  const int s = scale > 1 ? scale / 2 : 1;
  v = (v * k0 + (k1 >> 1)) & mask;
  v ^= v >> s;
  v = (v * k1) & mask;
  v ^= v >> (s + 1 < scale ? s + 1 : s);
  v = (v * 0x9E3779B97F4A7C15ull) & mask;
  v ^= v >> s;
  return v & mask;
}

// edge e of the graph (seed): unscrambled R-MAT descent then scramble.
GL_HD void rmat_edge(uint64_t e, int scale, uint64_t seed, uint64_t* src,
                     uint64_t* dst) {
  uint64_t state = splitmix64(seed ^ (e * 0xD1342543DE82EF95ull));
  uint64_t s = 0, d = 0;
  // thresholds on a 32-bit uniform: A=.57, A+B=.76, A+B+C=.95
  const uint32_t tA = 2448131359u, tAB = 3264175145u, tABC = 4080218931u;
  uint64_t r = 0;
  for (int level = 0; level < scale; ++level) {
    uint32_t t;
    if ((level & 1) == 0) {
      state = splitmix64(state);
      r = state;
      t = (uint32_t) (r >> 32);
    } else {
      t = (uint32_t) r;
    }
    uint32_t sb = (t >= tAB) ? 1u : 0u;               // C or D -> lower half
    uint32_t db = (t >= tA && t < tAB) || (t >= tABC) ? 1u : 0u;  // B or D
    s = (s << 1) | sb;
    d = (d << 1) | db;
  }
  *src = rmat_scramble(s, scale, seed + 1);
  *dst = rmat_scramble(d, scale, seed + 1);
}

// weight of edge e: mode 1 integer-valued 1..255, mode 2 real in (0,1]
GL_HD float rmat_weight(uint64_t e, uint64_t seed, int mode) {
  uint64_t r = splitmix64((seed + 2) ^ (e * 0xA0761D6478BD642Full));
  if (mode == 1) return (float) (1 + (r >> 32) % 255);
  return (float) ((r >> 40) + 1) * (1.0f / 16777216.0f);
}

}  // namespace gl
