// rmat_gen.h — the synthetic-input DEFINITION of bench.py, restated for the CPU
// arm (TEST / MEASUREMENT INFRASTRUCTURE; plain C++, no CUDA, no product code).
//
// Graph500-style R-MAT, SURVEY.md 8(d): (A,B,C,D) = (.57,.19,.19,.05), one
// counter-based splitmix64 stream per edge index, vertex ids scrambled by a
// fixed bijection of [0, 2^scale), self loops and duplicates kept, weights
// 1..255 (mode 1) or (0,1] (mode 2).  The reference ships no generator.  The
// GPU arm generates the same edges on the device
// (libgrape-lite_b200/csrc/rmat.h); tests/test_rmat_def.py checks that the two
// statements of the definition produce identical edge lists, so the CPU arm
// (oracle/_ref/ref_driver --rmat ...) never has to load the product library.
#pragma once
#include <stdint.h>

namespace rmatdef {

inline uint64_t mix(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

// bijection of [0, 2^scale): odd multiplications and xor-shifts are invertible mod 2^scale
inline uint64_t scramble(uint64_t v, int scale, uint64_t seed) {
  const uint64_t mask = scale >= 64 ? ~0ull : ((1ull << scale) - 1);
  const uint64_t k0 = mix(seed ^ 0x5851F42D4C957F2Dull) | 1ull;
  const uint64_t k1 = mix(seed ^ 0x14057B7EF767814Full) | 1ull;
  const int s = scale > 1 ? scale / 2 : 1;
  const int s2 = s + 1 < scale ? s + 1 : s;
  v = (v * k0 + (k1 >> 1)) & mask;
  v ^= v >> s;
  v = (v * k1) & mask;
  v ^= v >> s2;
  v = (v * 0x9E3779B97F4A7C15ull) & mask;
  v ^= v >> s;
  return v & mask;
}

inline void edge(uint64_t e, int scale, uint64_t seed, uint64_t* src, uint64_t* dst) {
  // quadrant thresholds on a 32-bit uniform: A = .57, A+B = .76, A+B+C = .95
  const uint32_t tA = 2448131359u, tAB = 3264175145u, tABC = 4080218931u;
  uint64_t st = mix(seed ^ (e * 0xD1342543DE82EF95ull));
  uint64_t s = 0, d = 0, r = 0;
  for (int level = 0; level < scale; ++level) {
    uint32_t t;
    if ((level & 1) == 0) {   // one 64-bit draw feeds two levels
      st = mix(st);
      r = st;
      t = (uint32_t) (r >> 32);
    } else {
      t = (uint32_t) r;
    }
    const uint32_t lower = t >= tAB ? 1u : 0u;                            // quadrants C, D
    const uint32_t right = ((t >= tA && t < tAB) || t >= tABC) ? 1u : 0u;  // quadrants B, D
    s = (s << 1) | lower;
    d = (d << 1) | right;
  }
  *src = scramble(s, scale, seed + 1);
  *dst = scramble(d, scale, seed + 1);
}

inline float weight(uint64_t e, uint64_t seed, int mode) {
  const uint64_t r = mix((seed + 2) ^ (e * 0xA0761D6478BD642Full));
  if (mode == 1) return (float) (1 + (r >> 32) % 255);
  return (float) ((r >> 40) + 1) * (1.0f / 16777216.0f);
}

}  // namespace rmatdef
