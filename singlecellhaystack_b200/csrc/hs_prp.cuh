// Keyed pseudo-random permutation of [0, n) for the seeded randomization entry points.
//
// The reference draws, per selected gene and randomization, `sample(x = count.cells, size = nnz)`
// (R/haystack_continuous.R:275): nnz distinct cells in random order, the i-th paired with the gene's i-th
// stored value.  Here the draw is pi(0), pi(1), ..., pi(nnz - 1) for a bijection pi of [0, n) keyed by
// (seed, stream): a 4-round Feistel network on an even number of bits >= log2(n), with cycle walking
// (values outside [0, n) are re-encrypted until they fall inside).  Being invertible, the kernel can ask
// for every cell c "is pi^-1(c) < nnz, and which value does it carry" and so emits the draw already
// ordered by cell -- no ids cross PCIe and nothing has to be sorted.
//
// The arithmetic is plain uint32/uint64 so that a numpy restatement (the test checker) matches bit for bit.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define HS_HD __host__ __device__ __forceinline__
#else
#define HS_HD inline
#endif

struct HsPrp {
  uint32_t key[4];
  uint32_t n;       // domain size, 1 <= n < 2^31
  int half;         // bits per Feistel half (1..16)
  uint32_t mask;    // (1 << half) - 1
};

HS_HD uint64_t hs_mix64(uint64_t z) {   // splitmix64 finaliser
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

HS_HD HsPrp hs_prp_make(uint64_t seed, uint64_t stream, uint32_t n) {
  HsPrp p;
  int bits = 2;
  while (bits < 32 && (1ull << bits) < (uint64_t)n) ++bits;
  bits += bits & 1;
  p.half = bits >> 1;
  p.mask = (1u << p.half) - 1u;
  p.n = n;
  const uint64_t z = seed + (stream + 1ull) * 0x9E3779B97F4A7C15ull;
  const uint64_t k01 = hs_mix64(z), k23 = hs_mix64(z + 0x9E3779B97F4A7C15ull);
  p.key[0] = (uint32_t)k01;
  p.key[1] = (uint32_t)(k01 >> 32);
  p.key[2] = (uint32_t)k23;
  p.key[3] = (uint32_t)(k23 >> 32);
  return p;
}

// Round function: the top `half` bits of a multiply / xor-shift / multiply mix of (x, key).  On the device
// it is two IMADs, one shift, one xor and an IMAD.HI (h >> (32 - half) == umulhi(h, 1 << half)), so the work
// splits evenly between the fma and alu pipes.
HS_HD uint32_t hs_prp_round(uint32_t x, uint32_t key, int half) {
  uint32_t h = x * 0x9E3779B1u + key;
  h ^= h >> 15;
  h *= 0x85EBCA77u;
#if defined(__CUDA_ARCH__)
  return __umulhi(h, 1u << half);
#else
  return h >> (32 - half);
#endif
}

HS_HD uint32_t hs_prp_enc1(const HsPrp& p, uint32_t x) {
  uint32_t l = x >> p.half, r = x & p.mask;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int i = 0; i < 4; ++i) {
    const uint32_t t = l ^ hs_prp_round(r, p.key[i], p.half);
    l = r;
    r = t;
  }
  return (l << p.half) | r;
}

HS_HD uint32_t hs_prp_dec1(const HsPrp& p, uint32_t y) {
  uint32_t l = y >> p.half, r = y & p.mask;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int i = 3; i >= 0; --i) {
    const uint32_t t = r ^ hs_prp_round(l, p.key[i], p.half);
    r = l;
    l = t;
  }
  return (l << p.half) | r;
}

// pi(i), 0 <= i < n
HS_HD uint32_t hs_prp_forward(const HsPrp& p, uint32_t i) {
  uint32_t x = hs_prp_enc1(p, i);
  while (x >= p.n) x = hs_prp_enc1(p, x);
  return x;
}

// pi^-1(c), 0 <= c < n
HS_HD uint32_t hs_prp_inverse(const HsPrp& p, uint32_t c) {
  uint32_t x = hs_prp_dec1(p, c);
  while (x >= p.n) x = hs_prp_dec1(p, x);
  return x;
}
