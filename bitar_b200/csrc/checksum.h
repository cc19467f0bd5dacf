// checksum.h -- CRC-32 / Adler-32 pieces fused into the deflate and inflate kernels.
//
// The reference requests the checksum through BlueFieldConfiguration::checksum_type
// (/root/reference/src/include/config.h:169-182 -> rte_comp_xform.checksum, src/config.cc:86-88) and the
// PMD returns it in rte_comp_op::output_chksum.  Here every thread checksums a contiguous slice of the
// chunk and the partial results are combined algebraically:
//   CRC-32 : register states combine over GF(2):  S(A||B) = S(A) * x^(8|B|) mod P  xor  S0(B)
//   Adler-32: A = 1 + sum b_i,  B = n + sum (n - i) b_i   (mod 65521)
// BITAR_HD so the arithmetic is unit-tested on the CPU against zlib (tests/test_core_host.py).
#pragma once
#include <stdint.h>

#include "deflate_common.h"

namespace bitar {
namespace cks {

constexpr uint32_t kPoly = 0xEDB88320u;  // reflected IEEE 802.3 polynomial (zlib convention)
constexpr uint32_t kAdlerMod = 65521u;

BITAR_HD uint32_t crc_table_entry(uint32_t i) {
  uint32_t c = i;
  for (int k = 0; k < 8; ++k) c = (c >> 1) ^ (kPoly & (0u - (c & 1u)));
  return c;
}

// register update for one byte (tab = 256-entry table, normally in shared memory)
BITAR_HD uint32_t crc_step(uint32_t state, uint32_t byte, const uint32_t* tab) {
  return tab[(state ^ byte) & 0xFFu] ^ (state >> 8);
}

// a(x) * b(x) mod P in the reflected representation (x^0 == 0x80000000)
BITAR_HD uint32_t crc_mulmod(uint32_t a, uint32_t b) {
  uint32_t p = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll 8
#endif
  for (int i = 31; i >= 0; --i) {
    p ^= b & (0u - ((a >> i) & 1u));
    b = (b >> 1) ^ (kPoly & (0u - (b & 1u)));
  }
  return p;
}

// x2n[k] = x^(2^k) mod P, k = 0..31
BITAR_HD void crc_x2n_init(uint32_t* x2n) {
  uint32_t p = 0x40000000u;  // x^1
  x2n[0] = p;
  for (int k = 1; k < 32; ++k) {
    p = crc_mulmod(p, p);
    x2n[k] = p;
  }
}

// x^n mod P
BITAR_HD uint32_t crc_xpow(uint32_t n, const uint32_t* x2n) {
  uint32_t p = 0x80000000u;
  for (int k = 0; n; ++k, n >>= 1)
    if (n & 1u) p = crc_mulmod(p, x2n[k]);
  return p;
}

// Partial sums of one slice for both checksums.
struct Partial {
  uint32_t crc_state;  // CRC register after the slice (start state 0, or 0xFFFFFFFF for the first slice)
  uint32_t s1;         // sum of bytes
  uint32_t s2;         // sum of (slice_len - j) * b_j  (position weights inside the slice)
};

// Contribution of a slice that is followed by `tail` more bytes.  XOR the crc parts, add the adler parts.
BITAR_HD uint32_t crc_contrib(uint32_t state, uint32_t tail_bytes, const uint32_t* x2n) {
  return tail_bytes ? crc_mulmod(state, crc_xpow(8u * tail_bytes, x2n)) : state;
}
BITAR_HD uint32_t adler_b_contrib(uint32_t s1, uint32_t s2, uint32_t tail_bytes) {
  uint64_t v = (uint64_t)s2 + (uint64_t)tail_bytes * (uint64_t)s1;
  return (uint32_t)(v % kAdlerMod);
}
BITAR_HD uint32_t adler_finish(uint32_t sum_s1, uint32_t sum_b, uint32_t n) {
  uint32_t a = (1u + sum_s1 % kAdlerMod) % kAdlerMod;
  uint32_t b = (n % kAdlerMod + sum_b % kAdlerMod) % kAdlerMod;
  return (b << 16) | a;
}

// rte_comp_op::output_chksum packing: CRC-32 in the low word, Adler-32 in the high word
BITAR_HD uint64_t pack(uint32_t crc, uint32_t adler) { return (uint64_t)crc | ((uint64_t)adler << 32); }

}  // namespace cks
}  // namespace bitar
