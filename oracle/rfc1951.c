/*
 * rfc1951.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement of the published algorithms the reference's codec path
 * relies on but does not contain (they live in zlib, reached through DPDK's
 * compress_zlib PMD, see bitar_oracle.c):
 *   - RFC 1951 raw DEFLATE decoding (stored / fixed / dynamic blocks),
 *   - CRC-32 (IEEE 802.3, reflected poly 0xEDB88320, zlib convention) and
 *     Adler-32 (RFC 1950), the checksum types selectable through
 *     /root/reference/src/include/config.h:169-182.
 *
 * It is deliberately independent of libz so that GPU-produced streams are checked
 * by two decoders (this one reports *where* a stream is invalid), and so that the
 * zlib-backed oracle itself can be cross-checked (tests/test_oracle.py).
 * It is bit-serial and slow on purpose: clarity over speed.
 */
#include <stdint.h>
#include <stddef.h>
#include <string.h>

#define ORACLE_API __attribute__((visibility("default")))

/* error codes (negative) */
enum {
  RFC_E_TRUNCATED = -1,     /* ran out of input */
  RFC_E_BTYPE = -2,         /* BTYPE == 3 */
  RFC_E_STORED_LEN = -3,    /* LEN != ~NLEN */
  RFC_E_OUTPUT_FULL = -4,   /* output capacity exceeded */
  RFC_E_BAD_CODE = -5,      /* code not in tree (incomplete set hit) */
  RFC_E_OVERSUBSCRIBED = -6,/* code length set over-subscribed */
  RFC_E_INCOMPLETE = -7,    /* incomplete set (other than the single-code case) */
  RFC_E_BAD_LENGTHS = -8,   /* repeat with no previous length / too many lengths */
  RFC_E_BAD_SYMBOL = -9,    /* litlen symbol 286/287 or dist symbol 30/31 */
  RFC_E_DIST_TOO_FAR = -10, /* distance beyond start of output */
  RFC_E_NO_EOB = -11        /* dynamic block without end-of-block code */
};

typedef struct {
  const uint8_t* in;
  size_t in_len;
  size_t bitpos; /* absolute bit position */
} bitreader;

static int getbit(bitreader* br) {
  if ((br->bitpos >> 3) >= br->in_len) return -1;
  int b = (br->in[br->bitpos >> 3] >> (br->bitpos & 7)) & 1;
  br->bitpos++;
  return b;
}

/* RFC 1951 3.1.1: data elements are packed LSB first */
static long getbits(bitreader* br, int n) {
  long v = 0;
  for (int i = 0; i < n; ++i) {
    int b = getbit(br);
    if (b < 0) return -1;
    v |= (long)b << i;
  }
  return v;
}

#define MAXBITS 15

typedef struct {
  uint16_t count[MAXBITS + 1];
  uint16_t symbol[288];
} hufftree;

/* Canonical code construction, RFC 1951 3.2.2.  Returns 0 complete, >0 incomplete
 * (left-over code space), <0 over-subscribed. */
static int build_tree(hufftree* t, const uint8_t* lengths, int n) {
  uint16_t offs[MAXBITS + 1];
  memset(t->count, 0, sizeof t->count);
  for (int i = 0; i < n; ++i) t->count[lengths[i]]++;
  if (t->count[0] == n) return 0; /* no codes: complete but unusable */
  int left = 1;
  for (int len = 1; len <= MAXBITS; ++len) {
    left <<= 1;
    left -= t->count[len];
    if (left < 0) return left;
  }
  offs[1] = 0;
  for (int len = 1; len < MAXBITS; ++len) offs[len + 1] = (uint16_t)(offs[len] + t->count[len]);
  for (int i = 0; i < n; ++i)
    if (lengths[i]) t->symbol[offs[lengths[i]]++] = (uint16_t)i;
  return left;
}

/* Huffman codes are packed MSB first (RFC 1951 3.1.1): walk bit by bit. */
static int decode_sym(bitreader* br, const hufftree* t) {
  int code = 0, first = 0, index = 0;
  for (int len = 1; len <= MAXBITS; ++len) {
    int b = getbit(br);
    if (b < 0) return RFC_E_TRUNCATED;
    code |= b;
    int count = t->count[len];
    if (code - count < first) return t->symbol[index + (code - first)];
    index += count;
    first += count;
    first <<= 1;
    code <<= 1;
  }
  return RFC_E_BAD_CODE;
}

static const uint16_t kLenBase[29] = {3,  4,  5,  6,  7,  8,  9,  10, 11,  13,  15,  17,  19,  23, 27,
                                      31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
static const uint8_t kLenExtra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2,
                                      2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
static const uint16_t kDistBase[30] = {1,   2,   3,   4,   5,   7,    9,    13,   17,   25,
                                       33,  49,  65,  97,  129, 193,  257,  385,  513,  769,
                                       1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
static const uint8_t kDistExtra[30] = {0, 0, 0, 0, 1, 1, 2, 2,  3,  3,  4,  4,  5,  5,  6,
                                       6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
static const uint8_t kClOrder[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

static int inflate_codes(bitreader* br, uint8_t* out, size_t cap, size_t* outpos,
                         const hufftree* lt, const hufftree* dt) {
  for (;;) {
    int sym = decode_sym(br, lt);
    if (sym < 0) return sym;
    if (sym < 256) {
      if (*outpos >= cap) return RFC_E_OUTPUT_FULL;
      out[(*outpos)++] = (uint8_t)sym;
    } else if (sym == 256) {
      return 0;
    } else {
      sym -= 257;
      if (sym >= 29) return RFC_E_BAD_SYMBOL;
      long eb = getbits(br, kLenExtra[sym]);
      if (eb < 0) return RFC_E_TRUNCATED;
      int len = kLenBase[sym] + (int)eb;
      int ds = decode_sym(br, dt);
      if (ds < 0) return ds;
      if (ds >= 30) return RFC_E_BAD_SYMBOL;
      eb = getbits(br, kDistExtra[ds]);
      if (eb < 0) return RFC_E_TRUNCATED;
      size_t dist = (size_t)kDistBase[ds] + (size_t)eb;
      if (dist > *outpos) return RFC_E_DIST_TOO_FAR;
      if (*outpos + (size_t)len > cap) return RFC_E_OUTPUT_FULL;
      for (int i = 0; i < len; ++i) {
        out[*outpos] = out[*outpos - dist];
        (*outpos)++;
      }
    }
  }
}

/*
 * Inflate one complete raw DEFLATE stream.  Returns bytes produced (>=0) or a negative
 * RFC_E_* code.  *consumed (optional) receives the number of input bytes used, *blocks the
 * number of blocks seen, *err_bit the bit position where decoding stopped on error.
 */
ORACLE_API long rfc1951_inflate(const uint8_t* in, size_t in_len, uint8_t* out, size_t cap,
                                size_t* consumed, int* blocks, size_t* err_bit) {
  bitreader br = {in, in_len, 0};
  size_t outpos = 0;
  int nblocks = 0, last, rc = 0;
  do {
    last = getbit(&br);
    long type = getbits(&br, 2);
    if (last < 0 || type < 0) { rc = RFC_E_TRUNCATED; break; }
    nblocks++;
    if (type == 0) {
      br.bitpos = (br.bitpos + 7) & ~(size_t)7;
      long len = getbits(&br, 16), nlen = getbits(&br, 16);
      if (len < 0 || nlen < 0) { rc = RFC_E_TRUNCATED; break; }
      if ((len ^ 0xffff) != nlen) { rc = RFC_E_STORED_LEN; break; }
      size_t at = br.bitpos >> 3;
      if (at + (size_t)len > in_len) { rc = RFC_E_TRUNCATED; break; }
      if (outpos + (size_t)len > cap) { rc = RFC_E_OUTPUT_FULL; break; }
      memcpy(out + outpos, in + at, (size_t)len);
      outpos += (size_t)len;
      br.bitpos += (size_t)len * 8;
    } else if (type == 1) {
      uint8_t lengths[288];
      hufftree lt, dt;
      int i = 0;
      for (; i < 144; ++i) lengths[i] = 8;
      for (; i < 256; ++i) lengths[i] = 9;
      for (; i < 280; ++i) lengths[i] = 7;
      for (; i < 288; ++i) lengths[i] = 8;
      build_tree(&lt, lengths, 288);
      for (i = 0; i < 30; ++i) lengths[i] = 5;
      build_tree(&dt, lengths, 30);
      rc = inflate_codes(&br, out, cap, &outpos, &lt, &dt);
      if (rc) break;
    } else if (type == 2) {
      long hlit = getbits(&br, 5), hdist = getbits(&br, 5), hclen = getbits(&br, 4);
      if (hlit < 0 || hdist < 0 || hclen < 0) { rc = RFC_E_TRUNCATED; break; }
      int nlen = (int)hlit + 257, ndist = (int)hdist + 1, ncode = (int)hclen + 4;
      if (nlen > 286 || ndist > 30) { rc = RFC_E_BAD_LENGTHS; break; }
      uint8_t lengths[320];
      hufftree ct, lt, dt;
      memset(lengths, 0, sizeof lengths);
      int i;
      for (i = 0; i < ncode; ++i) {
        long v = getbits(&br, 3);
        if (v < 0) { rc = RFC_E_TRUNCATED; break; }
        lengths[kClOrder[i]] = (uint8_t)v;
      }
      if (rc) break;
      int left = build_tree(&ct, lengths, 19);
      if (left < 0) { rc = RFC_E_OVERSUBSCRIBED; break; }
      if (left > 0) { rc = RFC_E_INCOMPLETE; break; } /* zlib: "invalid code lengths set" */
      uint8_t ll[320];
      memset(ll, 0, sizeof ll);
      int idx = 0;
      while (idx < nlen + ndist) {
        int sym = decode_sym(&br, &ct);
        if (sym < 0) { rc = sym; break; }
        if (sym < 16) {
          ll[idx++] = (uint8_t)sym;
        } else {
          int prev = 0, rep;
          long eb;
          if (sym == 16) {
            if (idx == 0) { rc = RFC_E_BAD_LENGTHS; break; }
            prev = ll[idx - 1];
            eb = getbits(&br, 2);
            rep = 3 + (int)eb;
          } else if (sym == 17) {
            eb = getbits(&br, 3);
            rep = 3 + (int)eb;
          } else {
            eb = getbits(&br, 7);
            rep = 11 + (int)eb;
          }
          if (eb < 0) { rc = RFC_E_TRUNCATED; break; }
          if (idx + rep > nlen + ndist) { rc = RFC_E_BAD_LENGTHS; break; }
          while (rep--) ll[idx++] = (uint8_t)prev;
        }
      }
      if (rc) break;
      if (ll[256] == 0) { rc = RFC_E_NO_EOB; break; }
      left = build_tree(&lt, ll, nlen);
      if (left < 0) { rc = RFC_E_OVERSUBSCRIBED; break; }
      /* zlib accepts an incomplete litlen/dist set only if it has a single code */
      if (left > 0 && nlen - lt.count[0] != 1) { rc = RFC_E_INCOMPLETE; break; }
      left = build_tree(&dt, ll + nlen, ndist);
      if (left < 0) { rc = RFC_E_OVERSUBSCRIBED; break; }
      if (left > 0 && ndist - dt.count[0] != 1) { rc = RFC_E_INCOMPLETE; break; }
      rc = inflate_codes(&br, out, cap, &outpos, &lt, &dt);
      if (rc) break;
    } else {
      rc = RFC_E_BTYPE;
      break;
    }
  } while (!last);
  if (consumed) *consumed = (br.bitpos + 7) >> 3;
  if (blocks) *blocks = nblocks;
  if (err_bit) *err_bit = br.bitpos;
  return rc ? rc : (long)outpos;
}

/* CRC-32, bitwise (reflected 0xEDB88320), init/final xor 0xFFFFFFFF. */
ORACLE_API uint32_t rfc_crc32(const uint8_t* p, size_t n) {
  uint32_t c = 0xFFFFFFFFu;
  for (size_t i = 0; i < n; ++i) {
    c ^= p[i];
    for (int k = 0; k < 8; ++k) c = (c >> 1) ^ (0xEDB88320u & (0u - (c & 1u)));
  }
  return c ^ 0xFFFFFFFFu;
}

/* Adler-32, RFC 1950 8.2: s1 = 1 + sum(b), s2 = sum(s1), both mod 65521. */
ORACLE_API uint32_t rfc_adler32(const uint8_t* p, size_t n) {
  uint32_t a = 1, b = 0;
  for (size_t i = 0; i < n; ++i) {
    a = (a + p[i]) % 65521u;
    b = (b + a) % 65521u;
  }
  return (b << 16) | a;
}
