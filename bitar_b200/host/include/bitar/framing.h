// framing.h -- RFC 1950 (zlib) / RFC 1952 (gzip) framing of compressed chunks, and the other direction.
//
// A chunk CompressDevice::Compress() returns is a raw DEFLATE stream (RFC 1951), followed by the parallel-inflate
// index unless the device was configured with emit_index = false (bitar_b200/csrc/deflate_common.h).  Arrow's and
// Parquet's GZIP / ZLIB codecs and every zlib tool expect one of the two wrappers around the raw stream; the reference
// points at that interop in its app (/root/reference/apps/demo_app.cc:148-150).  These helpers add the wrapper from the
// chunk and the checksum the kernels computed (Configuration: checksum_type), and locate the raw stream inside a
// wrapped buffer so that Decompress() can be pointed at it.  Header only, no CUDA: the same functions as
// bitar_b200/engine.py (stream_length, zlib_streams, gzip_members, unframe).
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>

namespace bitar::framing {

inline constexpr std::uint32_t kIndexMagic = 0xB17A0B02u;   // deflate_common.h

inline std::uint32_t LoadLe32(const std::uint8_t* p) {
  return (std::uint32_t)p[0] | ((std::uint32_t)p[1] << 8) | ((std::uint32_t)p[2] << 16) | ((std::uint32_t)p[3] << 24);
}

/// \brief Bytes of the raw DEFLATE stream at the start of a compressed chunk: the chunk minus the parallel-inflate
/// index, when it carries one.
inline std::size_t StreamLength(const std::uint8_t* chunk, std::size_t size) {
  if (size >= 16 && LoadLe32(chunk + size - 4) == kIndexMagic) {
    const std::uint64_t total = LoadLe32(chunk + size - 8), end_bit = LoadLe32(chunk + size - 12);
    const std::uint64_t full = total >> 16, rem = total & 65535u;
    const std::uint64_t entries = full * 33u + (rem ? 1u + (rem + 2047u) / 2048u : 0u);
    if ((end_bit + 7u) / 8u + 4u * (entries + 3u) == size) return (std::size_t)((end_bit + 7u) / 8u);
  }
  return size;
}

/// \brief One RFC 1950 stream: CMF/FLG 78 01 (32 KiB window, fastest level, no dictionary), the raw stream, Adler-32
/// big-endian.  \p adler32 is the high word of the chunk's bitar_result::checksum (checksum_type ADLER32).
inline std::string ZlibStream(const std::uint8_t* chunk, std::size_t size, std::uint32_t adler32) {
  const std::size_t n = StreamLength(chunk, size);
  std::string out;
  out.reserve(n + 6);
  out.append("\x78\x01", 2);
  out.append(reinterpret_cast<const char*>(chunk), n);
  const char tail[4] = {(char)(adler32 >> 24), (char)(adler32 >> 16), (char)(adler32 >> 8), (char)adler32};
  out.append(tail, 4);
  return out;
}

/// \brief One RFC 1952 member: the 10-byte header, the raw stream, CRC-32 and ISIZE little-endian.  Members of
/// consecutive chunks concatenate to a valid multi-member gzip file of the whole buffer.
inline std::string GzipMember(const std::uint8_t* chunk, std::size_t size, std::uint32_t crc32, std::uint32_t isize) {
  const std::size_t n = StreamLength(chunk, size);
  std::string out;
  out.reserve(n + 18);
  out.append("\x1f\x8b\x08\x00\x00\x00\x00\x00\x00\xff", 10);
  out.append(reinterpret_cast<const char*>(chunk), n);
  for (std::uint32_t v : {crc32, isize}) {
    const char le[4] = {(char)v, (char)(v >> 8), (char)(v >> 16), (char)(v >> 24)};
    out.append(le, 4);
  }
  return out;
}

struct Unframed {
  std::size_t offset, length;   // the raw DEFLATE stream (its own end is found by the decoder; `length` spans up to the trailer)
  bool gzip;                    // else zlib
  std::uint32_t checksum;       // the trailer's CRC-32 (gzip) or Adler-32 (zlib), to compare with the kernel's
  std::uint32_t isize;          // gzip: uncompressed size mod 2^32
};

/// \brief Locate the raw DEFLATE stream inside ONE zlib stream or ONE gzip member.  Throws std::invalid_argument for
/// anything else (preset dictionaries, unknown methods, truncated framing).
inline Unframed Unframe(const std::uint8_t* b, std::size_t n) {
  if (n >= 18 && b[0] == 0x1F && b[1] == 0x8B) {
    if (b[2] != 8) throw std::invalid_argument("gzip member: method is not DEFLATE");
    const unsigned flg = b[3];
    std::size_t at = 10;
    if (flg & 0xE0u) throw std::invalid_argument("gzip member: reserved flag bits set");
    if (flg & 4u) {   // FEXTRA
      if (at + 2 > n) throw std::invalid_argument("gzip member: truncated header");
      at += 2u + b[at] + 256u * b[at + 1];
    }
    for (unsigned bit : {8u, 16u})   // FNAME, FCOMMENT: zero-terminated
      if (flg & bit) {
        while (at < n && b[at] != 0) ++at;
        ++at;
      }
    if (flg & 2u) at += 2;   // FHCRC
    if (at + 8 > n) throw std::invalid_argument("gzip member: truncated");
    return {at, n - 8 - at, true, LoadLe32(b + n - 8), LoadLe32(b + n - 4)};
  }
  if (n >= 6 && (b[0] & 0x0Fu) == 8 && (((unsigned)b[0] << 8) | b[1]) % 31u == 0) {
    if ((b[0] >> 4) > 7) throw std::invalid_argument("zlib stream: window larger than 32 KiB");
    if (b[1] & 0x20u) throw std::invalid_argument("zlib stream: preset dictionary");
    const std::uint32_t adler = ((std::uint32_t)b[n - 4] << 24) | ((std::uint32_t)b[n - 3] << 16) | ((std::uint32_t)b[n - 2] << 8) | b[n - 1];
    return {2, n - 6, false, adler, 0};
  }
  throw std::invalid_argument("neither a zlib stream nor a gzip member");
}

}  // namespace bitar::framing
