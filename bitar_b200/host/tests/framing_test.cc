// framing_test -- bitar/framing.h against zlib itself (CPU only; run by tests/test_host_cpp.py).
#include <zlib.h>

#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "bitar/framing.h"

static std::vector<std::uint8_t> RawDeflate(const std::vector<std::uint8_t>& d) {
  z_stream zs{};
  deflateInit2(&zs, 1, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY);
  std::vector<std::uint8_t> out(d.size() + d.size() / 8 + 256);
  zs.next_in = const_cast<Bytef*>(d.data());
  zs.avail_in = (uInt)d.size();
  zs.next_out = out.data();
  zs.avail_out = (uInt)out.size();
  deflate(&zs, Z_FINISH);
  out.resize(out.size() - zs.avail_out);
  deflateEnd(&zs);
  return out;
}
static bool Inflates(const std::string& framed, int wbits, const std::vector<std::uint8_t>& want) {
  z_stream zs{};
  inflateInit2(&zs, wbits);
  std::vector<std::uint8_t> out(want.size() + 64);
  zs.next_in = reinterpret_cast<Bytef*>(const_cast<char*>(framed.data()));
  zs.avail_in = (uInt)framed.size();
  zs.next_out = out.data();
  zs.avail_out = (uInt)out.size();
  const int rc = inflate(&zs, Z_FINISH);
  const std::size_t got = out.size() - zs.avail_out;
  const bool all = zs.avail_in == 0;
  inflateEnd(&zs);
  return rc == Z_STREAM_END && all && got == want.size() && std::equal(want.begin(), want.end(), out.begin());
}
#define EXPECT(c) do { if (!(c)) { std::fprintf(stderr, "FAILED: %s (line %d)\n", #c, __LINE__); return 1; } } while (0)

int main() {
  namespace fr = bitar::framing;
  std::vector<std::uint8_t> data(100000);
  for (std::size_t i = 0; i < data.size(); ++i) data[i] = (std::uint8_t)((i * 7 / 13) ^ (i >> 9));
  auto raw = RawDeflate(data);
  // a chunk with a (structurally valid) index trailer appended: 100000 bytes = 1 full block + 34464 bytes (17 sub-ranges)
  std::vector<std::uint8_t> chunk = raw;
  const std::uint32_t entries = 33 + 1 + 17, end_bit = (std::uint32_t)raw.size() * 8;
  auto put = [&](std::uint32_t v) { for (int k = 0; k < 4; ++k) chunk.push_back((std::uint8_t)(v >> (8 * k))); };
  for (std::uint32_t i = 0; i < entries; ++i) put(0);
  put(end_bit);
  put((std::uint32_t)data.size());
  put(fr::kIndexMagic);
  EXPECT(fr::StreamLength(chunk.data(), chunk.size()) == raw.size());
  EXPECT(fr::StreamLength(raw.data(), raw.size()) == raw.size());
  const std::uint32_t adler = (std::uint32_t)adler32(adler32(0, nullptr, 0), data.data(), (uInt)data.size());
  const std::uint32_t crc = (std::uint32_t)crc32(crc32(0, nullptr, 0), data.data(), (uInt)data.size());
  const std::string z = fr::ZlibStream(chunk.data(), chunk.size(), adler), g = fr::GzipMember(chunk.data(), chunk.size(), crc, (std::uint32_t)data.size());
  EXPECT(Inflates(z, 15, data));
  EXPECT(Inflates(g, 15 + 16, data));
  const auto uz = fr::Unframe(reinterpret_cast<const std::uint8_t*>(z.data()), z.size());
  EXPECT(!uz.gzip && uz.offset == 2 && uz.length == raw.size() && uz.checksum == adler);
  const auto ug = fr::Unframe(reinterpret_cast<const std::uint8_t*>(g.data()), g.size());
  EXPECT(ug.gzip && ug.offset == 10 && ug.length == raw.size() && ug.checksum == crc && ug.isize == data.size());
  // zlib's own wrappers come apart the same way
  {
    std::vector<std::uint8_t> zz(compressBound((uLong)data.size()));
    uLongf zn = (uLongf)zz.size();
    compress2(zz.data(), &zn, data.data(), (uLong)data.size(), 6);
    const auto u = fr::Unframe(zz.data(), zn);
    EXPECT(!u.gzip && u.checksum == adler && u.offset == 2 && u.length == zn - 6);
  }
  bool threw = false;
  try { fr::Unframe(raw.data(), 5); } catch (const std::invalid_argument&) { threw = true; }
  EXPECT(threw);
  std::printf("framing: OK\n");
  return 0;
}
