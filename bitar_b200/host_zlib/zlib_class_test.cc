// zlib_class_test -- CompressDevice<Class_ZLIB> behind ZlibDeviceManager::Create (CPU only; run by tests/test_host_zlib_class.py).
#include <arrow/buffer.h>
#include <zlib.h>

#include <cstdio>
#include <cstring>
#include <vector>

#include "bitar/device_zlib.h"

#define EXPECT(c) do { if (!(c)) { std::fprintf(stderr, "FAILED: %s (line %d)\n", #c, __LINE__); return 1; } } while (0)

namespace bitar { namespace internal { enum class Nothing : int { kNone }; } }
using Class_NONE = std::integral_constant<bitar::internal::Nothing, bitar::internal::Nothing::kNone>;
namespace bitar { template <> class CompressDevice<Class_NONE> {}; }

int main() {
  auto* manager = bitar::ZlibDeviceManager::Instance();
  EXPECT(manager->Create<Class_NONE>(0, 1).status().IsNotImplemented());   // a class without a specialisation: src/include/device.h:203-207
  auto made = manager->Create<bitar::Class_ZLIB>(0, 2);
  EXPECT(made.ok());
  std::unique_ptr<bitar::ZlibCompressDevice> dev(*made);
  bitar::ZlibConfiguration cfg;
  cfg.decompressed_seg_size = 59460;
  cfg.max_preallocate_memzones = 8;
  {
    bitar::ZlibConfiguration bad = cfg;
    bad.decompressed_seg_size = 100;
    EXPECT(dev->Initialize(bad).IsInvalid());
  }
  EXPECT(dev->Compress(0, nullptr).status().IsInvalid());                  // not initialized yet
  EXPECT(dev->Initialize(cfg).ok());
  EXPECT(dev->Initialize(cfg).IsInvalid());
  EXPECT(dev->slots_total() == 8 && dev->slots_free() == 8);

  // columnar-looking input: 20 segments and a ragged tail (the pool of 8 slots has to grow)
  const std::size_t n = 20 * 59460 + 1234;
  std::vector<std::uint8_t> data(n);
  std::uint64_t x = 88172645463325252ull, key = 1000;
  for (std::size_t i = 0; i + 8 <= n; i += 8) {
    x ^= x << 13; x ^= x >> 7; x ^= x << 17;
    key += x % 7;
    std::memcpy(&data[i], &key, 8);
  }
  auto in = std::make_shared<arrow::Buffer>(data.data(), static_cast<std::int64_t>(n));
  auto comp = dev->Compress(1, in);
  EXPECT(comp.ok() && comp->size() == 21);
  EXPECT(dev->slots_total() == 24 && dev->slots_free() == 3);
  EXPECT(dev->Compress(2, in).status().IsInvalid());                       // queue pair out of range
  EXPECT(dev->Compress(0, std::make_shared<arrow::Buffer>(nullptr, 0))->empty());
  // every buffer is one complete raw DEFLATE stream that zlib inflates to its segment
  for (std::size_t i = 0; i < comp->size(); ++i) {
    std::vector<std::uint8_t> seg(59460);
    z_stream zs{};
    EXPECT(inflateInit2(&zs, -15) == Z_OK);
    zs.next_in = const_cast<Bytef*>((*comp)[i]->data());
    zs.avail_in = static_cast<uInt>((*comp)[i]->size());
    zs.next_out = seg.data();
    zs.avail_out = 59460;
    EXPECT(inflate(&zs, Z_FINISH) == Z_STREAM_END && zs.avail_in == 0);
    const std::size_t want = i + 1 < comp->size() ? 59460 : 1234;
    EXPECT(59460 - zs.avail_out == want && std::memcmp(seg.data(), data.data() + i * 59460, want) == 0);
    inflateEnd(&zs);
  }
  auto out = arrow::AllocateResizableBuffer(21 * 59460);
  EXPECT(out.ok());
  EXPECT(dev->Decompress(0, *comp, *out).ok());
  EXPECT((*out)->size() == static_cast<std::int64_t>(n) && std::memcmp((*out)->data(), data.data(), n) == 0);
  {
    auto small = arrow::AllocateResizableBuffer(59460);
    EXPECT(dev->Decompress(0, *comp, *small).IsCapacityError());
  }
  {   // a damaged stream is an IOError, as a failed op is in the reference (src/device.cc:512-520)
    std::vector<std::uint8_t> bad((*comp)[0]->data(), (*comp)[0]->data() + (*comp)[0]->size());
    bad[bad.size() / 2] ^= 0x5A;
    bad.resize(bad.size() / 2 + 1);
    bitar::BufferVector one;
    one.push_back(std::make_unique<arrow::Buffer>(bad.data(), static_cast<std::int64_t>(bad.size())));
    auto o2 = arrow::AllocateResizableBuffer(59460);
    EXPECT(dev->Decompress(0, one, *o2).IsIOError());
  }
  EXPECT(dev->Recycle(*comp) == 21 && dev->slots_free() == 24);
  EXPECT(dev->Recycle(*comp) == 0);                                         // Put() of a free slot returns 0 (src/memory.cc:191-209)
  std::printf("zlib class: OK (21 chunks, %zu -> round trip)\n", n);
  return 0;
}
