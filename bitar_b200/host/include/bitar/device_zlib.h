// device_zlib.h -- CompressDevice<Class_ZLIB>: a second device class behind DeviceManager::Create, the host-zlib
// software path as a first-class backend (what DPDK's compress_zlib PMD is to the reference: SURVEY.md 8(f) rank 4;
// the factory / class dispatch it exercises is /root/reference/src/include/device.h:191-219, src/driver.cc:129-150).
//
// SEPARATE LIBRARY (libbitar_host_zlib.so, bitar_b200/host_zlib/): it links zlib and Arrow, never libbitar_cuda.so, and
// neither libbitar_host.so nor anything on the CUDA path loads it.  Class_CUDA cannot fall back to it: an
// application chooses a class through DeviceManager::Create<Class>, as the reference chooses Class_MLX5_PCI.
//
// Same contract as the CUDA class (device.h): Compress() cuts the buffer into decompressed_seg_size() segments, one
// complete raw DEFLATE stream (zlib level 1, window 32 KiB: src/config.cc:83-91) per segment, returned in input order as
// non-owning views into pool slots that the caller must Recycle(); Decompress() inflates buffer i to offset i * seg of the
// output and resizes it to the total.  Queue pairs are independent (one caller thread each, like an lcore).
#pragma once
#include <arrow/result.h>
#include <arrow/status.h>

#include <cstddef>
#include <cstdint>
#include <memory>
#include <mutex>
#include <type_traits>
#include <vector>

#include "type_fwd.h"

namespace arrow {
class Buffer;
class ResizableBuffer;
}  // namespace arrow

namespace bitar {

namespace internal {
enum class HostDriverClass : std::int8_t { ZLIB };
}
using Class_ZLIB = std::integral_constant<internal::HostDriverClass, internal::HostDriverClass::ZLIB>;

/// \brief The knobs of src/include/config.h:76-121 that apply to a software codec.
struct ZlibConfiguration {
  std::uint32_t decompressed_seg_size = 2048;
  std::uint32_t max_preallocate_memzones = 2560;
  std::uint8_t window_size = 15;      // 8..15 (a stream written with 8 needs 9 to inflate: zlib's own rule)
  bool fixed_huffman = false;         // Z_FIXED instead of dynamic codes
  int level = 1;                      // src/config.cc:87 RTE_COMP_LEVEL_MIN + 1
  /// \brief src/config.cc:59-73, never below the stored-block bound (as include/bitar_cuda.h: bitar_compressed_seg_size).
  [[nodiscard]] std::uint32_t compressed_seg_size() const noexcept;
};

template <typename Class>
class CompressDevice;

template <>
class CompressDevice<Class_ZLIB> {
 public:
  CompressDevice(const CompressDevice&) = delete;
  CompressDevice& operator=(const CompressDevice&) = delete;
  ~CompressDevice();

  arrow::Status Initialize(const ZlibConfiguration& configuration);
  arrow::Result<BufferVector> Compress(std::uint16_t queue_pair_id, const std::shared_ptr<arrow::Buffer>& decompressed_buffer);
  arrow::Status Decompress(std::uint16_t queue_pair_id, const BufferVector& compressed_buffers,
                           const std::unique_ptr<arrow::ResizableBuffer>& decompressed_buffer);
  std::size_t Recycle(const BufferVector& buffers);

  [[nodiscard]] auto device_id() const noexcept { return device_id_; }
  [[nodiscard]] std::uint16_t num_qps() const noexcept { return num_qps_; }
  [[nodiscard]] std::size_t slots_free() const;
  [[nodiscard]] std::size_t slots_total() const;

 private:
  friend class ZlibDeviceManager;
  CompressDevice(std::uint8_t device_id, std::uint16_t num_qps) : device_id_(device_id), num_qps_(num_qps) {}
  std::uint8_t* Take();
  bool Put(const std::uint8_t* slot);
  void Grow(std::size_t slots);

  const std::uint8_t device_id_;
  const std::uint16_t num_qps_;
  ZlibConfiguration cfg_;
  bool started_ = false;
  std::size_t stride_ = 0;
  mutable std::mutex pool_mutex_;
  std::vector<std::unique_ptr<std::uint8_t[]>> slabs_;
  std::vector<std::pair<const std::uint8_t*, std::size_t>> slab_ranges_;   // base, slots
  std::vector<std::uint8_t*> free_;
  std::vector<std::uint8_t> taken_;   // one flag per slot of every slab, in slab order
};

using ZlibCompressDevice = CompressDevice<Class_ZLIB>;

/// \brief The DeviceManager::Create<..., Class> analogue for host classes (src/include/device.h:196-219): any class
/// without a specialisation is NotImplemented, Class_ZLIB yields a device (device ids are labels: the class has no
/// hardware to probe).
class ZlibDeviceManager {
 public:
  static ZlibDeviceManager* Instance();
  template <typename Class>
  arrow::Result<CompressDevice<Class>*> Create(std::uint8_t device_id, std::uint16_t /*num_qps*/) {
    return arrow::Status::NotImplemented("Unsupported compress device ", +device_id);
  }
};
template <>
arrow::Result<ZlibCompressDevice*> ZlibDeviceManager::Create<Class_ZLIB>(std::uint8_t device_id, std::uint16_t num_qps);

}  // namespace bitar
