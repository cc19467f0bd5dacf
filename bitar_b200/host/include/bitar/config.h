// config.h -- Configuration<Class> and its CUDA specialisation.
// Mirrors /root/reference/src/include/config.h:41-183 and src/config.cc:35-105 with a Class_CUDA tag;
// the rte_comp_xform accessors are replaced by the POD the C-ABI takes (bitar_cfg).
#pragma once
#include <cstdint>
#include <string>
#include <string_view>
#include <type_traits>

#include "bitar_cuda.h"

namespace bitar {

namespace internal {

static inline constexpr auto kExpanseRatio = 1.1;                       // config.h:41
static inline constexpr std::uint32_t kMinSegSize = BITAR_MIN_SEG_SIZE;  // config.h:45
// config.h:46-47 kMaxSegSize = (65535 - RTE_PKTMBUF_HEADROOM) / 1.1 = 59460: the reference limit, kept as
// the compat default; the CUDA engine accepts up to kWideMaxSegSize.
static inline constexpr std::uint32_t kMaxSegSize = BITAR_REF_MAX_SEG_SIZE;
static inline constexpr std::uint32_t kWideMaxSegSize = BITAR_MAX_SEG_SIZE;
static inline constexpr std::uint32_t kDefaultSegSize = 2048;           // config.h:48
static inline constexpr std::uint32_t kMaxPreallocateSlots = 1u << 22;  // RTE_MAX_MEMZONE analogue

enum class DriverClass : std::int8_t { CUDA };

}  // namespace internal

using Class_CUDA = std::integral_constant<internal::DriverClass, internal::DriverClass::CUDA>;

enum class HuffmanType : std::uint8_t { kDefault = BITAR_HUFFMAN_DEFAULT, kFixed = BITAR_HUFFMAN_FIXED, kDynamic = BITAR_HUFFMAN_DYNAMIC };
enum class ChecksumType : std::uint8_t {
  kNone = BITAR_CHECKSUM_NONE, kCrc32 = BITAR_CHECKSUM_CRC32, kAdler32 = BITAR_CHECKSUM_ADLER32, kCrc32Adler32 = BITAR_CHECKSUM_CRC32_ADLER32
};
enum class SlotMemory : std::uint8_t { kDevice = BITAR_MEM_DEVICE, kPinnedHost = BITAR_MEM_PINNED };

template <typename Class>
class Configuration {
 public:
  Configuration() { UpdateCompressedSegSize(); }
  virtual ~Configuration() = default;

  [[nodiscard]] virtual std::string ToString() const;
  [[nodiscard]] virtual std::string_view type_name() const noexcept = 0;
  /// \brief The POD handed to bitar_dev_open (replaces compress_xform()/decompress_xform()).
  [[nodiscard]] virtual bitar_cfg to_c() const noexcept;

  [[nodiscard]] auto burst_size() const noexcept { return burst_size_; }
  void set_burst_size(std::uint16_t v) { burst_size_ = v; }
  [[nodiscard]] auto max_sgl_segs() const noexcept { return max_sgl_segs_; }
  void set_max_sgl_segs(std::uint16_t v) { max_sgl_segs_ = v; }
  [[nodiscard]] auto decompressed_seg_size() const noexcept { return decompressed_seg_size_; }
  void set_decompressed_seg_size(std::uint32_t v) {
    decompressed_seg_size_ = v;
    UpdateCompressedSegSize();
  }
  [[nodiscard]] auto compressed_seg_size() const noexcept { return compressed_seg_size_; }
  [[nodiscard]] auto window_size() const noexcept { return window_size_; }
  void set_window_size(std::uint8_t v) { window_size_ = v; }
  [[nodiscard]] auto huffman_enc() const noexcept { return huffman_enc_; }
  void set_huffman_enc(HuffmanType v) { huffman_enc_ = v; }
  [[nodiscard]] auto max_preallocate_memzones() const noexcept { return max_preallocate_memzones_; }
  void set_max_preallocate_memzones(std::uint32_t v) { max_preallocate_memzones_ = v; }

 private:
  void UpdateCompressedSegSize() noexcept { compressed_seg_size_ = bitar_compressed_seg_size(decompressed_seg_size_); }

  std::uint16_t burst_size_ = 32;
  std::uint16_t max_sgl_segs_ = 1U;
  std::uint32_t decompressed_seg_size_ = internal::kDefaultSegSize;
  std::uint32_t compressed_seg_size_{};
  std::uint8_t window_size_ = 0U;
  HuffmanType huffman_enc_ = HuffmanType::kDynamic;
  std::uint32_t max_preallocate_memzones_ = 2560;  // RTE_MAX_MEMZONE
};

static inline constexpr std::string_view kCudaConfigurationTypeName{"cuda"};

/// \brief Specific configuration for CUDA devices (the BlueFieldConfiguration analogue, config.h:157-183).
class CudaConfiguration : public Configuration<Class_CUDA> {
 public:
  [[nodiscard]] std::string_view type_name() const noexcept override { return kCudaConfigurationTypeName; }
  [[nodiscard]] std::string ToString() const override;
  [[nodiscard]] bitar_cfg to_c() const noexcept override;

  [[nodiscard]] auto checksum_type() const noexcept { return checksum_type_; }
  void set_checksum_type(ChecksumType v) { checksum_type_ = v; }
  /// \brief Where the output slots live: device memory, or pinned host memory the kernels write over PCIe.
  [[nodiscard]] auto slot_memory() const noexcept { return slot_memory_; }
  void set_slot_memory(SlotMemory v) { slot_memory_ = v; }
  /// \brief Whether a compressed chunk carries the parallel-inflate index after its final block (default).  Without it a
  /// chunk is a bare RFC 1951 stream whose length is the buffer's size; it still inflates here, through the slower
  /// whole-stream kernel.
  [[nodiscard]] auto emit_index() const noexcept { return emit_index_; }
  void set_emit_index(bool v) { emit_index_ = v; }

  static CudaConfiguration Defaults() { return {}; }

 private:
  ChecksumType checksum_type_ = ChecksumType::kNone;
  SlotMemory slot_memory_ = SlotMemory::kDevice;
  bool emit_index_ = true;
};

}  // namespace bitar
