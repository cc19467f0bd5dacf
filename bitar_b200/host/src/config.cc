// config.cc -- Configuration<Class_CUDA> / CudaConfiguration.
// Mirrors /root/reference/src/config.cc:35-105: ToString() and the translation of the knobs into what the
// device consumes (there: rte_comp_xform; here: the bitar_cfg POD of the C-ABI).
#include "bitar/config.h"

#include <sstream>

namespace bitar {

template <typename Class>
std::string Configuration<Class>::ToString() const {
  std::ostringstream os;
  os << "burst_size: " << burst_size_ << ", max_sgl_segs: " << max_sgl_segs_
     << ", decompressed_seg_size: " << decompressed_seg_size_ << ", compressed_seg_size: " << compressed_seg_size_
     << ", window_size: " << static_cast<unsigned>(window_size_) << ", huffman_enc: "
     << (huffman_enc_ == HuffmanType::kFixed ? "FIXED" : huffman_enc_ == HuffmanType::kDynamic ? "DYNAMIC" : "DEFAULT")
     << ", max_preallocate_memzones: " << max_preallocate_memzones_;
  return os.str();
}

template <typename Class>
bitar_cfg Configuration<Class>::to_c() const noexcept {
  bitar_cfg c{};
  c.decompressed_seg_size = decompressed_seg_size_;
  c.compressed_seg_size = compressed_seg_size_;
  c.max_preallocate_slots = max_preallocate_memzones_;
  c.burst_size = burst_size_;
  c.max_sgl_segs = max_sgl_segs_;
  c.window_size = window_size_;
  c.huffman_enc = static_cast<std::uint8_t>(huffman_enc_);
  c.checksum_type = BITAR_CHECKSUM_NONE;
  c.slot_mem_kind = BITAR_MEM_DEVICE;
  return c;
}

std::string CudaConfiguration::ToString() const {
  static const char* const kChecksum[] = {"NONE", "CRC32", "ADLER32", "CRC32_ADLER32"};
  return Configuration<Class_CUDA>::ToString() + ", checksum_type: " + kChecksum[static_cast<unsigned>(checksum_type_) & 3u] +
         ", slot_memory: " + (slot_memory_ == SlotMemory::kDevice ? "DEVICE" : "PINNED_HOST");
}

bitar_cfg CudaConfiguration::to_c() const noexcept {
  bitar_cfg c = Configuration<Class_CUDA>::to_c();
  c.checksum_type = static_cast<std::uint8_t>(checksum_type_);
  c.slot_mem_kind = static_cast<std::uint8_t>(slot_memory_);
  c.no_index = emit_index_ ? 0 : 1;
  return c;
}

template class Configuration<Class_CUDA>;

}  // namespace bitar
