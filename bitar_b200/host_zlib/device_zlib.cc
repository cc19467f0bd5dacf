// device_zlib.cc -- CompressDevice<Class_ZLIB> (bitar/device_zlib.h): the reference's chunking contract
// (/root/reference/src/device.cc:156-318, src/memory.cc:350-505) over zlib on the calling thread.
// Built into libbitar_host_zlib.so; links zlib and Arrow only.
#include "bitar/device_zlib.h"

#include <arrow/buffer.h>
#include <zlib.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>

namespace bitar {

std::uint32_t ZlibConfiguration::compressed_seg_size() const noexcept {
  const auto s = decompressed_seg_size;
  const auto expanded = static_cast<std::uint32_t>(std::ceil(static_cast<double>(s) * 1.1));   // src/config.cc:59-73
  const std::uint32_t stored = s + 5u * ((s + 65534u) / 65535u) + 8u;                          // what a stored stream needs
  return std::max(expanded, stored);
}

CompressDevice<Class_ZLIB>::~CompressDevice() = default;

void CompressDevice<Class_ZLIB>::Grow(std::size_t slots) {
  auto slab = std::make_unique<std::uint8_t[]>(slots * stride_);
  for (std::size_t i = slots; i-- > 0;) free_.push_back(slab.get() + i * stride_);
  slab_ranges_.emplace_back(slab.get(), slots);
  taken_.resize(taken_.size() + slots, 0);
  slabs_.push_back(std::move(slab));
}

arrow::Status CompressDevice<Class_ZLIB>::Initialize(const ZlibConfiguration& c) {   // src/device.cc:114-154, 352-415
  if (started_) return arrow::Status::Invalid("the device is already initialized");
  if (num_qps_ == 0) return arrow::Status::Invalid("at least one queue pair is required");
  if (c.decompressed_seg_size < 1024u || c.decompressed_seg_size > (1u << 20))
    return arrow::Status::Invalid("decompressed_seg_size must be in [1024, 1048576], got ", c.decompressed_seg_size);
  if (c.window_size < 8 || c.window_size > 15) return arrow::Status::Invalid("window_size must be in [8, 15]");
  if (c.level < 0 || c.level > 9) return arrow::Status::Invalid("level must be in [0, 9]");
  if (c.max_preallocate_memzones == 0) return arrow::Status::Invalid("max_preallocate_memzones must be positive");
  cfg_ = c;
  stride_ = (static_cast<std::size_t>(c.compressed_seg_size()) + 63u) & ~std::size_t{63};
  std::lock_guard<std::mutex> lock(pool_mutex_);
  Grow(c.max_preallocate_memzones);
  started_ = true;
  return arrow::Status::OK();
}

std::uint8_t* CompressDevice<Class_ZLIB>::Take() {   // DeviceMemory::Take, src/memory.cc:160-189: grows with a warning
  std::lock_guard<std::mutex> lock(pool_mutex_);
  if (free_.empty()) {
    std::fprintf(stderr, "bitar(zlib): slot pool exhausted, allocating %u more slots\n", cfg_.max_preallocate_memzones);
    Grow(cfg_.max_preallocate_memzones);
  }
  auto* slot = free_.back();
  free_.pop_back();
  std::size_t base_index = 0;
  for (const auto& [base, n] : slab_ranges_) {
    if (slot >= base && slot < base + n * stride_) {
      taken_[base_index + static_cast<std::size_t>(slot - base) / stride_] = 1;
      break;
    }
    base_index += n;
  }
  return slot;
}

bool CompressDevice<Class_ZLIB>::Put(const std::uint8_t* slot) {   // DeviceMemory::Put, src/memory.cc:191-209: 1 / 0
  std::lock_guard<std::mutex> lock(pool_mutex_);
  std::size_t base_index = 0;
  for (const auto& [base, n] : slab_ranges_) {
    if (slot >= base && slot < base + n * stride_ && static_cast<std::size_t>(slot - base) % stride_ == 0) {
      auto& flag = taken_[base_index + static_cast<std::size_t>(slot - base) / stride_];
      if (!flag) return false;
      flag = 0;
      free_.push_back(const_cast<std::uint8_t*>(slot));
      return true;
    }
    base_index += n;
  }
  return false;
}

std::size_t CompressDevice<Class_ZLIB>::slots_free() const {
  std::lock_guard<std::mutex> lock(pool_mutex_);
  return free_.size();
}
std::size_t CompressDevice<Class_ZLIB>::slots_total() const {
  std::lock_guard<std::mutex> lock(pool_mutex_);
  return taken_.size();
}

arrow::Result<BufferVector> CompressDevice<Class_ZLIB>::Compress(std::uint16_t queue_pair_id,
                                                                 const std::shared_ptr<arrow::Buffer>& decompressed_buffer) {
  if (!started_) return arrow::Status::Invalid("the device has not been initialized");
  if (queue_pair_id >= num_qps_) return arrow::Status::Invalid("queue pair id ", queue_pair_id, " is out of range");
  BufferVector out;
  if (decompressed_buffer == nullptr || decompressed_buffer->size() == 0) return out;   // src/device.cc:160-163
  const auto seg = static_cast<std::int64_t>(cfg_.decompressed_seg_size);
  const auto total = decompressed_buffer->size();
  z_stream zs{};
  if (deflateInit2(&zs, cfg_.level, Z_DEFLATED, -static_cast<int>(std::max<std::uint8_t>(cfg_.window_size, 9)), 8,
                   cfg_.fixed_huffman ? Z_FIXED : Z_DEFAULT_STRATEGY) != Z_OK)
    return arrow::Status::IOError("deflateInit2 failed");
  arrow::Status st = arrow::Status::OK();
  for (std::int64_t off = 0; off < total; off += seg) {   // one stateless op with flush FINAL per segment, src/memory.cc:106-116
    const auto n = std::min(seg, total - off);
    auto* slot = Take();
    deflateReset(&zs);
    zs.next_in = const_cast<Bytef*>(decompressed_buffer->data() + off);
    zs.avail_in = static_cast<uInt>(n);
    zs.next_out = slot;
    zs.avail_out = static_cast<uInt>(cfg_.compressed_seg_size());
    const int rc = deflate(&zs, Z_FINISH);
    if (rc != Z_STREAM_END) {
      Put(slot);
      st = arrow::Status::IOError("compress op ", off / seg, " failed (", rc == Z_OK || rc == Z_BUF_ERROR ? "out of space" : "error", ")");
      break;
    }
    out.push_back(std::make_unique<arrow::Buffer>(slot, static_cast<std::int64_t>(cfg_.compressed_seg_size() - zs.avail_out)));
  }
  deflateEnd(&zs);
  if (!st.ok()) {
    Recycle(out);
    return st;
  }
  return out;
}

arrow::Status CompressDevice<Class_ZLIB>::Decompress(std::uint16_t queue_pair_id, const BufferVector& compressed_buffers,
                                                     const std::unique_ptr<arrow::ResizableBuffer>& decompressed_buffer) {
  if (!started_) return arrow::Status::Invalid("the device has not been initialized");
  if (queue_pair_id >= num_qps_) return arrow::Status::Invalid("queue pair id ", queue_pair_id, " is out of range");
  if (decompressed_buffer == nullptr) return arrow::Status::Invalid("null output buffer");
  const auto seg = static_cast<std::int64_t>(cfg_.decompressed_seg_size);
  const auto need = seg * static_cast<std::int64_t>(compressed_buffers.size());
  if (decompressed_buffer->capacity() < need)   // src/device.cc:252-262
    return arrow::Status::CapacityError("the output buffer holds ", decompressed_buffer->capacity(), " bytes, ", need, " are needed");
  z_stream zs{};
  if (inflateInit2(&zs, -15) != Z_OK) return arrow::Status::IOError("inflateInit2 failed");
  std::int64_t total = 0;
  arrow::Status st = arrow::Status::OK();
  for (std::size_t i = 0; i < compressed_buffers.size(); ++i) {   // segment i goes to out + i * seg, src/memory.cc:482-493
    const auto& b = compressed_buffers[i];
    inflateReset(&zs);
    zs.next_in = const_cast<Bytef*>(b->data());
    zs.avail_in = static_cast<uInt>(b->size());
    zs.next_out = decompressed_buffer->mutable_data() + static_cast<std::int64_t>(i) * seg;
    zs.avail_out = static_cast<uInt>(seg);
    const int rc = inflate(&zs, Z_FINISH);
    if (rc != Z_STREAM_END) {   // (bytes after the final block -- the CUDA class's index trailer -- stay unread)
      st = arrow::Status::IOError("decompress op ", i, " failed (", rc == Z_BUF_ERROR || rc == Z_OK ? "out of space" : "data error", ")");
      break;
    }
    const auto produced = seg - static_cast<std::int64_t>(zs.avail_out);
    if (i + 1 < compressed_buffers.size() && produced != seg) {
      st = arrow::Status::IOError("decompress op ", i, " produced ", produced, " bytes, a full segment has ", seg);
      break;
    }
    total += produced;
  }
  inflateEnd(&zs);
  ARROW_RETURN_NOT_OK(st);
  return decompressed_buffer->Resize(total, /*shrink_to_fit=*/false);
}

std::size_t CompressDevice<Class_ZLIB>::Recycle(const BufferVector& buffers) {   // src/device.cc:320-327
  std::size_t n = 0;
  for (const auto& b : buffers)
    if (b != nullptr && Put(b->data())) ++n;
  return n;
}

ZlibDeviceManager* ZlibDeviceManager::Instance() {
  static ZlibDeviceManager instance;
  return &instance;
}

template <>
arrow::Result<ZlibCompressDevice*> ZlibDeviceManager::Create<Class_ZLIB>(std::uint8_t device_id, std::uint16_t num_qps) {
  return new ZlibCompressDevice(device_id, num_qps);
}

}  // namespace bitar
