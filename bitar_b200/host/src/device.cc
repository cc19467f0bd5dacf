// device.cc -- CompressDevice<Class_CUDA> over the C-ABI (include/bitar_cuda.h).
// Mirrors /root/reference/src/device.cc:114-588: same public methods, state machine, argument meaning
// and arrow::Status codes; the enqueue/dequeue burst loop (device.cc:204-235) collapses into one enqueue
// of the whole op list on the queue pair's CUDA stream plus one wait.
#include "bitar/device.h"

#include <arrow/buffer.h>
#include <arrow/util/logging.h>

#include <algorithm>
#include <cstring>
#include <string>

namespace bitar {

namespace internal {

arrow::Status StatusFromC(int rc) {
  if (rc == BITAR_OK) return arrow::Status::OK();
  const std::string msg = bitar_last_error();
  switch (rc) {   // negative arrow::StatusCode, src/include/util.h:166-205
    case BITAR_E_OUT_OF_MEMORY: return arrow::Status::OutOfMemory(msg);
    case BITAR_E_INVALID: return arrow::Status::Invalid(msg);
    case BITAR_E_IO_ERROR: return arrow::Status::IOError(msg);
    case BITAR_E_CAPACITY: return arrow::Status::CapacityError(msg);
    case BITAR_E_CANCELLED: return arrow::Status::Cancelled(msg);
    case BITAR_E_NOT_IMPLEMENTED: return arrow::Status::NotImplemented(msg);
    default: return arrow::Status::UnknownError(msg);
  }
}

}  // namespace internal

template <typename Class>
CompressDevice<Class>::CompressDevice(std::uint8_t device_id, std::uint16_t num_qps)
    : device_id_{device_id}, num_qps_{num_qps}, qp_state_(num_qps) {}

template <typename Class>
CompressDevice<Class>::~CompressDevice() {   // stop / close, src/device.cc:329-343
  if (handle_ != nullptr) {
    for (std::uint16_t q = 0; q < num_qps_; ++q) {
      bitar_qp_wait(handle_, q);
      Unregister(q);
      if (qp_state_[q].staged != nullptr) bitar_mem_free(BITAR_MEM_PINNED, device_id_, qp_state_[q].staged);
      qp_state_[q].staged = nullptr;
    }
    bitar_dev_close(handle_);
  }
  handle_ = nullptr;
  state_ = internal::DeviceState::kUndefined;
}

template <typename Class>
arrow::Status CompressDevice<Class>::set_configuration(std::unique_ptr<Configuration<Class>> configuration) {
  configuration_ = std::move(configuration);
  return arrow::Status::OK();
}

template <typename Class>
arrow::Status CompressDevice<Class>::ValidateConfiguration() {   // src/device.cc:352-415
  if (configuration_ == nullptr) return arrow::Status::Invalid("Configuration is not set");
  bitar_dev_info info{};
  ARROW_RETURN_NOT_OK(internal::StatusFromC(bitar_cuda_device_info(device_id_, &info)));
  if (configuration_->burst_size() == 0) return arrow::Status::Invalid("burst_size must be greater than 0");
  if (configuration_->max_sgl_segs() > 1 && !info.supports_sgl)
    return arrow::Status::Invalid("Compress device does not support chained mbufs. max_sgl_segs must be <= 1");
  const auto seg = configuration_->decompressed_seg_size();
  if (seg < internal::kMinSegSize || seg > internal::kWideMaxSegSize)
    return arrow::Status::Invalid("decompressed_seg_size is not in the range of [", internal::kMinSegSize, ", ",
                                  internal::kWideMaxSegSize, "]");
  const auto window = configuration_->window_size();
  if (window != 0 && (window < info.window_min || window > info.window_max))
    return arrow::Status::Invalid("window_size is not in the range of [", static_cast<unsigned>(info.window_min), ", ",
                                  static_cast<unsigned>(info.window_max), "]");
  if (configuration_->huffman_enc() == HuffmanType::kFixed && !info.supports_fixed)
    return arrow::Status::Invalid("Compress device does not support fixed Huffman encoding");
  if (configuration_->huffman_enc() == HuffmanType::kDynamic && !info.supports_dynamic)
    return arrow::Status::Invalid("Compress device does not support dynamic Huffman encoding");
  const auto zones = configuration_->max_preallocate_memzones();
  if (zones < BITAR_MIN_PREALLOCATE_SLOTS || zones > internal::kMaxPreallocateSlots)
    return arrow::Status::Invalid("max_preallocate_memzones (", zones, ") is not in the range of [",
                                  BITAR_MIN_PREALLOCATE_SLOTS, ", ", internal::kMaxPreallocateSlots, "]");
  return arrow::Status::OK();
}

template <typename Class>
arrow::Status CompressDevice<Class>::Initialize(std::unique_ptr<Configuration<Class>> configuration) {
  if (state_ == internal::DeviceState::kStarted)
    return arrow::Status::Invalid("Compress device ", static_cast<unsigned>(device_id_), " has already started");
  ARROW_RETURN_NOT_OK(set_configuration(std::move(configuration)));
  ARROW_RETURN_NOT_OK(ValidateConfiguration());
  const bitar_cfg c = configuration_->to_c();
  // configure + queue_pair_setup x N + start + slot preallocation, src/device.cc:121-151,429-441
  ARROW_RETURN_NOT_OK(internal::StatusFromC(bitar_dev_open(device_id_, num_qps_, &c, &handle_)));
  state_ = internal::DeviceState::kStarted;
  return arrow::Status::OK();
}

template <typename Class>
arrow::Status CompressDevice<Class>::EntryGuard(std::uint16_t queue_pair_id) {   // src/device.cc:443-462
  if (state_ != internal::DeviceState::kStarted || handle_ == nullptr)
    return arrow::Status::Invalid("Compress device ", static_cast<unsigned>(device_id_), " has not started");
  if (queue_pair_id >= num_qps_)
    return arrow::Status::Invalid("queue_pair_id must be in the range of [0, ", num_qps_, ")");
  if (bitar_qp_busy(handle_, queue_pair_id) != 0)
    return arrow::Status::Cancelled("Queue pair ", queue_pair_id, " of compress device ",
                                    static_cast<unsigned>(device_id_), " is busy");
  return arrow::Status::OK();
}

template <typename Class>
void CompressDevice<Class>::ReleaseSlots(std::uint16_t queue_pair_id) {   // ReleaseAll, src/device.cc:537-542
  auto& q = qp_state_[queue_pair_id];
  for (auto it = q.slots.rbegin(); it != q.slots.rend(); ++it) bitar_slot_put(handle_, *it);
  q.slots.clear();
}

template <typename Class>
void CompressDevice<Class>::Unregister(std::uint16_t queue_pair_id) {
  auto& q = qp_state_[queue_pair_id];
  if (q.registered != nullptr) bitar_host_unregister(q.registered);
  q.registered = nullptr;
}

namespace {
// Caller memory must be reachable from the device: device memory and pinned / registered host memory are
// used in place (the rte_mem_virt2iova analogue, src/memory.cc:388-391); pageable host memory is
// registered for the duration of the call.
arrow::Status MakeAccessible(const void* ptr, std::size_t size, void** registered) {
  *registered = nullptr;
  if (size == 0 || bitar_ptr_kind(ptr, nullptr) != 0) return arrow::Status::OK();
  ARROW_RETURN_NOT_OK(internal::StatusFromC(bitar_host_register(const_cast<void*>(ptr), size)));
  *registered = const_cast<void*>(ptr);
  return arrow::Status::OK();
}
}  // namespace

template <typename Class>
arrow::Status CompressDevice<Class>::EnqueueCompress(std::uint16_t queue_pair_id,
                                                     const std::shared_ptr<arrow::Buffer>& decompressed_buffer) {
  ARROW_RETURN_NOT_OK(EntryGuard(queue_pair_id));
  Unregister(queue_pair_id);   // left over from an asynchronous call
  auto& q = qp_state_[queue_pair_id];
  q.ops.clear();
  q.results.clear();
  q.slots.clear();
  if (decompressed_buffer == nullptr || decompressed_buffer->size() == 0) return arrow::Status::OK();
  const auto seg = static_cast<std::size_t>(configuration_->decompressed_seg_size());
  const auto size = static_cast<std::size_t>(decompressed_buffer->size());
  const std::size_t n = (size + seg - 1) / seg;   // src/device.cc:168-170
  ARROW_RETURN_NOT_OK(MakeAccessible(decompressed_buffer->data(), size, &q.registered));
  q.slots.resize(n);
  if (bitar_slot_take_n(handle_, static_cast<std::uint32_t>(n), q.slots.data()) != BITAR_OK) {
    q.slots.clear();
    return arrow::Status::IOError("Unable to get enough output slots: ", bitar_last_error());   // src/memory.cc:407-410
  }
  const std::uint32_t slot = bitar_slot_size(handle_);
  // AssembleFrom(span, offset), src/memory.cc:350-430: max_sgl_segs segments chained into one operation = one stream
  // over k * S bytes whose destination is the k slots taken for them (they must be one contiguous range)
  const std::size_t k = std::max<std::size_t>(1, configuration_->max_sgl_segs());
  const std::size_t n_ops = (n + k - 1) / k;
  q.ops.resize(n_ops);
  q.results.assign(n_ops, bitar_result{0, BITAR_OP_NOT_RUN, 0});
  for (std::size_t g = 0; g < n_ops; ++g) {
    const std::size_t cnt = std::min(k, n - g * k);
    for (std::size_t j = 1; j < cnt; ++j)
      if (static_cast<std::uint8_t*>(q.slots[g * k + j]) != static_cast<std::uint8_t*>(q.slots[g * k]) + j * slot) {
        ReleaseSlots(queue_pair_id);
        return arrow::Status::IOError("The output slots of a chained operation are not contiguous");
      }
    q.ops[g].src = decompressed_buffer->data() + g * k * seg;
    q.ops[g].src_len = static_cast<std::uint32_t>(std::min(k * seg, size - g * k * seg));
    q.ops[g].dst = q.slots[g * k];
    q.ops[g].dst_cap = static_cast<std::uint32_t>(cnt * slot);
  }
  const int rc = bitar_qp_deflate(handle_, queue_pair_id, q.ops.data(), static_cast<std::uint32_t>(n_ops), q.results.data());
  if (rc != BITAR_OK) {
    ReleaseSlots(queue_pair_id);
    return internal::StatusFromC(rc);
  }
  return arrow::Status::OK();
}

template <typename Class>
arrow::Result<BufferVector> CompressDevice<Class>::FinishCompress(std::uint16_t queue_pair_id, bool in_callback) {
  auto& q = qp_state_[queue_pair_id];
  BufferVector out;
  if (q.ops.empty()) return out;
  // the dequeue busy-poll, src/device.cc:228-235 (inside a completion callback the work is already done)
  const int rc = in_callback ? bitar_qp_result(handle_, queue_pair_id) : bitar_qp_wait(handle_, queue_pair_id);
  if (!in_callback) Unregister(queue_pair_id);
  if (rc != BITAR_OK) {   // per-op status != SUCCESS -> IOError, src/device.cc:512-520
    ReleaseSlots(queue_pair_id);
    return internal::StatusFromC(rc);
  }
  out.reserve(q.slots.size());
  const std::size_t k = std::max<std::size_t>(1, configuration_->max_sgl_segs());
  if (k == 1) {
    for (std::size_t i = 0; i < q.ops.size(); ++i)   // non-owning views into the slots, src/device.cc:183-195
      out.emplace_back(std::make_unique<arrow::Buffer>(static_cast<const std::uint8_t*>(q.slots[i]),
                                                       static_cast<std::int64_t>(q.results[i].produced)));
  } else {
    // chained segments: one buffer per destination slot that holds data, all of them full but the last, as the
    // reference's dequeue callback lists them (src/device.cc:183-195).  A stream that ends exactly at a slot boundary
    // gets an empty buffer after it, so that the end of a stream is always a buffer shorter than a slot (what
    // Decompress() goes by).  Slots a stream did not reach go back to the pool (the reference leaves them occupied).
    const std::size_t slot = bitar_slot_size(handle_), n = q.slots.size();
    for (std::size_t g = 0; g < q.ops.size(); ++g) {
      const std::size_t p = q.results[g].produced, cnt = std::min(k, n - g * k);
      std::size_t used = (p + slot - 1) / slot;
      const auto* base = static_cast<const std::uint8_t*>(q.slots[g * k]);
      for (std::size_t j = 0; j < used; ++j)
        out.emplace_back(std::make_unique<arrow::Buffer>(base + j * slot, static_cast<std::int64_t>(std::min(slot, p - j * slot))));
      if (p % slot == 0) {
        out.emplace_back(std::make_unique<arrow::Buffer>(base + p, 0));
        if (used < cnt) ++used;   // the empty buffer sits on (and keeps) the next slot
      }
      for (std::size_t j = cnt; j-- > used;) bitar_slot_put(handle_, q.slots[g * k + j]);
    }
  }
  q.slots.clear();
  return out;
}

template <typename Class>
arrow::Result<BufferVector> CompressDevice<Class>::Compress(std::uint16_t queue_pair_id,
                                                            const std::shared_ptr<arrow::Buffer>& decompressed_buffer) {
  if (decompressed_buffer == nullptr || decompressed_buffer->size() == 0) return BufferVector{};   // src/device.cc:161-164
  ARROW_RETURN_NOT_OK(EnqueueCompress(queue_pair_id, decompressed_buffer));
  return FinishCompress(queue_pair_id);
}

template <typename Class>
arrow::Status CompressDevice<Class>::EnqueueDecompress(std::uint16_t queue_pair_id, const BufferVector& compressed_buffers,
                                                       const std::unique_ptr<arrow::ResizableBuffer>& decompressed_buffer) {
  auto& q = qp_state_[queue_pair_id < num_qps_ ? queue_pair_id : 0];
  if (compressed_buffers.empty()) {   // src/device.cc:244-246
    ARROW_RETURN_NOT_OK(EntryGuard(queue_pair_id));
    q.ops.clear();
    q.results.clear();
    return arrow::Status::OK();
  }
  const auto seg = static_cast<std::size_t>(configuration_ ? configuration_->decompressed_seg_size() : 0);
  const std::size_t k = configuration_ ? std::max<std::size_t>(1, configuration_->max_sgl_segs()) : 1;
  if (k > 1) return EnqueueDecompressChained(queue_pair_id, compressed_buffers, decompressed_buffer);
  const std::size_t n = compressed_buffers.size();
  if (decompressed_buffer == nullptr || static_cast<std::size_t>(decompressed_buffer->capacity()) < n * seg)
    return arrow::Status::CapacityError("The decompressed_buffer is required to be >= ", n * seg, " bytes");   // src/device.cc:248-254
  ARROW_RETURN_NOT_OK(EntryGuard(queue_pair_id));
  Unregister(queue_pair_id);
  // Compressed buffers in pageable memory (read from a file into heap buffers, say) are copied into a pinned stage
  // of this queue pair, one buffer per row of a constant stride, which is the layout the library gathers with one
  // strided copy-engine transfer per batch (pool slots look the same).  Device-accessible buffers are used in place.
  std::size_t widest = 0;
  bool pageable = false;
  for (const auto& b : compressed_buffers) {
    if (b == nullptr) return arrow::Status::Invalid("null compressed buffer");
    widest = std::max(widest, static_cast<std::size_t>(b->size()));
    pageable = pageable || (b->size() > 0 && bitar_ptr_kind(b->data(), nullptr) == 0);
  }
  const std::size_t stride = (widest + 16 + 255) & ~static_cast<std::size_t>(255);
  if (pageable && q.staged_cap < n * stride) {
    if (q.staged != nullptr) bitar_mem_free(BITAR_MEM_PINNED, device_id_, q.staged);
    q.staged = nullptr;
    q.staged_cap = 0;
    ARROW_RETURN_NOT_OK(internal::StatusFromC(bitar_mem_alloc(BITAR_MEM_PINNED, device_id_, n * stride, 256, &q.staged)));
    q.staged_cap = n * stride;
  }
  ARROW_RETURN_NOT_OK(MakeAccessible(decompressed_buffer->mutable_data(), n * seg, &q.registered));
  q.ops.resize(n);
  q.results.assign(n, bitar_result{0, BITAR_OP_NOT_RUN, 0});
  for (std::size_t i = 0; i < n; ++i) {   // AssembleFrom(buffers, index, span, offset), src/memory.cc:432-505
    const auto& b = compressed_buffers[i];
    if (b->size() > 0 && bitar_ptr_kind(b->data(), nullptr) == 0) {
      std::memcpy(static_cast<std::uint8_t*>(q.staged) + i * stride, b->data(), static_cast<std::size_t>(b->size()));
      q.ops[i].src = static_cast<std::uint8_t*>(q.staged) + i * stride;
    } else {
      q.ops[i].src = b->data();
    }
    q.ops[i].src_len = static_cast<std::uint32_t>(b->size());
    q.ops[i].dst = decompressed_buffer->mutable_data() + i * seg;
    q.ops[i].dst_cap = static_cast<std::uint32_t>(seg);
  }
  return internal::StatusFromC(
      bitar_qp_inflate(handle_, queue_pair_id, q.ops.data(), static_cast<std::uint32_t>(n), q.results.data()));
}

// Chained segments (AssembleFrom(buffers, ...) with max_sgl_segs > 1, src/memory.cc:432-505): the buffers of one stream are
// consecutive full slots and end with a shorter one; they must be one contiguous, device-accessible range.  Stream g
// inflates to decompressed_buffer + g * k * S.
template <typename Class>
arrow::Status CompressDevice<Class>::EnqueueDecompressChained(std::uint16_t queue_pair_id, const BufferVector& compressed_buffers,
                                                              const std::unique_ptr<arrow::ResizableBuffer>& decompressed_buffer) {
  auto& q = qp_state_[queue_pair_id < num_qps_ ? queue_pair_id : 0];
  const auto seg = static_cast<std::size_t>(configuration_->decompressed_seg_size());
  const std::size_t k = configuration_->max_sgl_segs(), slot = bitar_slot_size(handle_);
  std::vector<bitar_chunk> ops;
  for (std::size_t i = 0; i < compressed_buffers.size();) {
    if (compressed_buffers[i] == nullptr) return arrow::Status::Invalid("null compressed buffer");
    const std::uint8_t* first = compressed_buffers[i]->data();
    std::size_t total = 0;
    for (;;) {
      const auto& b = compressed_buffers[i];
      if (b == nullptr || b->data() != first + total)
        return arrow::Status::Invalid("The compressed buffers of a chained operation are not contiguous");
      total += static_cast<std::size_t>(b->size());
      ++i;
      if (static_cast<std::size_t>(b->size()) < slot) break;
      if (i == compressed_buffers.size())
        return arrow::Status::Invalid("The last chained operation has no end (a buffer shorter than a slot)");
    }
    if (bitar_ptr_kind(first, nullptr) == 0) return arrow::Status::Invalid("Chained compressed buffers must be device-accessible memory");
    ops.push_back(bitar_chunk{first, nullptr, static_cast<std::uint32_t>(total), static_cast<std::uint32_t>(k * seg)});
  }
  const std::size_t need = std::max(ops.size() * k * seg, compressed_buffers.size() * seg);   // (the second: src/device.cc:248-254)
  if (decompressed_buffer == nullptr || static_cast<std::size_t>(decompressed_buffer->capacity()) < ops.size() * k * seg)
    return arrow::Status::CapacityError("The decompressed_buffer is required to be >= ", need, " bytes");
  ARROW_RETURN_NOT_OK(EntryGuard(queue_pair_id));
  Unregister(queue_pair_id);
  ARROW_RETURN_NOT_OK(MakeAccessible(decompressed_buffer->mutable_data(), ops.size() * k * seg, &q.registered));
  for (std::size_t g = 0; g < ops.size(); ++g) ops[g].dst = decompressed_buffer->mutable_data() + g * k * seg;
  q.ops = std::move(ops);
  q.results.assign(q.ops.size(), bitar_result{0, BITAR_OP_NOT_RUN, 0});
  return internal::StatusFromC(
      bitar_qp_inflate(handle_, queue_pair_id, q.ops.data(), static_cast<std::uint32_t>(q.ops.size()), q.results.data()));
}

template <typename Class>
arrow::Status CompressDevice<Class>::FinishDecompress(std::uint16_t queue_pair_id,
                                                      const std::unique_ptr<arrow::ResizableBuffer>& decompressed_buffer,
                                                      bool in_callback) {
  auto& q = qp_state_[queue_pair_id];
  if (q.ops.empty()) return arrow::Status::OK();
  const int rc = in_callback ? bitar_qp_result(handle_, queue_pair_id) : bitar_qp_wait(handle_, queue_pair_id);
  if (!in_callback) Unregister(queue_pair_id);
  // unlike the reference (src/device.cc:275-310 leaves the queue pair busy after an error) the queue pair is idle again
  ARROW_RETURN_NOT_OK(internal::StatusFromC(rc));
  std::int64_t total = 0;
  for (const auto& r : q.results) total += r.produced;
  return decompressed_buffer->Resize(total, /*shrink_to_fit=*/false);   // src/device.cc:315
}

template <typename Class>
arrow::Status CompressDevice<Class>::Decompress(std::uint16_t queue_pair_id, const BufferVector& compressed_buffers,
                                                const std::unique_ptr<arrow::ResizableBuffer>& decompressed_buffer) {
  if (compressed_buffers.empty()) return arrow::Status::OK();
  ARROW_RETURN_NOT_OK(EnqueueDecompress(queue_pair_id, compressed_buffers, decompressed_buffer));
  return FinishDecompress(queue_pair_id, decompressed_buffer);
}

template <typename Class>
std::size_t CompressDevice<Class>::Recycle(const BufferVector& buffers) {   // src/device.cc:320-327
  std::size_t recycled = 0;
  for (auto it = buffers.rbegin(); it != buffers.rend(); ++it)
    if (*it != nullptr) recycled += static_cast<std::size_t>(bitar_slot_put(handle_, (*it)->data()));
  return recycled;
}

template <typename Class>
void* CompressDevice<Class>::StreamOf(std::uint16_t queue_pair_id) const {
  return handle_ != nullptr ? bitar_qp_stream(handle_, queue_pair_id) : nullptr;
}

template <typename Class>
const std::vector<bitar_result>& CompressDevice<Class>::LastResults(std::uint16_t queue_pair_id) const {
  return qp_state_[queue_pair_id].results;
}

template <typename Class>
arrow::Status CompressDevice<Class>::LastElapsedMs(std::uint16_t queue_pair_id, float* kernel_ms, float* total_ms) const {
  return internal::StatusFromC(bitar_qp_last_ms(handle_, queue_pair_id, kernel_ms, total_ms));
}

template <typename Class>
arrow::Status CompressDevice<Class>::OnComplete(std::uint16_t queue_pair_id, void (*fn)(void*), void* arg) {
  return internal::StatusFromC(bitar_qp_on_complete(handle_, queue_pair_id, fn, arg));
}

template class CompressDevice<Class_CUDA>;

// ---- CudaCompressDevice (the BlueFieldCompressDevice analogue, src/device.cc:558-588) ----
arrow::Status CudaCompressDevice::ValidateConfiguration() {
  ARROW_RETURN_NOT_OK(CudaCompressDeviceBase::ValidateConfiguration());
  const auto* cfg = dynamic_cast<const CudaConfiguration*>(configuration().get());
  if (cfg == nullptr) return arrow::Status::Invalid("Configuration is not a CudaConfiguration");
  return arrow::Status::OK();
}

arrow::Status CudaCompressDevice::set_configuration(std::unique_ptr<Configuration<Class_CUDA>> configuration) {
  if (configuration == nullptr || configuration->type_name() != kCudaConfigurationTypeName)   // src/device.cc:579-588
    return arrow::Status::Invalid("Unable to set configuration of type [", configuration ? configuration->type_name() : "null",
                                  "] to a CUDA compress device");
  return CudaCompressDeviceBase::set_configuration(std::move(configuration));
}

DeviceManager* DeviceManager::Instance() {
  static DeviceManager instance;
  return &instance;
}

arrow::Result<CompressDevice<Class_CUDA>*> DeviceManager::Create(std::uint8_t device_id, std::uint16_t num_qps) {
  bitar_dev_info info{};
  ARROW_RETURN_NOT_OK(internal::StatusFromC(bitar_cuda_device_info(device_id, &info)));
  if (info.cc_major < 10)   // DeviceManager::Create default -> NotImplemented, src/include/device.h:203-207
    return arrow::Status::NotImplemented("Unsupported compress device ", static_cast<unsigned>(device_id), " (", info.name,
                                         ", sm_", info.cc_major, info.cc_minor, "): the kernels are built for sm_100a");
  return new CudaCompressDevice(device_id, num_qps);
}

}  // namespace bitar
