// device.h -- CompressDevice<Class_CUDA>: one CUDA device, its queue pairs (CUDA streams) and its
// output-slot pool.  Mirrors /root/reference/src/include/device.h:83-134,191-240 and src/device.cc:114-588.
#pragma once
#include <arrow/result.h>
#include <arrow/status.h>

#include <cstddef>
#include <cstdint>
#include <memory>
#include <vector>

#include "bitar_cuda.h"
#include "config.h"
#include "type_fwd.h"

namespace arrow {
class Buffer;
class ResizableBuffer;
}  // namespace arrow

namespace bitar {

namespace internal {
enum class DeviceState { kUndefined = 1U << 0U, kConfigured = 1U << 1U, kStarted = 1U << 2U };  // device.h:64-68
/// \brief arrow::Status from the negative StatusCode a C-ABI call returned (src/include/util.h:166-205).
arrow::Status StatusFromC(int rc);
}  // namespace internal

template <typename Class>
class CompressDevice {
 public:
  CompressDevice(const CompressDevice&) = delete;
  CompressDevice& operator=(const CompressDevice&) = delete;

  /// \brief Initialize this compress device with a corresponding type of configuration
  /// (validate -> preallocate slots -> set up the queue pairs -> start), src/device.cc:114-154.
  virtual arrow::Status Initialize(std::unique_ptr<Configuration<Class>> configuration);

  /// \brief Compress a buffer via \p queue_pair_id: one complete raw DEFLATE stream per
  /// decompressed_seg_size() bytes, returned in input order as non-owning views into pool slots that
  /// the caller must Recycle() (src/device.cc:156-238).  Null/empty input returns {}.
  arrow::Result<BufferVector> Compress(std::uint16_t queue_pair_id,
                                       const std::shared_ptr<arrow::Buffer>& decompressed_buffer);

  /// \brief Decompress buffers via \p queue_pair_id: buffer i inflates to offset i * seg of
  /// \p decompressed_buffer, which is then resized to the total (src/device.cc:240-318).
  arrow::Status Decompress(std::uint16_t queue_pair_id, const BufferVector& compressed_buffers,
                           const std::unique_ptr<arrow::ResizableBuffer>& decompressed_buffer);

  /// \brief Return the slots of buffers produced by Compress(); returns how many were recycled
  /// (src/device.cc:320-327).
  std::size_t Recycle(const BufferVector& buffers);

  /// \brief The queue pair's cudaStream_t (the analogue of LcoreOf(), device.h:121-124).
  [[nodiscard]] void* StreamOf(std::uint16_t queue_pair_id) const;

  [[nodiscard]] auto device_id() const noexcept { return device_id_; }
  [[nodiscard]] std::uint16_t num_qps() const noexcept { return num_qps_; }

  /// \brief Per-op results of the last call on a queue pair (produced, status, checksum): the value the
  /// reference drops (rte_comp_op::output_chksum is never read, src/memory.cc:106-116).
  [[nodiscard]] const std::vector<bitar_result>& LastResults(std::uint16_t queue_pair_id) const;
  /// \brief Device time (ms) of the last call on a queue pair: kernel only / whole call.
  arrow::Status LastElapsedMs(std::uint16_t queue_pair_id, float* kernel_ms, float* total_ms) const;

  virtual ~CompressDevice();

  // --- asynchronous halves used by CompressAsync / DecompressAsync (util.h) ---
  struct PendingCompress;
  arrow::Status EnqueueCompress(std::uint16_t queue_pair_id, const std::shared_ptr<arrow::Buffer>& decompressed_buffer);
  /// \p in_callback: called from an OnComplete() callback (a CUDA driver thread, where no CUDA call is allowed):
  /// the result is read without waiting and a temporarily registered buffer stays registered until the next
  /// call on the queue pair.
  arrow::Result<BufferVector> FinishCompress(std::uint16_t queue_pair_id, bool in_callback = false);
  arrow::Status EnqueueDecompress(std::uint16_t queue_pair_id, const BufferVector& compressed_buffers,
                                  const std::unique_ptr<arrow::ResizableBuffer>& decompressed_buffer);
  arrow::Status EnqueueDecompressChained(std::uint16_t queue_pair_id, const BufferVector& compressed_buffers,
                                  const std::unique_ptr<arrow::ResizableBuffer>& decompressed_buffer);
  arrow::Status FinishDecompress(std::uint16_t queue_pair_id, const std::unique_ptr<arrow::ResizableBuffer>& decompressed_buffer,
                                 bool in_callback = false);
  arrow::Status OnComplete(std::uint16_t queue_pair_id, void (*fn)(void*), void* arg);

 protected:
  CompressDevice(std::uint8_t device_id, std::uint16_t num_qps);
  virtual arrow::Status ValidateConfiguration();
  [[nodiscard]] const auto& configuration() const noexcept { return configuration_; }
  virtual arrow::Status set_configuration(std::unique_ptr<Configuration<Class>> configuration);
  [[nodiscard]] auto state() const noexcept { return state_; }
  void set_state(internal::DeviceState state) { state_ = state; }

 private:
  arrow::Status EntryGuard(std::uint16_t queue_pair_id);
  void ReleaseSlots(std::uint16_t queue_pair_id);
  void Unregister(std::uint16_t queue_pair_id);

  struct QueuePairState {
    std::vector<bitar_chunk> ops;
    std::vector<bitar_result> results;
    std::vector<void*> slots;
    void* registered = nullptr;   // pageable input registered for the duration of a call
    void* staged = nullptr;       // pinned copy of compressed buffers that arrived in pageable memory (Decompress)
    std::size_t staged_cap = 0;
  };

  const std::uint8_t device_id_;
  const std::uint16_t num_qps_;
  std::unique_ptr<Configuration<Class>> configuration_;
  internal::DeviceState state_{internal::DeviceState::kUndefined};
  bitar_dev* handle_ = nullptr;
  std::vector<QueuePairState> qp_state_;
};

class DeviceManager {
 public:
  static DeviceManager* Instance();
  /// \brief Create the device object for a probed CUDA device (the Create<Vendor,Device,Class> analogue,
  /// device.h:196-219): NotImplemented unless the device is compute capability 10.x.
  arrow::Result<CompressDevice<Class_CUDA>*> Create(std::uint8_t device_id, std::uint16_t num_qps);
};

using CudaCompressDeviceBase = CompressDevice<Class_CUDA>;

class CudaCompressDevice : public CudaCompressDeviceBase {
 public:
  ~CudaCompressDevice() override = default;

 protected:
  using CompressDevice::CompressDevice;
  arrow::Status ValidateConfiguration() override;
  arrow::Status set_configuration(std::unique_ptr<Configuration<Class_CUDA>> configuration) override;
  friend class DeviceManager;
};

}  // namespace bitar
