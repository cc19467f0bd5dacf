// util.h -- asynchronous Compress/Decompress.  Mirrors /root/reference/src/include/util.h:45-101,133-151,
// 216-236: the call is enqueued on the queue pair's CUDA stream and the user callback runs on a CUDA
// driver thread when it completes (the worker-lcore analogue); WaitForAsync() is rte_eal_wait_lcore().
#pragma once
#include <arrow/result.h>
#include <arrow/status.h>

#include <atomic>
#include <cerrno>
#include <condition_variable>
#include <cstdint>
#include <memory>
#include <mutex>

#include "config.h"
#include "device.h"
#include "type_fwd.h"

namespace bitar {

static inline constexpr auto kAsyncReturnOK = 2;   // util.h:45

namespace internal {
struct AsyncSlot {                                  // completion state shared with the driver thread
  std::mutex mu;
  std::condition_variable cv;
  bool done = false;
  int ret = 0;                                      // 0 == never ran (apps/demo_app.cc:267-274)
  void Finish(int r) {
    { std::lock_guard<std::mutex> lock(mu); ret = r; done = true; }
    cv.notify_all();
  }
  int Wait() {
    std::unique_lock<std::mutex> lock(mu);
    cv.wait(lock, [&] { return done; });
    return ret;
  }
};
}  // namespace internal

/// The signature of the result_callback should be equivalent to:
///   int func(std::uint8_t device_id, std::uint16_t queue_pair_id, arrow::Result<bitar::BufferVector>&& result);
/// Params hold REFERENCES: device, buffer and callback must outlive the call (util.h:69-72).
template <typename Class, typename Callback>
struct CompressParam {
  CompressParam(const std::unique_ptr<bitar::CompressDevice<Class>>& device, std::uint16_t queue_pair_id,
                const std::shared_ptr<arrow::Buffer>& decompressed_buffer, const Callback& result_callback)
      : device_{device}, queue_pair_id_{queue_pair_id}, decompressed_buffer_{decompressed_buffer},
        result_callback_{result_callback} {}
  const std::unique_ptr<bitar::CompressDevice<Class>>& device_;
  const std::uint16_t queue_pair_id_{};
  const std::shared_ptr<arrow::Buffer>& decompressed_buffer_;
  const Callback& result_callback_;
  internal::AsyncSlot slot_;
};

/// result_callback: int func(std::uint8_t device_id, std::uint16_t queue_pair_id, const arrow::Status& status);
template <typename Class, typename Callback>
struct DecompressParam {
  DecompressParam(const std::unique_ptr<bitar::CompressDevice<Class>>& device, std::uint16_t queue_pair_id,
                  const BufferVector& compressed_buffers,
                  const std::unique_ptr<arrow::ResizableBuffer>& decompressed_buffer, const Callback& result_callback)
      : device_{device}, queue_pair_id_{queue_pair_id}, compressed_buffers_{compressed_buffers},
        decompressed_buffer_{decompressed_buffer}, result_callback_{result_callback} {}
  const std::unique_ptr<bitar::CompressDevice<Class>>& device_;
  const std::uint16_t queue_pair_id_{};
  const BufferVector& compressed_buffers_;
  const std::unique_ptr<arrow::ResizableBuffer>& decompressed_buffer_;
  const Callback& result_callback_;
  internal::AsyncSlot slot_;
};

namespace internal {
template <typename Class, typename Callback>
void CompressDone(void* p) {
  auto* param = static_cast<CompressParam<Class, Callback>*>(p);
  auto result = param->device_->FinishCompress(param->queue_pair_id_, /*in_callback=*/true);
  param->slot_.Finish(param->result_callback_(param->device_->device_id(), param->queue_pair_id_, std::move(result)));
}
template <typename Class, typename Callback>
void DecompressDone(void* p) {
  auto* param = static_cast<DecompressParam<Class, Callback>*>(p);
  auto status = param->device_->FinishDecompress(param->queue_pair_id_, param->decompressed_buffer_, /*in_callback=*/true);
  param->slot_.Finish(param->result_callback_(param->device_->device_id(), param->queue_pair_id_, status));
}
}  // namespace internal

/// \brief Asynchronous CompressDevice::Compress().  \return 0 if started, -EBUSY if the queue pair still
/// has a call in flight (util.h:211-221), another negative value if the call could not be enqueued.
template <typename Class, typename Callback>
int CompressAsync(const std::unique_ptr<CompressParam<Class, Callback>>& param) {
  auto st = param->device_->EnqueueCompress(param->queue_pair_id_, param->decompressed_buffer_);
  if (st.IsCancelled()) return -EBUSY;
  if (!st.ok()) return -static_cast<int>(st.code());
  st = param->device_->OnComplete(param->queue_pair_id_, &internal::CompressDone<Class, Callback>, param.get());
  return st.ok() ? 0 : -static_cast<int>(st.code());
}

template <typename Class, typename Callback>
int DecompressAsync(const std::unique_ptr<DecompressParam<Class, Callback>>& param) {
  auto st = param->device_->EnqueueDecompress(param->queue_pair_id_, param->compressed_buffers_, param->decompressed_buffer_);
  if (st.IsCancelled()) return -EBUSY;
  if (!st.ok()) return -static_cast<int>(st.code());
  st = param->device_->OnComplete(param->queue_pair_id_, &internal::DecompressDone<Class, Callback>, param.get());
  return st.ok() ? 0 : -static_cast<int>(st.code());
}

/// \brief rte_eal_wait_lcore() analogue: blocks until the callback ran and returns its int.
template <typename Param>
int WaitForAsync(const std::unique_ptr<Param>& param) { return param->slot_.Wait(); }

}  // namespace bitar
