// memory_pool.h -- arrow::MemoryPool back ends handing out device-accessible memory.
// Mirrors /root/reference/src/include/memory_pool.h:36-74 (RtemallocAllocator / RtememzoneAllocator and
// the address tracker) with CUDA allocators: cudaMallocAsync device memory and pinned host memory.
#pragma once
#include <cstddef>
#include <cstdint>
#include <mutex>
#include <unordered_map>

#include <arrow/buffer.h>
#include <arrow/result.h>

#include <memory>

namespace arrow {
class MemoryPool;
}  // namespace arrow

namespace bitar {

/// \brief Maps a base address to its allocation (the RtememzoneAllocatorTracker analogue).
class CudaAllocatorTracker {
 public:
  struct Allocation {
    std::size_t size;
    int kind;    // BITAR_MEM_*
    int device;
  };
  /// \brief Look up the allocation that starts at \p addr; nullptr if unknown.
  const Allocation* Of(const std::uint8_t* addr) const noexcept;
  [[nodiscard]] std::size_t count() const noexcept;
  static CudaAllocatorTracker* Instance();

  void Emplace(const std::uint8_t* addr, Allocation a);
  void Release(const std::uint8_t* addr);

 private:
  std::unordered_map<const std::uint8_t*, Allocation> allocations_{};
  mutable std::mutex mutex_{};
};

enum class MemoryPoolBackend : std::uint8_t { System, CudaPinnedHost };

/// \brief Get the memory pool for the selected backend.  CudaPinnedHost is the Rtememzone analogue: memory the
/// device reads and writes in place (zero-copy over PCIe) and the CPU can touch, which Arrow's allocation
/// helpers do (they zero the padding of every buffer on the CPU) -- the reason there is no device-memory POOL.
arrow::MemoryPool* GetMemoryPool(MemoryPoolBackend backend);

/// \brief A resizable buffer in the memory of CUDA device \p device_id (cudaMallocAsync), for callers that keep
/// their data resident in HBM.  Typed as a CPU arrow::Buffer whose data() is a device address (Arrow C++ in this
/// image ships without arrow::cuda); never dereference it on the host.
arrow::Result<std::unique_ptr<arrow::ResizableBuffer>> AllocateDeviceBuffer(std::int64_t size, int device_id);

}  // namespace bitar
