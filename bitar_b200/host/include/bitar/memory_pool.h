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

enum class MemoryPoolBackend : std::uint8_t { System, CudaPinnedHost, CudaDevice };

/// \brief Get the memory pool for the selected backend (/root/reference/src/include/memory_pool.h:65-74).
/// CudaPinnedHost is the Rtememzone analogue for buffers the CPU also touches: memory the device reads and writes in
/// place (zero-copy over PCIe).  CudaDevice hands out memory of the calling thread's current CUDA device
/// (cudaMallocAsync): Allocate / Reallocate (device-side copy) / Free and the pool statistics work like any pool's, but
/// the bytes are not addressable by the CPU -- Arrow's own allocation helpers (arrow::AllocateBuffer, builders) zero the
/// padding of what they allocate on the CPU and must not be given this pool; use AllocateDeviceBuffer() below.
arrow::MemoryPool* GetMemoryPool(MemoryPoolBackend backend);

/// \brief A resizable buffer in the memory of CUDA device \p device_id, allocated from the CudaDevice pool (so that
/// it shows in the pool's statistics and the tracker), for callers that keep their data resident in HBM.  Typed as a
/// CPU arrow::Buffer whose data() is a device address (Arrow C++ in this image ships without arrow::cuda); never
/// dereference it on the host.  Resize() / Reserve() never touch the bytes from the CPU.
arrow::Result<std::unique_ptr<arrow::ResizableBuffer>> AllocateDeviceBuffer(std::int64_t size, int device_id);

}  // namespace bitar
