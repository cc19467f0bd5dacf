// memory_pool.h -- arrow::MemoryPool back ends handing out device-accessible memory.
// Mirrors /root/reference/src/include/memory_pool.h:36-74 (RtemallocAllocator / RtememzoneAllocator and
// the address tracker) with CUDA allocators: cudaMallocAsync device memory and pinned host memory.
#pragma once
#include <cstddef>
#include <cstdint>
#include <mutex>
#include <unordered_map>

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

enum class MemoryPoolBackend : std::uint8_t { System, CudaDevice, CudaPinnedHost };

/// \brief Get the memory pool for the selected backend (CudaDevice uses the current CUDA device).
arrow::MemoryPool* GetMemoryPool(MemoryPoolBackend backend);

}  // namespace bitar
