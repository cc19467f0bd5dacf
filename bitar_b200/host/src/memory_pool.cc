// memory_pool.cc -- arrow::MemoryPool back ends over the C-ABI allocators.
// Mirrors /root/reference/src/memory_pool.cc:70-350: an allocator per backend (there rte_malloc and
// rte_memzone, here cudaMallocAsync device memory and pinned host memory), the statistics every Arrow pool
// keeps, the zero-size sentinel (memory_pool.cc:58-67) and the address tracker (memory_pool.cc:295-319).
#include "bitar/memory_pool.h"

#include <arrow/buffer.h>
#include <arrow/memory_pool.h>
#include <arrow/result.h>
#include <arrow/status.h>

#include <atomic>
#include <cstring>
#include <string>

#include "bitar_cuda.h"

namespace bitar {

namespace {

alignas(64) std::uint8_t zero_size_area[64];   // what zero-byte allocations point at

class CudaMemoryPool : public arrow::MemoryPool {
 public:
  CudaMemoryPool(int kind, std::string name) : kind_{kind}, name_{std::move(name)} {}

  arrow::Status Allocate(std::int64_t size, std::int64_t alignment, std::uint8_t** out) override {
    if (size < 0) return arrow::Status::Invalid("negative malloc size");
    if (size == 0) {
      *out = zero_size_area;
      return arrow::Status::OK();
    }
    void* p = nullptr;
    int device = -1;
    if (kind_ == BITAR_MEM_DEVICE && bitar_current_device(&device) != BITAR_OK)
      return arrow::Status::Invalid("no current CUDA device: ", bitar_last_error());
    const int rc = bitar_mem_alloc(kind_, device, static_cast<std::size_t>(size), static_cast<std::size_t>(alignment), &p);
    if (rc != BITAR_OK) return arrow::Status::OutOfMemory("malloc of size ", size, " failed: ", bitar_last_error());
    *out = static_cast<std::uint8_t*>(p);
    CudaAllocatorTracker::Instance()->Emplace(*out, {static_cast<std::size_t>(size), kind_, device});
    Did(size);
    return arrow::Status::OK();
  }

  // realloc = alloc + copy + free, like the Rtememzone allocator (memory_pool.cc:151-174); the copy runs on
  // the device for device memory
  arrow::Status Reallocate(std::int64_t old_size, std::int64_t new_size, std::int64_t alignment, std::uint8_t** ptr) override {
    std::uint8_t* old = *ptr;
    std::uint8_t* fresh = nullptr;
    ARROW_RETURN_NOT_OK(Allocate(new_size, alignment, &fresh));
    const std::int64_t n = old_size < new_size ? old_size : new_size;
    if (n > 0) {
      if (kind_ == BITAR_MEM_DEVICE) {
        if (bitar_mem_copy(fresh, old, static_cast<std::size_t>(n)) != BITAR_OK) {
          Free(fresh, new_size, alignment);
          return arrow::Status::IOError("device copy failed: ", bitar_last_error());
        }
      } else {
        std::memcpy(fresh, old, static_cast<std::size_t>(n));
      }
    }
    Free(old, old_size, alignment);
    *ptr = fresh;
    return arrow::Status::OK();
  }

  void Free(std::uint8_t* buffer, std::int64_t size, std::int64_t /*alignment*/) override {
    if (buffer == zero_size_area || buffer == nullptr) return;
    const auto* a = CudaAllocatorTracker::Instance()->Of(buffer);
    const int device = a != nullptr ? a->device : 0;
    bitar_mem_free(kind_, device, buffer);
    CudaAllocatorTracker::Instance()->Release(buffer);
    bytes_.fetch_sub(size);
  }

  std::int64_t bytes_allocated() const override { return bytes_.load(); }
  std::int64_t max_memory() const override { return max_.load(); }
  std::int64_t total_bytes_allocated() const override { return total_.load(); }
  std::int64_t num_allocations() const override { return count_.load(); }
  std::string backend_name() const override { return name_; }

 private:
  void Did(std::int64_t size) {
    const auto now = bytes_.fetch_add(size) + size;
    auto prev = max_.load();
    while (now > prev && !max_.compare_exchange_weak(prev, now)) {
    }
    total_.fetch_add(size);
    count_.fetch_add(1);
  }
  const int kind_;
  const std::string name_;
  std::atomic<std::int64_t> bytes_{0}, max_{0}, total_{0}, count_{0};
};

// device-resident ResizableBuffer over the CudaDevice pool: Resize() never touches the bytes from the CPU
class DeviceBuffer : public arrow::ResizableBuffer {
 public:
  DeviceBuffer(std::uint8_t* data, std::int64_t size, int device, arrow::MemoryPool* pool)
      : arrow::ResizableBuffer(data, size), device_{device}, pool_{pool} {
    capacity_ = size;
  }
  ~DeviceBuffer() override {
    if (data_ != nullptr) pool_->Free(const_cast<std::uint8_t*>(data_), capacity_, 256);
  }
  arrow::Status Resize(const std::int64_t new_size, bool /*shrink_to_fit*/) override {
    if (new_size < 0) return arrow::Status::Invalid("Negative buffer resize: ", new_size);
    if (new_size > capacity_) ARROW_RETURN_NOT_OK(Reserve(new_size));
    size_ = new_size;
    return arrow::Status::OK();
  }
  arrow::Status Reserve(const std::int64_t new_capacity) override {
    if (new_capacity <= capacity_) return arrow::Status::OK();
    std::uint8_t* p = const_cast<std::uint8_t*>(data_);
    ARROW_RETURN_NOT_OK(pool_->Reallocate(capacity_, new_capacity, 256, &p));   // alloc + device-side copy + free
    data_ = p;
    capacity_ = new_capacity;
    return arrow::Status::OK();
  }

 private:
  const int device_;
  arrow::MemoryPool* const pool_;
};

}  // namespace

arrow::Result<std::unique_ptr<arrow::ResizableBuffer>> AllocateDeviceBuffer(std::int64_t size, int device_id) {
  if (size < 0) return arrow::Status::Invalid("negative size");
  int current = -1;
  if (bitar_current_device(&current) != BITAR_OK) return arrow::Status::Invalid("no current CUDA device: ", bitar_last_error());
  if (current != device_id && bitar_set_device(device_id) != BITAR_OK)
    return arrow::Status::Invalid("cannot select CUDA device ", device_id, ": ", bitar_last_error());
  auto* pool = GetMemoryPool(MemoryPoolBackend::CudaDevice);
  std::uint8_t* p = nullptr;
  auto st = pool->Allocate(size > 0 ? size : 1, 256, &p);
  if (current != device_id) bitar_set_device(current);
  ARROW_RETURN_NOT_OK(st);
  return std::unique_ptr<arrow::ResizableBuffer>(new DeviceBuffer(p, size, device_id, pool));
}

CudaAllocatorTracker* CudaAllocatorTracker::Instance() {
  static CudaAllocatorTracker tracker;
  return &tracker;
}
const CudaAllocatorTracker::Allocation* CudaAllocatorTracker::Of(const std::uint8_t* addr) const noexcept {
  std::lock_guard<std::mutex> lock(mutex_);
  const auto it = allocations_.find(addr);
  return it == allocations_.end() ? nullptr : &it->second;
}
std::size_t CudaAllocatorTracker::count() const noexcept {
  std::lock_guard<std::mutex> lock(mutex_);
  return allocations_.size();
}
void CudaAllocatorTracker::Emplace(const std::uint8_t* addr, Allocation a) {
  std::lock_guard<std::mutex> lock(mutex_);   // the reference's Emplace is unlocked (a race): not copied
  allocations_[addr] = a;
}
void CudaAllocatorTracker::Release(const std::uint8_t* addr) {
  std::lock_guard<std::mutex> lock(mutex_);
  allocations_.erase(addr);
}

arrow::MemoryPool* GetMemoryPool(MemoryPoolBackend backend) {   // memory_pool.cc:321-350: process-lifetime statics
  switch (backend) {
    case MemoryPoolBackend::CudaPinnedHost: {
      static CudaMemoryPool pool(BITAR_MEM_PINNED, "cuda_pinned_host");
      return &pool;
    }
    case MemoryPoolBackend::CudaDevice: {
      static CudaMemoryPool pool(BITAR_MEM_DEVICE, "cuda_device");
      return &pool;
    }
    case MemoryPoolBackend::System:
    default:
      return arrow::system_memory_pool();
  }
}

}  // namespace bitar
