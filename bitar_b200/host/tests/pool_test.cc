// pool_test -- the CudaDevice arrow::MemoryPool and AllocateDeviceBuffer() (needs a GPU; run by tests/test_host_cpp.py).
#include <arrow/buffer.h>
#include <arrow/memory_pool.h>

#include <cstdio>
#include <cstring>
#include <vector>

#include "bitar/memory_pool.h"
#include "bitar_cuda.h"

#define EXPECT(c) do { if (!(c)) { std::fprintf(stderr, "FAILED: %s (line %d): %s\n", #c, __LINE__, bitar_last_error()); return 1; } } while (0)

int main() {
  auto* pool = bitar::GetMemoryPool(bitar::MemoryPoolBackend::CudaDevice);
  EXPECT(pool->backend_name() == "cuda_device");
  const auto tracked0 = bitar::CudaAllocatorTracker::Instance()->count();
  std::uint8_t* p = nullptr;
  EXPECT(pool->Allocate(1 << 20, 256, &p).ok() && p != nullptr);
  int dev = -1;
  EXPECT(bitar_ptr_kind(p, &dev) == 1);                       // device memory
  EXPECT(pool->bytes_allocated() == (1 << 20) && pool->num_allocations() == 1);
  std::vector<std::uint8_t> host(1 << 20), back(3 << 20);
  for (std::size_t i = 0; i < host.size(); ++i) host[i] = (std::uint8_t)(i * 31 + 7);
  EXPECT(bitar_mem_copy(p, host.data(), host.size()) == BITAR_OK);
  EXPECT(pool->Reallocate(1 << 20, 3 << 20, 256, &p).ok());   // device-side copy
  EXPECT(pool->bytes_allocated() == (3 << 20) && pool->max_memory() >= (4 << 20));
  EXPECT(bitar_mem_copy(back.data(), p, 1 << 20) == BITAR_OK && std::memcmp(back.data(), host.data(), 1 << 20) == 0);
  pool->Free(p, 3 << 20, 256);
  EXPECT(pool->bytes_allocated() == 0);
  std::uint8_t* z = nullptr;
  EXPECT(pool->Allocate(0, 64, &z).ok() && z != nullptr);     // the zero-size sentinel (src/memory_pool.cc:58-67)
  pool->Free(z, 0, 64);
  {
    auto r = bitar::AllocateDeviceBuffer(100000, 0);
    EXPECT(r.ok());
    auto buf = std::move(*r);
    EXPECT(buf->size() == 100000 && bitar_ptr_kind(buf->data(), &dev) == 1 && pool->bytes_allocated() >= 100000);
    EXPECT(bitar_mem_copy(const_cast<std::uint8_t*>(buf->data()), host.data(), 100000) == BITAR_OK);
    EXPECT(buf->Resize(5 << 20, false).ok() && buf->size() == (5 << 20));
    EXPECT(bitar_mem_copy(back.data(), buf->data(), 100000) == BITAR_OK && std::memcmp(back.data(), host.data(), 100000) == 0);
    EXPECT(bitar::CudaAllocatorTracker::Instance()->count() == tracked0 + 1);
  }
  EXPECT(pool->bytes_allocated() == 0 && bitar::CudaAllocatorTracker::Instance()->count() == tracked0);
  std::printf("device pool: OK (%lld allocations, %lld bytes in total)\n", (long long)pool->num_allocations(),
              (long long)pool->total_bytes_allocated());
  return 0;
}
