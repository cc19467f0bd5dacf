// bitar_demo -- the demo_app flow of the reference on CUDA devices (apps/demo_app.cc:332-357, 487-693):
//   EvaluateSync : device 0 / queue pair 0, kNumTests x { Compress, Decompress, memcmp, Recycle }
//   EvaluateAsync: the buffer split evenly over every (device, queue pair), CompressAsync /
//                  DecompressAsync with a callback per queue pair, per-part memcmp
// and prints microseconds and Gbps (bits) like apps/demo_app.cc:82-89.  Input: --file raw bytes, or a
// synthetic columnar mix (--bytes).  Buffers live in the pinned-host pool (the Rtememzone analogue:
// zero-copy for the device) or, with --device, in the device pool.
#include <arrow/buffer.h>
#include <arrow/io/file.h>
#include <arrow/io/memory.h>
#include <arrow/ipc/feather.h>
#include <arrow/ipc/reader.h>
#include <arrow/ipc/writer.h>
#include <arrow/memory_pool.h>
#include <arrow/result.h>
#include <arrow/status.h>
#include <arrow/table.h>
#include <parquet/arrow/reader.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "bitar/config.h"
#include "bitar/device.h"
#include "bitar/driver.h"
#include "bitar/memory_pool.h"
#include "bitar/util.h"

namespace {

constexpr int kNumTests = 3;                 // apps/demo_app.h:45
using Device = bitar::CompressDevice<bitar::Class_CUDA>;
using Clock = std::chrono::steady_clock;

#define CHECK_OK(expr)                                                       \
  do {                                                                       \
    auto st_ = (expr);                                                       \
    if (!st_.ok()) {                                                         \
      std::fprintf(stderr, "%s: %s\n", #expr, st_.ToString().c_str());       \
      std::exit(EXIT_FAILURE);                                               \
    }                                                                        \
  } while (0)

void PrintPerf(const char* what, std::int64_t bytes, Clock::time_point t0, Clock::time_point t1) {
  const double us = std::chrono::duration<double, std::micro>(t1 - t0).count();
  std::printf("  %-22s %12.0f us  %8.2f Gbps  (%.2f GB/s)\n", what, us, bytes * 8.0 / 1e3 / us, bytes / 1e3 / us);
}

// a lineitem-like mix: sorted int64 keys, dictionary int32 codes, 2-decimal float64 prices (thirds)
void FillSynthetic(std::uint8_t* p, std::size_t n) {
  std::uint64_t x = 20261018;
  auto rnd = [&]() { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return x; };
  const std::size_t third = n / 3 / 8 * 8;
  std::int64_t key = 0;
  for (std::size_t i = 0; i + 8 <= third; i += 8) { key += (std::int64_t)(rnd() & 3); std::memcpy(p + i, &key, 8); }
  for (std::size_t i = third; i + 4 <= 2 * third; i += 4) { std::int32_t c = (std::int32_t)(rnd() % 100 < 50 ? rnd() % 2 : rnd() % 7); std::memcpy(p + i, &c, 4); }
  for (std::size_t i = 2 * third; i + 8 <= n; i += 8) { double v = (double)(90000 + rnd() % 10410000) / 100.0; std::memcpy(p + i, &v, 8); }
}

// "content mode" of the reference (apps/demo_app.cc:144-247, 297-330): a Parquet / Feather file is read into an
// arrow::Table and serialised to Arrow IPC stream bytes -- straight into the pinned-host pool, so the
// compressor reads them in place; anything else is taken as raw bytes.
bool EndsWith(const std::string& s, const std::string& suffix) {
  return s.size() >= suffix.size() && s.compare(s.size() - suffix.size(), suffix.size(), suffix) == 0;
}

arrow::Result<std::shared_ptr<arrow::Buffer>> SerializeTable(const std::shared_ptr<arrow::Table>& table, arrow::MemoryPool* pool) {
  const auto options = arrow::ipc::IpcWriteOptions::Defaults();
  ARROW_ASSIGN_OR_RAISE(auto sink, arrow::io::BufferOutputStream::Create(1 << 20, pool));
  ARROW_ASSIGN_OR_RAISE(auto writer, arrow::ipc::MakeStreamWriter(sink, table->schema(), options));
  ARROW_RETURN_NOT_OK(writer->WriteTable(*table));
  ARROW_RETURN_NOT_OK(writer->Close());
  return sink->Finish();
}

arrow::Result<std::shared_ptr<arrow::Buffer>> ReadFile(const std::string& path, arrow::MemoryPool* pool, bool* is_table) {
  ARROW_ASSIGN_OR_RAISE(auto file, arrow::io::MemoryMappedFile::Open(path, arrow::io::FileMode::READ));
  std::shared_ptr<arrow::Table> table;
  *is_table = true;
  if (EndsWith(path, ".parquet")) {
    ARROW_ASSIGN_OR_RAISE(auto reader, parquet::arrow::OpenFile(file, arrow::default_memory_pool()));
    ARROW_RETURN_NOT_OK(reader->ReadTable(&table));
  } else if (EndsWith(path, ".feather") || EndsWith(path, ".arrow")) {
    ARROW_ASSIGN_OR_RAISE(auto reader, arrow::ipc::feather::Reader::Open(file, arrow::ipc::IpcReadOptions::Defaults()));
    ARROW_RETURN_NOT_OK(reader->Read(&table));
  } else {
    *is_table = false;
    ARROW_ASSIGN_OR_RAISE(auto size, file->GetSize());
    ARROW_ASSIGN_OR_RAISE(auto buffer, arrow::AllocateBuffer(size, pool));
    ARROW_ASSIGN_OR_RAISE(auto got, file->ReadAt(0, size, buffer->mutable_data()));
    if (got != size) return arrow::Status::IOError("Unable to read ", size, " bytes from file");
    return std::shared_ptr<arrow::Buffer>(std::move(buffer));
  }
  std::printf("table: %lld rows x %d columns\n", (long long)table->num_rows(), table->num_columns());
  return SerializeTable(table, pool);
}

// DeserializeTable of the reference (apps/demo_app.cc:225-243): the round-tripped bytes must still be an IPC stream
arrow::Result<std::shared_ptr<arrow::Table>> DeserializeTable(const std::shared_ptr<arrow::Buffer>& buffer) {
  arrow::io::BufferReader reader(buffer);
  ARROW_ASSIGN_OR_RAISE(auto batches, arrow::ipc::RecordBatchStreamReader::Open(&reader));
  return batches->ToTable();
}

bool SameBytes(const std::uint8_t* a, const std::uint8_t* b, std::size_t n, bool on_device) {
  if (!on_device) return std::memcmp(a, b, n) == 0;
  std::vector<std::uint8_t> ha(n), hb(n);
  bitar_mem_copy(ha.data(), a, n);
  bitar_mem_copy(hb.data(), b, n);
  return ha == hb;
}

}  // namespace

int main(int argc, char** argv) {
  std::size_t bytes = 64u << 20;
  std::uint32_t seg = 59460, qps_per_device = 2, sgl = 1;
  std::string file, mode = "both";
  bool on_device = false, bytes_given = false;
  for (int i = 1; i < argc; ++i) {
    const std::string a = argv[i];
    auto next = [&]() { return i + 1 < argc ? std::string(argv[++i]) : std::string(); };
    if (a == "--bytes") { bytes = std::stoull(next()); bytes_given = true; }
    else if (a == "--seg") seg = (std::uint32_t)std::stoul(next());
    else if (a == "--qps") qps_per_device = (std::uint32_t)std::stoul(next());
    else if (a == "--file") file = next();
    else if (a == "--mode") mode = next();
    else if (a == "--device") on_device = true;
    else if (a == "--sgl") sgl = (std::uint32_t)std::stoul(next());
    else { std::fprintf(stderr, "usage: %s [--file F | --bytes N] [--seg S] [--qps Q] [--mode sync|async|both] [--device] [--sgl K]\n", argv[0]); return 2; }
  }

  auto* driver = bitar::CompressDriver<bitar::Class_CUDA>::Instance();
  auto ids_r = driver->ListAvailableDeviceIds();
  CHECK_OK(ids_r.status());
  const auto ids = *ids_r;
  auto devs_r = driver->GetDevices(ids, (std::uint32_t)ids.size() * qps_per_device);
  CHECK_OK(devs_r.status());
  auto devices = std::move(*devs_r);

  auto* pool = bitar::GetMemoryPool(bitar::MemoryPoolBackend::CudaPinnedHost);
  auto allocate = [&](std::int64_t size, int device) -> arrow::Result<std::unique_ptr<arrow::ResizableBuffer>> {
    if (on_device) return bitar::AllocateDeviceBuffer(size, device);
    return arrow::AllocateResizableBuffer(size, pool);
  };
  std::shared_ptr<arrow::Buffer> input;
  bool is_table = false;
  if (!file.empty()) {
    auto host_r = ReadFile(file, pool, &is_table);   // pinned host memory, device-accessible in place
    CHECK_OK(host_r.status());
    if (bytes_given && (std::int64_t)bytes < (*host_r)->size()) *host_r = arrow::SliceBuffer(*host_r, 0, (std::int64_t)bytes);
    bytes = (std::size_t)(*host_r)->size();
    if (on_device) {
      auto in_r = allocate((std::int64_t)bytes, ids[0]);
      CHECK_OK(in_r.status());
      input = std::move(*in_r);
      bitar_mem_copy(const_cast<std::uint8_t*>(input->data()), (*host_r)->data(), bytes);
    } else {
      input = *host_r;
    }
  } else {
    std::vector<std::uint8_t> host(bytes);
    FillSynthetic(host.data(), bytes);
    auto in_r = allocate((std::int64_t)bytes, ids[0]);
    CHECK_OK(in_r.status());
    input = std::move(*in_r);
    bitar_mem_copy(const_cast<std::uint8_t*>(input->data()), host.data(), bytes);
  }

  const std::size_t n_chunks = (bytes + seg - 1) / seg;
  std::size_t total_qps = 0;
  for (auto& d : devices) total_qps += d->num_qps();
  for (auto& d : devices) {   // app_common.cc:87-100: memzones = ceil((chunks + nQP) / nDev)
    auto cfg = std::make_unique<bitar::CudaConfiguration>();
    cfg->set_decompressed_seg_size(seg);
    cfg->set_max_sgl_segs((std::uint16_t)sgl);   // K segments chained into one stream (src/include/config.h:90-96)
    cfg->set_max_preallocate_memzones((std::uint32_t)std::max<std::size_t>(20, (n_chunks + total_qps) / devices.size() + total_qps));
    cfg->set_slot_memory(on_device ? bitar::SlotMemory::kDevice : bitar::SlotMemory::kPinnedHost);
    CHECK_OK(d->Initialize(std::move(cfg)));
  }
  std::printf("%zu bytes, seg %u (%zu chunks), %zu device(s), %zu queue pair(s), buffers in %s memory\n", bytes, seg, n_chunks,
              devices.size(), total_qps, on_device ? "device" : "pinned host");
  int failures = 0;

  if (mode == "sync" || mode == "both") {   // EvaluateSync, apps/demo_app.cc:487-548
    std::printf("sync (device %u, queue pair 0):\n", (unsigned)devices[0]->device_id());
    auto& dev = devices[0];
    for (int t = 0; t < kNumTests; ++t) {
      auto t0 = Clock::now();
      auto comp_r = dev->Compress(0, input);
      auto t1 = Clock::now();
      CHECK_OK(comp_r.status());
      auto compressed = std::move(*comp_r);
      std::int64_t csize = 0;
      for (auto& b : compressed) csize += b->size();
      auto out_r = allocate((std::int64_t)((n_chunks + sgl - 1) / sgl * sgl * (std::size_t)seg), dev->device_id());
      CHECK_OK(out_r.status());
      auto output = std::move(*out_r);
      auto t2 = Clock::now();
      CHECK_OK(dev->Decompress(0, compressed, output));
      auto t3 = Clock::now();
      PrintPerf("Compress", (std::int64_t)bytes, t0, t1);
      PrintPerf("Decompress", (std::int64_t)bytes, t2, t3);
      const bool ok = (std::size_t)output->size() == bytes && SameBytes(output->data(), input->data(), bytes, on_device);
      std::printf("  ratio %.3f  round trip %s\n", (double)bytes / (double)csize, ok ? "OK" : "MISMATCH");
      failures += !ok;
      if (ok && is_table && !on_device && !bytes_given && t == 0) {   // the inflated bytes are an Arrow IPC stream again
        auto table_r = DeserializeTable(std::shared_ptr<arrow::Buffer>(std::move(output)));
        std::printf("  deserialised table: %s\n", table_r.ok() ? "OK" : table_r.status().ToString().c_str());
        failures += !table_r.ok();
        if (table_r.ok()) std::printf("  %lld rows x %d columns\n", (long long)(*table_r)->num_rows(), (*table_r)->num_columns());
      }
      if (t == 0 && !on_device && sgl == 1) {
        // the same compressed segments from ordinary heap memory (as if read back from storage): Decompress()
        // stages them itself
        bitar::BufferVector heap;
        std::vector<std::shared_ptr<arrow::Buffer>> keep;
        for (auto& b : compressed) {
          auto copy_r = arrow::AllocateBuffer(b->size());
          CHECK_OK(copy_r.status());
          std::shared_ptr<arrow::Buffer> copy = std::move(*copy_r);
          if (b->size()) std::memcpy(copy->mutable_data(), b->data(), (std::size_t)b->size());
          heap.emplace_back(std::make_unique<arrow::Buffer>(copy->data(), copy->size()));
          keep.push_back(std::move(copy));
        }
        auto out2_r = allocate((std::int64_t)(heap.size() * (std::size_t)seg), dev->device_id());
        CHECK_OK(out2_r.status());
        auto output2 = std::move(*out2_r);
        auto t4 = Clock::now();
        CHECK_OK(dev->Decompress(0, heap, output2));
        auto t5 = Clock::now();
        PrintPerf("Decompress (pageable input)", (std::int64_t)bytes, t4, t5);
        const bool ok2 = (std::size_t)output2->size() == bytes && SameBytes(output2->data(), input->data(), bytes, on_device);
        std::printf("  pageable compressed input: %s\n", ok2 ? "OK" : "MISMATCH");
        failures += !ok2;
      }
      if (dev->Recycle(compressed) != compressed.size()) { std::printf("  recycle count mismatch\n"); ++failures; }
    }
  }

  if (mode == "async" || mode == "both") {   // EvaluateAsync, apps/demo_app.cc:550-693
    std::printf("async (%zu queue pairs):\n", total_qps);
    struct Part { Device* dev; const std::unique_ptr<Device>* owner; std::uint16_t qp; std::size_t off, len; };
    std::vector<Part> parts;
    const std::size_t per = ((n_chunks + total_qps - 1) / total_qps + sgl - 1) / sgl * sgl * seg;   // even split on segment (chain) boundaries
    std::size_t off = 0;
    for (auto& d : devices)
      for (std::uint16_t q = 0; q < d->num_qps() && off < bytes; ++q) {
        parts.push_back({d.get(), &d, q, off, std::min(per, bytes - off)});
        off += per;
      }
    for (int t = 0; t < kNumTests; ++t) {
      std::vector<bitar::BufferVector> results(parts.size());
      std::vector<std::shared_ptr<arrow::Buffer>> slices;
      for (auto& p : parts) slices.push_back(arrow::SliceBuffer(input, (std::int64_t)p.off, (std::int64_t)p.len));
      auto ccb = [&](std::uint8_t dev_id, std::uint16_t qp, arrow::Result<bitar::BufferVector>&& r) -> int {
        if (!r.ok()) return EXIT_FAILURE;
        for (std::size_t i = 0; i < parts.size(); ++i)
          if (parts[i].dev->device_id() == dev_id && parts[i].qp == qp) results[i] = std::move(*r);
        return bitar::kAsyncReturnOK;
      };
      using CParam = bitar::CompressParam<bitar::Class_CUDA, decltype(ccb)>;
      std::vector<std::unique_ptr<CParam>> cparams;
      auto t0 = Clock::now();
      for (std::size_t i = 0; i < parts.size(); ++i) {
        cparams.push_back(std::make_unique<CParam>(*parts[i].owner, parts[i].qp, slices[i], ccb));
        if (bitar::CompressAsync(cparams.back()) != 0) { std::printf("  CompressAsync launch failed\n"); ++failures; }
      }
      for (auto& p : cparams) failures += bitar::WaitForAsync(p) != bitar::kAsyncReturnOK;   // demo_app.cc:258-280
      auto t1 = Clock::now();
      PrintPerf("CompressAsync", (std::int64_t)bytes, t0, t1);

      std::vector<std::unique_ptr<arrow::ResizableBuffer>> outs;
      for (std::size_t i = 0; i < parts.size(); ++i) {
        auto r = allocate((std::int64_t)(((parts[i].len + seg - 1) / seg + sgl - 1) / sgl * sgl * (std::size_t)seg), parts[i].dev->device_id());
        CHECK_OK(r.status());
        outs.push_back(std::move(*r));
      }
      auto dcb = [&](std::uint8_t, std::uint16_t, const arrow::Status& st) -> int { return st.ok() ? bitar::kAsyncReturnOK : EXIT_FAILURE; };
      using DParam = bitar::DecompressParam<bitar::Class_CUDA, decltype(dcb)>;
      std::vector<std::unique_ptr<DParam>> dparams;
      auto t2 = Clock::now();
      for (std::size_t i = 0; i < parts.size(); ++i) {
        dparams.push_back(std::make_unique<DParam>(*parts[i].owner, parts[i].qp, results[i], outs[i], dcb));
        if (bitar::DecompressAsync(dparams.back()) != 0) { std::printf("  DecompressAsync launch failed\n"); ++failures; }
      }
      for (auto& p : dparams) failures += bitar::WaitForAsync(p) != bitar::kAsyncReturnOK;
      auto t3 = Clock::now();
      PrintPerf("DecompressAsync", (std::int64_t)bytes, t2, t3);
      bool ok = true;
      for (std::size_t i = 0; i < parts.size(); ++i) {   // per-part memcmp, demo_app.cc:671-686
        ok = ok && (std::size_t)outs[i]->size() == parts[i].len &&
             SameBytes(outs[i]->data(), input->data() + parts[i].off, parts[i].len, on_device);
        if (parts[i].dev->Recycle(results[i]) != results[i].size()) ok = false;
      }
      std::printf("  round trip %s\n", ok ? "OK" : "MISMATCH");
      failures += !ok;
    }
  }
  devices.clear();
  input.reset();
  std::printf("%s (kernel launches: %llu, pool allocations tracked: %zu)\n", failures ? "FAILED" : "PASSED",
              (unsigned long long)bitar_kernel_launches(), bitar::CudaAllocatorTracker::Instance()->count());
  return failures ? EXIT_FAILURE : EXIT_SUCCESS;
}
