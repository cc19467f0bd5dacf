// driver.cc -- CompressDriver<Class_CUDA>: CUDA device enumeration behind the C-ABI shim.
// Mirrors /root/reference/src/driver.cc:100-223 (probe, id validation, even spread of the workers).
#include "bitar/driver.h"

#include <algorithm>

namespace bitar {

namespace internal {
std::vector<std::uint16_t> DistributeWorkers(std::uint32_t num_workers, std::size_t num_devices) {
  std::vector<std::uint16_t> out(num_devices, 0);
  if (num_devices == 0) return out;
  const auto min_per = static_cast<std::uint16_t>(num_workers / num_devices);   // src/driver.cc:105-117
  const auto extra = num_workers % num_devices;
  for (std::size_t i = 0; i < num_devices; ++i) out[i] = static_cast<std::uint16_t>(min_per + (i < extra ? 1 : 0));
  return out;
}
}  // namespace internal

template <typename Class>
CompressDriver<Class>* CompressDriver<Class>::Instance() {   // src/driver.cc:162-166
  static CompressDriver<Class> instance;
  return &instance;
}

template <typename Class>
arrow::Result<std::vector<std::uint8_t>> CompressDriver<Class>::ListAvailableDeviceIds() {
  const int n = bitar_cuda_device_count();
  if (n <= 0) return arrow::Status::Invalid("No compress device is available with driver name: cuda");   // src/driver.cc:184-187
  std::vector<std::uint8_t> ids(static_cast<std::size_t>(std::min(n, 255)));
  for (std::size_t i = 0; i < ids.size(); ++i) ids[i] = static_cast<std::uint8_t>(i);
  return ids;
}

template <typename Class>
arrow::Result<std::vector<std::unique_ptr<CompressDevice<Class>>>> CompressDriver<Class>::GetDevices(
    const std::vector<std::uint8_t>& device_ids, std::uint32_t num_workers) {
  ARROW_ASSIGN_OR_RAISE(auto available, ListAvailableDeviceIds());
  for (auto id : device_ids)   // src/driver.cc:55-73
    if (std::find(available.begin(), available.end(), id) == available.end())
      return arrow::Status::Invalid("Device id ", static_cast<unsigned>(id), " is not available");
  if (device_ids.empty()) return std::vector<std::unique_ptr<CompressDevice<Class>>>{};
  if (num_workers == 0) num_workers = static_cast<std::uint32_t>(device_ids.size());
  if (device_ids.size() > num_workers)   // src/driver.cc:204-210
    return arrow::Status::Invalid("The number of devices to set up (", device_ids.size(),
                                  ") is greater than the number of available workers (", num_workers, ").");
  const auto per_device = internal::DistributeWorkers(num_workers, device_ids.size());
  std::vector<std::unique_ptr<CompressDevice<Class>>> devices;
  devices.reserve(device_ids.size());
  for (std::size_t i = 0; i < device_ids.size(); ++i) {
    ARROW_ASSIGN_OR_RAISE(auto* dev, DeviceManager::Instance()->Create(device_ids[i], per_device[i]));
    devices.emplace_back(dev);
  }
  return devices;
}

template class CompressDriver<Class_CUDA>;

}  // namespace bitar
