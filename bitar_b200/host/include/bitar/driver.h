// driver.h -- CompressDriver<Class_CUDA>: probe of the CUDA devices and creation of device objects.
// Mirrors /root/reference/src/include/driver.h:40-66 and src/driver.cc:100-223.
#pragma once
#include <arrow/result.h>

#include <cstdint>
#include <memory>
#include <vector>

#include "config.h"
#include "device.h"

namespace bitar {

template <typename Class>
class CompressDriver {
 public:
  CompressDriver(const CompressDriver&) = delete;
  CompressDriver& operator=(const CompressDriver&) = delete;

  static CompressDriver<Class>* Instance();

  /// \brief Get devices with the selected ids.  \p num_workers queue pairs (CUDA streams; the reference
  /// uses its worker lcores) are spread over the devices as evenly as possible, each device at least
  /// one, the first num_workers % n devices one more (src/driver.cc:100-158).  0 = one per device.
  arrow::Result<std::vector<std::unique_ptr<CompressDevice<Class>>>> GetDevices(
      const std::vector<std::uint8_t>& device_ids, std::uint32_t num_workers = 0);

  /// \brief All device ids managed by this driver on the current machine (src/driver.cc:173-190).
  arrow::Result<std::vector<std::uint8_t>> ListAvailableDeviceIds();

 private:
  CompressDriver() = default;
  ~CompressDriver() = default;
};

namespace internal {
/// \brief The queue-pair distribution rule alone (host logic, src/driver.cc:103-117,152-155).
std::vector<std::uint16_t> DistributeWorkers(std::uint32_t num_workers, std::size_t num_devices);
}  // namespace internal

}  // namespace bitar
