// type_fwd.h -- mirrors /root/reference/src/include/type_fwd.h:30
#pragma once
#include <arrow/buffer.h>

#include <memory>
#include <vector>

namespace bitar {

using BufferVector = std::vector<std::unique_ptr<arrow::Buffer>>;

}  // namespace bitar
