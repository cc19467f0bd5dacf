"""bitar_b200 -- B200-native block DEFLATE engine behind bitar's API (host-side Python view).

The product is the C-ABI shared library ``bitar_b200/csrc/libbitar_cuda.so`` (include/bitar_cuda.h)
and the C++ facade in ``bitar_b200/host``; this package only binds the C-ABI for tests and bench.
"""
import os as _os

# A device with 8 queue pairs runs up to 16 streams (staged calls add a copy-back stream per queue pair).  The CUDA
# driver maps streams onto 8 hardware queues unless told otherwise; streams that share one serialise (measured:
# decompress from pinned host memory 30 instead of 44 GB/s at 4+ queue pairs).  Read at CUDA initialisation, so it
# is set here, before the first CUDA call of the process; libbitar_cuda.so does the same when it is loaded.
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

__version__ = "0.1.0"
