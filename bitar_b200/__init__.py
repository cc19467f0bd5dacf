"""bitar_b200 -- B200-native block DEFLATE engine behind bitar's API (host-side Python view).

The product is the C-ABI shared library ``bitar_b200/csrc/libbitar_cuda.so`` (include/bitar_cuda.h)
and the C++ facade in ``bitar_b200/host``; this package only binds the C-ABI for tests and bench.
"""
__version__ = "0.1.0"
