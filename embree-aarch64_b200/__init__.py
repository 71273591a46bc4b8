"""b200-rayquery: a B200-native ray-query engine behind Embree's rtcore C API.

The product is `lib/libembree3.so` (CUDA kernels + C-ABI, built from `csrc/`); this package is the
thin Python host mirror used by tests and the benchmark: `rtcore` (ctypes binding of the C ABI),
`fixtures` (synthetic scenes / ray streams of the BASELINE configurations) and `build`.
"""
from . import rtcore, fixtures  # noqa: F401
from .rtcore import RTCore, PRODUCT_LIB  # noqa: F401
