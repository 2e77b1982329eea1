"""qcs_b200 -- B200-native state-vector engine behind the QCS C89 API.

The product is two native libraries (qcs_b200/lib/libqcs.so, libqcs_cuda.so)
built from qcs_b200/csrc; this package only builds and binds them.
"""
from .circuit import Circuit, QcsError, set_default  # noqa: F401
from . import _ffi  # noqa: F401

__all__ = ["Circuit", "QcsError", "set_default"]
