"""unreal_b200: B200-native (sm_100a) rollout-and-target hot path of kvas7andy/unreal.

Host code is Python/PyTorch; every kernel lives in libunreal_b200.so behind the C ABI of
include/unreal_b200.h and is reached through ctypes (unreal_b200._lib).  Importing the
package requires the built library; there is no CPU fallback.
"""
from . import _lib  # noqa: F401  (raises when libunreal_b200.so is missing)
from . import kernels  # noqa: F401

__all__ = ["_lib", "kernels"]
