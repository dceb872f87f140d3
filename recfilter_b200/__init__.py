"""recfilter_b200 -- B200-native tiled recursive-filter engine.

Python side = thin ctypes binding of the C ABI declared in ``include/recfilter_b200.h``
(plumbing for tests, bench and the multi-GPU driver).  The product is the shared
library ``librecfilter_b200.so`` (hand-written sm_100a kernels + launch planner) and
the C++ operator surface in ``include/recfilter.h``.

There is no CPU fallback: if the CUDA library is missing or no device is present,
every compute call raises.
"""
from .capi import (  # noqa: F401
    Plan,
    MultiGpuPlan,
    Exchange,
    Scan,
    RecFilterError,
    lib,
    lib_path,
    device_count,
    DTYPES,
    exported_symbols,
)
from .filters import (  # noqa: F401
    gaussian_weights,
    integral_image_coeff,
    overlap_feedback_coeff,
    gaussian_box_filter,
)

__all__ = [
    "Plan", "MultiGpuPlan", "Exchange", "Scan", "RecFilterError", "lib", "lib_path", "device_count", "DTYPES",
    "exported_symbols", "gaussian_weights", "integral_image_coeff",
    "overlap_feedback_coeff", "gaussian_box_filter",
]
