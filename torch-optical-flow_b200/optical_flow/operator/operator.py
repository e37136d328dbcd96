"""`optical_flow.operator.operator` of the reference (optical_flow/operator/operator.py:8-165), bound to the K1 / K4a
kernels: the implementation lives in `ofb200.ops.operator`."""
from ofb200.ops.operator import (  # noqa: F401
    denormalize, integrate, normalize, resize, scale, warp, warp_grid,
)
