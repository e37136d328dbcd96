"""Drop-in `optical_flow` package: the hot-path subset of awaelchli/torch-optical-flow's call
surface (reference optical_flow/__init__.py:2,4), backed by sm_100a kernels through libofb200."""
from optical_flow.operator.operator import denormalize, integrate, normalize, resize, scale, warp  # noqa: F401
from optical_flow.metrics.epe import AverageEndPointError  # noqa: F401
