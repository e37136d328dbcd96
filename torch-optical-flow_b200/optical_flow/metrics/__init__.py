"""`optical_flow.metrics` of the reference (optical_flow/metrics/__init__.py:1-2) on the K4c reduction kernels."""
import pkgutil

__path__ = pkgutil.extend_path(__path__, __name__)

from optical_flow.metrics.epe import AverageEndPointError, end_point_error  # noqa: E402,F401
from optical_flow.metrics.f1 import OutlierRatio  # noqa: E402,F401
