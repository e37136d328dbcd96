from optical_flow.metrics.epe import AverageEndPointError, end_point_error  # noqa: F401
from optical_flow.metrics.f1 import OutlierRatio  # noqa: F401
