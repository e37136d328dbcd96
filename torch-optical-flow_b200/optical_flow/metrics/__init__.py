from optical_flow.metrics.epe import AverageEndPointError, end_point_error  # noqa: F401
