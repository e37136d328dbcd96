"""`optical_flow.metrics.epe` of the reference (optical_flow/metrics/epe.py) on the K4c kernel (`ofb200.ops.epe`).

With a reference checkout behind this package on `sys.path` AND torchmetrics importable, `AverageEndPointError` is the
reference's own `torchmetrics.Metric` subclass -- states, `compute`, cross-rank sync and Lightning logging unchanged --
whose `update` runs the kernel (`ofb200.overlay`); otherwise it is the self-contained class of `ofb200.ops.epe`."""
from ofb200.ops.epe import AverageEndPointError, end_point_error  # noqa: F401
from ofb200.overlay import reference_metric_class as _ref_class

_cls = _ref_class(__name__, __file__, "epe.py", "AverageEndPointError")
if _cls is not None:
    AverageEndPointError = _cls
