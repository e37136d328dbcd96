"""`optical_flow.metrics.f1` of the reference (optical_flow/metrics/f1.py) on the K4c kernel in outlier mode
(`ofb200.ops.f1`); with a reference checkout and torchmetrics present, the reference's own Metric subclass with its
`update` on the kernel (see optical_flow/metrics/epe.py)."""
from ofb200.ops.f1 import OutlierRatio  # noqa: F401
from ofb200.overlay import reference_metric_class as _ref_class

_cls = _ref_class(__name__, __file__, "f1.py", "OutlierRatio")
if _cls is not None:
    OutlierRatio = _cls
