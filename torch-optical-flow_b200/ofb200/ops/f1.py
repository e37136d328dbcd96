"""Outlier ratio (F1) with the reference's interface (reference optical_flow/metrics/f1.py), computed by the
K4c streaming-reduction kernel in outlier mode.  Same state handling as AverageEndPointError
(`ofb200.ops.epe.SumCountMetric`; dist_reduce_fx="sum", reference f1.py:30-31)."""
from typing import Optional

import torch
from torch import Tensor

import ofb200
from ofb200.ops.epe import SumCountMetric, _prep_valid


def _accumulate_outliers(acc: Tensor, pred: Tensor, target: Tensor, valid: Optional[Tensor], abs_threshold: float,
                         rel_threshold: float) -> None:
    b, _, h, w = pred.shape
    valid = _prep_valid(valid, b, h, w)
    with torch.cuda.device(pred.device):
        rc = ofb200.load().ofb_outlier_reduce_f32(
            ofb200.ptr(pred), ofb200.ptr(target), ofb200.ptr(valid), ofb200.ptr(acc), b, h, w,
            float(abs_threshold), float(rel_threshold), ofb200.stream_ptr(),
        )
    ofb200.check(rc, "ofb_outlier_reduce_f32")


class OutlierRatio(SumCountMetric):
    """Ratio of pixels whose end-point error exceeds `abs_threshold` *and* whose relative error exceeds
    `rel_threshold` (reference f1.py:10-51).

    dim            flow-component dimension of `pred` / `target` (only 1 is supported by the kernel)
    abs_threshold  a pixel can only be an outlier when its end-point error is above this many pixels ...
    rel_threshold  ... and above this fraction of the ground-truth flow magnitude (KITTI: 3 px and 5 %)
    """

    def __init__(self, dim: int = 1, abs_threshold: float = 3.0, rel_threshold: float = 0.05) -> None:
        super().__init__()
        self.dim = dim
        self.abs_threshold = abs_threshold
        self.rel_threshold = rel_threshold

    def _accumulate(self, acc: Tensor, pred: Tensor, target: Tensor, valid: Optional[Tensor]) -> None:
        _accumulate_outliers(acc, pred, target, valid, self.abs_threshold, self.rel_threshold)

    @property
    def sum_outliers(self) -> Tensor:
        return self._acc[0].to(torch.float32) if self._acc is not None else torch.tensor(0.0)

    @property
    def total(self) -> Tensor:
        return self._acc[1].to(torch.int64) if self._acc is not None else torch.tensor(0)
