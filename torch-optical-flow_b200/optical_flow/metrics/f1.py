"""Outlier ratio (F1) with the reference's interface (reference optical_flow/metrics/f1.py), computed by the
K4c streaming-reduction kernel in outlier mode.  Same state handling as AverageEndPointError: the two
metric states live in one 16-byte device buffer, `sync()` all-reduces it (dist_reduce_fx="sum",
reference f1.py:30-31)."""
from typing import Optional

import torch
from torch import Tensor

import ofb200
from optical_flow.metrics.epe import _prep


class OutlierRatio:
    """Ratio of pixels whose end-point error exceeds `abs_threshold` *and* whose relative error exceeds
    `rel_threshold` (reference f1.py:10-51).

    dim            flow-component dimension of `pred` / `target` (only 1 is supported by the kernel)
    abs_threshold  a pixel can only be an outlier when its end-point error is above this many pixels ...
    rel_threshold  ... and above this fraction of the ground-truth flow magnitude (KITTI: 3 px and 5 %)
    """

    def __init__(self, dim: int = 1, abs_threshold: float = 3.0, rel_threshold: float = 0.05) -> None:
        self.dim = dim
        self.abs_threshold = abs_threshold
        self.rel_threshold = rel_threshold
        self._acc: Optional[Tensor] = None   # double[2] on the device: (sum_outliers, total)

    def update(self, pred: Tensor, target: Tensor, valid: Optional[Tensor] = None) -> None:
        pred_d, target_d = _prep(pred, target, self.dim)
        b, _, h, w = pred_d.shape
        if self._acc is None:
            self._acc = torch.zeros(2, dtype=torch.float64, device=pred_d.device)
        if valid is not None:
            valid = ofb200.to_device(valid).detach()
            if valid.numel() != b * h * w:
                raise RuntimeError("valid must have B*H*W elements")
            valid = valid.reshape(b, h, w).to(torch.float32).contiguous()
        with torch.cuda.device(pred_d.device):
            rc = ofb200.load().ofb_outlier_reduce_f32(
                ofb200.ptr(pred_d), ofb200.ptr(target_d), ofb200.ptr(valid), ofb200.ptr(self._acc), b, h, w,
                float(self.abs_threshold), float(self.rel_threshold), ofb200.stream_ptr(),
            )
        ofb200.check(rc, "ofb_outlier_reduce_f32")

    __call__ = update

    @property
    def sum_outliers(self) -> Tensor:
        return self._acc[0].to(torch.float32) if self._acc is not None else torch.tensor(0.0)

    @property
    def total(self) -> Tensor:
        return self._acc[1].to(torch.int64) if self._acc is not None else torch.tensor(0)

    def sync(self, group=None) -> None:
        import torch.distributed as dist

        if self._acc is not None and dist.is_available() and dist.is_initialized():
            dist.all_reduce(self._acc, op=dist.ReduceOp.SUM, group=group)

    def compute(self) -> Tensor:
        if self._acc is None:
            return torch.tensor(float("nan"))
        return (self._acc[0] / self._acc[1]).to(torch.float32)

    def reset(self) -> None:
        if self._acc is not None:
            self._acc.zero_()
