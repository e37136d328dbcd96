"""End-point error with the reference's interface (reference optical_flow/metrics/epe.py), computed
by the K4c streaming-reduction kernel.  torchmetrics is not required: the two metric states
(`sum_epe`, `total`) live in one 16-byte device buffer and `sync()` sums it across ranks with a
single all-reduce -- the semantics of `dist_reduce_fx="sum"` (reference epe.py:22-23)."""
from typing import Optional

import torch
from torch import Tensor

import ofb200


def _prep(pred: Tensor, target: Tensor, dim: int):
    if dim != 1 or pred.dim() != 4 or pred.shape[1] != 2:
        raise NotImplementedError("end-point error kernel expects (B, 2, H, W) flows and dim=1")
    if pred.shape != target.shape:
        raise RuntimeError(f"pred {tuple(pred.shape)} and target {tuple(target.shape)} differ in shape")
    for t in (pred, target):
        if t.dtype != torch.float32:
            raise NotImplementedError(f"ofb200 kernels are fp32 only, got {t.dtype}")
    return ofb200.to_device(pred).detach().contiguous(), ofb200.to_device(target).detach().contiguous()


def _accumulate(acc: Tensor, pred: Tensor, target: Tensor, valid: Optional[Tensor]) -> None:
    b, _, h, w = pred.shape
    if valid is not None:
        valid = ofb200.to_device(valid).detach()
        if valid.numel() != b * h * w:
            raise RuntimeError("valid must have B*H*W elements")
        valid = valid.reshape(b, h, w).to(torch.float32).contiguous()
    with torch.cuda.device(pred.device):
        rc = ofb200.load().ofb_epe_reduce_f32(
            ofb200.ptr(pred), ofb200.ptr(target), ofb200.ptr(valid), ofb200.ptr(acc), b, h, w, ofb200.stream_ptr()
        )
    ofb200.check(rc, "ofb_epe_reduce_f32")


class AverageEndPointError:
    """Average End-to-end Point Error (reference epe.py:8-38): streaming mean of ||pred - target||_2.

    Args:
        dim: the dimension along which to compute the end-point-error (only 1 is supported)
    """

    def __init__(self, dim: int = 1) -> None:
        self.dim = dim
        self._acc: Optional[Tensor] = None   # double[2] on the device: (sum_epe, total)

    def _state(self, device) -> Tensor:
        if self._acc is None:
            self._acc = torch.zeros(2, dtype=torch.float64, device=device)
        return self._acc

    def update(self, pred: Tensor, target: Tensor, valid: Optional[Tensor] = None) -> None:
        pred_d, target_d = _prep(pred, target, self.dim)
        _accumulate(self._state(pred_d.device), pred_d, target_d, valid)

    __call__ = update

    @property
    def sum_epe(self) -> Tensor:
        return self._acc[0].to(torch.float32) if self._acc is not None else torch.tensor(0.0)

    @property
    def total(self) -> Tensor:
        return self._acc[1].to(torch.int64) if self._acc is not None else torch.tensor(0)

    def sync(self, group=None) -> None:
        """Sum the metric states over all ranks (one 16-byte all-reduce)."""
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


def end_point_error(pred: Tensor, target: Tensor, dim: int = 1, reduce: bool = True) -> Tensor:
    """End-to-end Point Error (reference epe.py:41-61): mean EPE, or the (B, H, W) map with reduce=False."""
    on_host = not pred.is_cuda
    pred_d, target_d = _prep(pred, target, dim)
    b, _, h, w = pred_d.shape
    if reduce:
        acc = torch.zeros(2, dtype=torch.float64, device=pred_d.device)
        _accumulate(acc, pred_d, target_d, None)
        out = (acc[0] / acc[1]).to(torch.float32)
    else:
        out = torch.empty((b, h, w), dtype=torch.float32, device=pred_d.device)
        with torch.cuda.device(pred_d.device):
            rc = ofb200.load().ofb_epe_map_f32(
                ofb200.ptr(pred_d), ofb200.ptr(target_d), ofb200.ptr(out), b, h, w, ofb200.stream_ptr()
            )
        ofb200.check(rc, "ofb_epe_map_f32")
    return out.cpu() if on_host else out
