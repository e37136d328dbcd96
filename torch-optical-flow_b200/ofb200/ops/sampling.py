"""Sampling utilities with the reference's interface (reference methods/raft/model/utils.py)."""
from typing import List, Sequence, Tuple, Union

import torch
import torch.nn.functional as F
from torch import Tensor

import ofb200
from ofb200.ops.operator import _check_f32, _resize_raw, _run


class InputPadder:
    """Replicate-pads images up to the next multiple of 8 in both dimensions and crops results back (behaviour of
    reference utils.py:38-61).  Caller-side helper, not on the accelerated path.

    mode "sintel": the extra rows / columns are split evenly (the odd one goes to the bottom / right);
    any other mode: columns split evenly, all extra rows at the bottom (KITTI)."""

    MULTIPLE = 8

    def __init__(self, dims: Sequence[int], mode: str = "sintel") -> None:
        height, width = int(dims[-2]), int(dims[-1])
        extra_h = -height % self.MULTIPLE
        extra_w = -width % self.MULTIPLE
        self.left, self.right = extra_w // 2, extra_w - extra_w // 2
        if mode == "sintel":
            self.top, self.bottom = extra_h // 2, extra_h - extra_h // 2
        else:
            self.top, self.bottom = 0, extra_h
        self.ht, self.wd = height, width

    def pad(self, *inputs: Tensor) -> List[Tensor]:
        amounts = (self.left, self.right, self.top, self.bottom)
        return [F.pad(image, amounts, mode="replicate") for image in inputs]

    def unpad(self, x: Tensor) -> Tensor:
        rows, cols = x.shape[-2] - self.top - self.bottom, x.shape[-1] - self.left - self.right
        return x.narrow(-2, self.top, rows).narrow(-1, self.left, cols)


def bilinear_sampler(
    img: Tensor, coords: Tensor, mode: str = "bilinear", mask: bool = False
) -> Union[Tensor, Tuple[Tensor, Tensor]]:
    """grid_sample with pixel coordinates (reference utils.py:64-80).

    img (N, C, H, W), coords (N, Ho, Wo, 2) -> (N, C, Ho, Wo) [, mask (N, Ho, Wo, 1) float].
    The reference accepts `mode` but never forwards it (utils.py:74); same here."""
    n, c, h, w = img.shape
    _, ho, wo, _ = coords.shape
    _check_f32(img, coords)

    def run(img_d: Tensor, coords_d: Tensor):
        img_d, coords_d = img_d.contiguous(), coords_d.contiguous()
        out = torch.empty((n, c, ho, wo), dtype=torch.float32, device=img_d.device)
        m = torch.empty((n, ho, wo, 1), dtype=torch.float32, device=img_d.device) if mask else None
        rc = ofb200.load().ofb_bilinear_sampler_f32(
            ofb200.ptr(img_d), ofb200.ptr(coords_d), ofb200.ptr(out), ofb200.ptr(m), n, c, h, w, ho, wo,
            ofb200.stream_ptr(),
        )
        ofb200.check(rc, "ofb_bilinear_sampler_f32")
        return (out, m) if mask else out

    return _run(run, "bilinear_sampler", img, coords)


def coords_grid(batch: int, ht: int, wd: int) -> Tensor:
    """(batch, 2, ht, wd) fp32 grid, channel 0 = x (column), channel 1 = y (row) (reference utils.py:83-86)."""
    ys, xs = torch.meshgrid(torch.arange(ht), torch.arange(wd), indexing="ij")
    grid = torch.stack((xs, ys), dim=0).to(torch.float32)
    return grid.unsqueeze(0).expand(batch, -1, -1, -1).contiguous()


def upflow8(flow: Tensor, mode: str = "bilinear") -> Tensor:
    """8 * bilinear x8 upsampling with align_corners=True (reference utils.py:89-91)."""
    if mode != "bilinear":
        raise NotImplementedError(f"upflow8: mode={mode!r} has no B200 kernel (bilinear only)")
    _check_f32(flow)
    h, w = flow.shape[-2:]
    return _resize_raw(flow, (8 * h, 8 * w), True, 8.0, 8.0, "upflow8")
