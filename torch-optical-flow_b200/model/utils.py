"""Sampling utilities with the reference's interface (reference methods/raft/model/utils.py)."""
from typing import List, Sequence, Tuple, Union

import torch
import torch.nn.functional as F
from torch import Tensor

import ofb200
from optical_flow.operator.operator import _check_f32, _resize_raw, _run


class InputPadder:
    """Pads images such that dimensions are divisible by 8 (reference utils.py:38-61).

    Caller-side helper, not on the accelerated path: plain replicate padding."""

    def __init__(self, dims: Sequence[int], mode: str = "sintel") -> None:
        self.ht, self.wd = dims[-2:]
        pad_ht = (((self.ht // 8) + 1) * 8 - self.ht) % 8
        pad_wd = (((self.wd // 8) + 1) * 8 - self.wd) % 8
        if mode == "sintel":
            self._pad = [pad_wd // 2, pad_wd - pad_wd // 2, pad_ht // 2, pad_ht - pad_ht // 2]
        else:
            self._pad = [pad_wd // 2, pad_wd - pad_wd // 2, 0, pad_ht]

    def pad(self, *inputs: Tensor) -> List[Tensor]:
        return [F.pad(x, self._pad, mode="replicate") for x in inputs]

    def unpad(self, x: Tensor) -> Tensor:
        ht, wd = x.shape[-2:]
        c = [self._pad[2], ht - self._pad[3], self._pad[0], wd - self._pad[1]]
        return x[..., c[0] : c[1], c[2] : c[3]]


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
    coords = torch.stack((xs, ys), dim=0).float()
    return coords[None].repeat(batch, 1, 1, 1)


def upflow8(flow: Tensor, mode: str = "bilinear") -> Tensor:
    """8 * bilinear x8 upsampling with align_corners=True (reference utils.py:89-91)."""
    if mode != "bilinear":
        raise NotImplementedError(f"upflow8: mode={mode!r} has no B200 kernel (bilinear only)")
    _check_f32(flow)
    new_size = (8 * flow.shape[2], 8 * flow.shape[3])
    return _resize_raw(flow, new_size, True, 8.0, 8.0, "upflow8")
