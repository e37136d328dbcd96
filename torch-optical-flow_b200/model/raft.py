"""RAFT's convex flow upsampling (reference methods/raft/model/raft.py:73-85) on the K4b kernel.

Only the hot-path static methods of the reference's `RAFT` class are provided; the network itself
(encoders, GRU, training loop) is out of scope and keeps running from the reference."""
from typing import Tuple

import torch
from torch import Tensor

import ofb200
from model.utils import coords_grid
from optical_flow.operator.operator import _check_f32, _run


def upsample_flow(flow: Tensor, mask: Tensor) -> Tensor:
    """Upsample flow field [H/8, W/8, 2] -> [H, W, 2] using convex combination (reference raft.py:73-85).

    flow (N, 2, h, w), mask (N, 576, h, w) -> (N, 2, 8h, 8w)."""
    n, _, h, w = flow.shape
    if flow.shape[1] != 2 or tuple(mask.shape) != (n, 576, h, w):
        raise RuntimeError(f"upsample_flow: expected flow (N,2,h,w) and mask (N,576,h,w), got "
                           f"{tuple(flow.shape)} and {tuple(mask.shape)}")
    _check_f32(flow, mask)

    def run(flow_d: Tensor, mask_d: Tensor):
        flow_d, mask_d = flow_d.contiguous(), mask_d.contiguous()
        out = torch.empty((n, 2, 8 * h, 8 * w), dtype=torch.float32, device=flow_d.device)
        rc = ofb200.load().ofb_convex_upsample_f32(
            ofb200.ptr(flow_d), ofb200.ptr(mask_d), ofb200.ptr(out), n, h, w, ofb200.stream_ptr()
        )
        ofb200.check(rc, "ofb_convex_upsample_f32")
        return out

    return _run(run, "upsample_flow", flow, mask)


class RAFT:
    """Namespace for the static hot-path methods of the reference's RAFT LightningModule."""

    upsample_flow = staticmethod(upsample_flow)

    @staticmethod
    def initialize_flow(img: Tensor) -> Tuple[Tensor, Tensor]:
        """flow = coords1 - coords0 on the 1/8 grid (reference raft.py:64-71)."""
        n, c, h, w = img.shape
        coords0 = coords_grid(n, h // 8, w // 8).to(img.device)
        coords1 = coords_grid(n, h // 8, w // 8).to(img.device)
        return coords0, coords1
