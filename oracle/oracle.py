"""numpy/ctypes front-end of oracle.c -- TEST INFRASTRUCTURE ONLY (see oracle.c header).

Every function mirrors one reference entry point (paths relative to /root/reference):

    warp              optical_flow/operator/operator.py:8-56
    scale/normalize/  optical_flow/operator/operator.py:59-82,117-146
    denormalize
    resize            optical_flow/operator/operator.py:85-114
    integrate         optical_flow/operator/operator.py:149-165
    corr_volume       methods/raft/model/corr.py:79-87      (numpy sgemm replaces torch.matmul)
    corr_pyramid      methods/raft/model/corr.py:38-54
    corr_lookup       methods/raft/model/corr.py:56-77 + methods/raft/model/utils.py:64-80
    bilinear_sampler  methods/raft/model/utils.py:64-80
    coords_grid       methods/raft/model/utils.py:83-86
    upflow8           methods/raft/model/utils.py:89-91
    upsample_flow     methods/raft/model/raft.py:73-85
    end_point_error / epe_sum_count   optical_flow/metrics/epe.py:25-61
    outlier_sum_count  optical_flow/metrics/f1.py:33-48
    sequence_loss      methods/raft/model/raft.py:231-260

All arrays are C-contiguous float32 numpy arrays in the reference's layouts (NCHW).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_c_f = ctypes.POINTER(ctypes.c_float)
_c_u8 = ctypes.POINTER(ctypes.c_uint8)
_c_i32 = ctypes.POINTER(ctypes.c_int32)


def _cpu_has_v3():
    try:
        with open("/proc/cpuinfo") as fh:
            for line in fh:
                if line.startswith("flags"):
                    flags = set(line.split(":", 1)[1].split())
                    return {"avx2", "fma", "bmi2"} <= flags
    except OSError:
        pass
    return False


def build(force=False):
    """Compile liboracle_*.so with the committed Makefile (gcc only)."""
    if force:
        subprocess.run(["make", "-C", _HERE, "clean"], check=True, capture_output=True)
    subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)


def _load():
    name = "liboracle_v3.so" if _cpu_has_v3() else "liboracle_base.so"
    path = os.path.join(_HERE, name)
    if not os.path.exists(path):
        build()
    return ctypes.CDLL(path)


_lib = _load()
for _n in (
    "orc_linspace_f32", "orc_grid_sample_f32", "orc_warp_f32", "orc_resize_bilinear_f32",
    "orc_avg_pool2_f32", "orc_corr_lookup_f32", "orc_convex_upsample_f32", "orc_epe_f32",
    "orc_epe_map_f32", "orc_round_bf16_f32", "orc_outlier_f32", "orc_sequence_loss_f32",
):
    getattr(_lib, _n).restype = None
_lib.orc_linspace_f32.argtypes = [ctypes.c_float, ctypes.c_float, ctypes.c_int, _c_f]
_lib.orc_grid_sample_f32.argtypes = [_c_f, _c_f, _c_f] + [ctypes.c_int] * 9
_lib.orc_warp_f32.argtypes = [_c_f, _c_f, _c_f, _c_u8] + [ctypes.c_int] * 7
_lib.orc_resize_bilinear_f32.argtypes = [_c_f, _c_f] + [ctypes.c_int] * 7 + [ctypes.c_float] * 2
_lib.orc_avg_pool2_f32.argtypes = [_c_f, _c_f, ctypes.c_int64, ctypes.c_int, ctypes.c_int]
_lib.orc_corr_lookup_f32.argtypes = [
    ctypes.POINTER(_c_f), ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int),
    ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int), _c_f, _c_f, _c_i32, _c_u8,
] + [ctypes.c_int] * 5
_lib.orc_convex_upsample_f32.argtypes = [_c_f, _c_f, _c_f] + [ctypes.c_int] * 3
_lib.orc_epe_f32.argtypes = [_c_f, _c_f, _c_f, ctypes.POINTER(ctypes.c_double),
                             ctypes.POINTER(ctypes.c_int64)] + [ctypes.c_int] * 3
_lib.orc_epe_map_f32.argtypes = [_c_f, _c_f, _c_f] + [ctypes.c_int] * 3
_lib.orc_outlier_f32.argtypes = [_c_f, _c_f, _c_f, ctypes.POINTER(ctypes.c_double),
                                 ctypes.POINTER(ctypes.c_int64)] + [ctypes.c_int] * 3 + [ctypes.c_float] * 2
_lib.orc_sequence_loss_f32.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, _c_f, _c_f,
                                       ctypes.POINTER(ctypes.c_double)] + [ctypes.c_int] * 3 + [ctypes.c_double, ctypes.c_float]
_lib.orc_round_bf16_f32.argtypes = [_c_f, _c_f, ctypes.c_int64]

_MODES = {"bilinear": 0, "nearest": 1}
_PADS = {"zeros": 0, "border": 1, "reflection": 2}


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return a.ctypes.data_as(_c_f)


def linspace(start, end, n):
    out = np.empty(n, np.float32)
    _lib.orc_linspace_f32(start, end, n, _p(out))
    return out


def grid_sample(img, grid, mode="bilinear", padding_mode="zeros", align_corners=False):
    img, grid = _f32(img), _f32(grid)
    n, c, h, w = img.shape
    _, ho, wo, _ = grid.shape
    out = np.empty((n, c, ho, wo), np.float32)
    _lib.orc_grid_sample_f32(_p(img), _p(grid), _p(out), n, c, h, w, ho, wo,
                             _MODES[mode], _PADS[padding_mode], int(align_corners))
    return out


def warp(frame, flow, mode="bilinear", padding_mode="border", align_corners=False, return_mask=False):
    frame, flow = _f32(frame), _f32(flow)
    b, c, h, w = frame.shape
    assert flow.shape == (b, 2, h, w)
    out = np.empty_like(frame)
    valid = np.empty((b, h, w), np.uint8) if return_mask else None
    _lib.orc_warp_f32(_p(frame), _p(flow), _p(out),
                      valid.ctypes.data_as(_c_u8) if return_mask else None,
                      b, c, h, w, _MODES[mode], _PADS[padding_mode], int(align_corners))
    return (out, valid) if return_mask else out


def scale(flow, factor=1.0):
    flow = np.asarray(flow)
    assert flow.shape[1] == 2
    if isinstance(factor, (float, int)):
        factor = (factor, factor)
    assert len(factor) == 2
    f = np.array(factor, dtype=flow.dtype).reshape(1, 2, 1, 1)
    return flow * f


def normalize(flow):
    assert flow.shape[1] == 2
    h, w = flow.shape[-2:]
    return scale(flow, (2.0 / max(w - 1, 1), 2.0 / max(h - 1, 1)))


def denormalize(flow):
    assert flow.shape[1] == 2
    h, w = flow.shape[-2:]
    return scale(flow, (max(w - 1, 1) / 2, max(h - 1, 1) / 2))


def _resize_raw(x, size, align_corners, mul_x=1.0, mul_y=1.0):
    x = _f32(x)
    n, c, h, w = x.shape
    out = np.empty((n, c, size[0], size[1]), np.float32)
    _lib.orc_resize_bilinear_f32(_p(x), _p(out), n, c, h, w, size[0], size[1],
                                 int(align_corners), mul_x, mul_y)
    return out


def resize(flow, size=None, scale_factor=None):
    flow = _f32(flow)
    assert flow.shape[1] == 2
    assert flow.ndim == 4
    _, _, h, w = flow.shape
    if scale_factor:
        size = (round(h * scale_factor), round(w * scale_factor))
    sy = size[0] / h
    sx = size[1] / w
    return _resize_raw(flow, size, False, sx, sy)


def upflow8(flow):
    flow = _f32(flow)
    return _resize_raw(flow, (8 * flow.shape[2], 8 * flow.shape[3]), True, 8.0, 8.0)


def integrate(*flows):
    assert len(flows) >= 2
    total = _f32(flows[-1])
    for flow in reversed(flows[:-1]):
        assert flow.shape == total.shape, "All flows must have the same size."
        total = _f32(flow) + warp(total, flow)
    return total


def round_bf16(x):
    x = _f32(x)
    out = np.empty_like(x)
    _lib.orc_round_bf16_f32(_p(x), _p(out), x.size)
    return out


def corr_volume(fmap1, fmap2):
    """(B,C,h,w) x2 -> (B,h,w,1,h,w); numpy sgemm stands in for torch.matmul."""
    fmap1, fmap2 = _f32(fmap1), _f32(fmap2)
    b, c, h, w = fmap1.shape
    a = fmap1.reshape(b, c, h * w)
    bb = fmap2.reshape(b, c, h * w)
    corr = np.matmul(a.transpose(0, 2, 1), bb)
    corr = corr / np.sqrt(np.float32(c))
    return corr.reshape(b, h, w, 1, h, w).astype(np.float32, copy=False)


def avg_pool2(x):
    x = _f32(x)
    n, one, h, w = x.shape
    assert one == 1
    out = np.empty((n, 1, h // 2, w // 2), np.float32)
    _lib.orc_avg_pool2_f32(_p(x), _p(out), n, h, w)
    return out


def corr_pyramid(fmap1, fmap2, num_levels=4):
    corr = corr_volume(fmap1, fmap2)
    b, h1, w1, dim, h2, w2 = corr.shape
    corr = corr.reshape(b * h1 * w1, dim, h2, w2)
    pyr = [corr]
    for _ in range(num_levels - 1):
        corr = avg_pool2(corr)
        pyr.append(corr)
    return pyr


def corr_lookup(pyramid, coords, radius=4, return_index=False, shared_slice=False):
    """pyramid: list of (B*h*w, 1, h_l, w_l) float32; coords (B,2,h,w) -> (B, L*(2r+1)^2, h, w).

    With return_index=True also returns idx (Q,L,2,2r+1) int32 and valid (Q,L,(2r+1)^2) u8.
    shared_slice=True: the levels are (1, 1, h_l, w_l) and every query reads the same slice (query stride 0) --
    for checks of the floor indices / validity masks of many queries, which do not depend on the volume.
    """
    coords = _f32(coords)
    b, two, h, w = coords.shape
    assert two == 2
    pyr = [_f32(p) for p in pyramid]
    L = len(pyr)
    d = 2 * radius + 1
    q = b * h * w
    ptrs = (_c_f * L)(*[_p(p) for p in pyr])
    qs = (ctypes.c_int64 * L)(*[0 if shared_slice else p.shape[2] * p.shape[3] for p in pyr])
    assert all(p.shape[0] == (1 if shared_slice else q) for p in pyr)
    rp = (ctypes.c_int * L)(*[p.shape[3] for p in pyr])
    lh = (ctypes.c_int * L)(*[p.shape[2] for p in pyr])
    lw = (ctypes.c_int * L)(*[p.shape[3] for p in pyr])
    out = np.empty((b, L * d * d, h, w), np.float32)
    idx = np.empty((q, L, 2, d), np.int32) if return_index else None
    valid = np.empty((q, L, d * d), np.uint8) if return_index else None
    _lib.orc_corr_lookup_f32(ptrs, qs, rp, lh, lw, _p(coords), _p(out),
                             idx.ctypes.data_as(_c_i32) if return_index else None,
                             valid.ctypes.data_as(_c_u8) if return_index else None,
                             b, h, w, L, radius)
    return (out, idx, valid) if return_index else out


def bilinear_sampler(img, coords, mask=False):
    """utils.py:64-80 -- pixel coords (N,Ho,Wo,2) -> grid_sample(align_corners=True, zeros)."""
    img, coords = _f32(img), _f32(coords)
    H, W = img.shape[-2:]
    xg = np.float32(2) * coords[..., 0:1] / np.float32(W - 1) - np.float32(1)
    yg = np.float32(2) * coords[..., 1:2] / np.float32(H - 1) - np.float32(1)
    grid = np.concatenate([xg, yg], axis=-1).astype(np.float32)
    out = grid_sample(img, grid, "bilinear", "zeros", True)
    if mask:
        m = (xg > -1) & (yg > -1) & (xg < 1) & (yg < 1)
        return out, m.astype(np.float32)
    return out


def coords_grid(batch, ht, wd):
    ys, xs = np.meshgrid(np.arange(ht), np.arange(wd), indexing="ij")
    coords = np.stack([xs, ys], axis=0).astype(np.float32)
    return np.repeat(coords[None], batch, axis=0)


def upsample_flow(flow, mask):
    flow, mask = _f32(flow), _f32(mask)
    n, two, h, w = flow.shape
    assert two == 2 and mask.shape == (n, 576, h, w)
    out = np.empty((n, 2, 8 * h, 8 * w), np.float32)
    _lib.orc_convex_upsample_f32(_p(flow), _p(mask), _p(out), n, h, w)
    return out


def epe_sum_count(pred, target, valid=None):
    pred, target = _f32(pred), _f32(target)
    b, two, h, w = pred.shape
    assert two == 2 and target.shape == pred.shape
    v = _f32(valid) if valid is not None else None
    s = ctypes.c_double(0.0)
    c = ctypes.c_int64(0)
    _lib.orc_epe_f32(_p(pred), _p(target), _p(v) if v is not None else None,
                     ctypes.byref(s), ctypes.byref(c), b, h, w)
    return s.value, c.value


def outlier_sum_count(pred, target, valid=None, abs_threshold=3.0, rel_threshold=0.05):
    """optical_flow/metrics/f1.py:33-48 -> (number of outliers, number of selected pixels)."""
    pred, target = _f32(pred), _f32(target)
    b, two, h, w = pred.shape
    assert two == 2 and target.shape == pred.shape
    v = _f32(valid) if valid is not None else None
    s = ctypes.c_double(0.0)
    c = ctypes.c_int64(0)
    _lib.orc_outlier_f32(_p(pred), _p(target), _p(v) if v is not None else None, ctypes.byref(s), ctypes.byref(c),
                         b, h, w, abs_threshold, rel_threshold)
    return s.value, c.value


def sequence_loss(flow_preds, flow_gt, valid, gamma=0.8, max_flow=400.0):
    """methods/raft/model/raft.py:231-260 -> (loss, {"1px", "3px", "5px"}, extras {"epe_sum", "kept"})."""
    preds = [_f32(p) for p in flow_preds]
    gt, v = _f32(flow_gt), _f32(valid)
    b, two, h, w = gt.shape
    assert two == 2 and v.shape == (b, h, w) and all(p.shape == gt.shape for p in preds)
    arr = (ctypes.c_void_p * len(preds))(*[p.ctypes.data for p in preds])
    out = (ctypes.c_double * 6)()
    _lib.orc_sequence_loss_f32(arr, len(preds), _p(gt), _p(v), out, b, h, w, float(gamma), float(max_flow))
    kept = out[2]
    metrics = {k: (out[i] / kept if kept > 0 else float("nan")) for k, i in (("1px", 3), ("3px", 4), ("5px", 5))}
    return out[0], metrics, {"epe_sum": out[1], "kept": int(kept)}


def end_point_error(pred, target, reduce=True):
    pred, target = _f32(pred), _f32(target)
    b, _, h, w = pred.shape
    out = np.empty((b, h, w), np.float32)
    _lib.orc_epe_map_f32(_p(pred), _p(target), _p(out), b, h, w)
    return out.mean(dtype=np.float32) if reduce else out
