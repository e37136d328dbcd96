"""Patch an already-imported reference checkout of awaelchli/torch-optical-flow in place.

Two ways to put the B200 kernels behind the reference's call surface:

1. *path overlay* -- `torch-optical-flow_b200/` ahead of the reference on `sys.path`: the `optical_flow` and `model`
   packages of this repository answer the hot-path names and fall through to the reference for the rest
   (`optical_flow/__init__.py`, `model/__init__.py`);
2. *in-place patch* (this module) -- the caller imported the reference as usual; `patch_reference()` rebinds, inside the
   reference's own modules, exactly the functions / classes of SURVEY.md section 8a:

       optical_flow.operator.operator.{warp, warp_grid, scale, resize, normalize, denormalize, integrate}   operator.py:8-165
       optical_flow.{warp, scale, resize, normalize, denormalize, integrate}                                __init__.py:2
       model.corr.CorrBlock                                                                                corr.py:37-87
       model.utils.{bilinear_sampler, upflow8}                                                             utils.py:64-91
       model.raft.{CorrBlock, upflow8, sequence_loss}, model.raft.RAFT.upsample_flow                       raft.py:6-9,73-85,231-260
       optical_flow.metrics.epe.{end_point_error, AverageEndPointError.update}                             epe.py:25-61
       optical_flow.metrics.f1.OutlierRatio.update                                                         f1.py:33-48

   The metric CLASSES stay the reference's `torchmetrics.Metric` subclasses -- states, `compute`, `dist_reduce_fx="sum"`
   synchronisation and Lightning logging are untouched; only their `update` runs the K4c kernel and adds its
   (sum, count) into the `sum_epe` / `sum_outliers` and `total` states.  Everything else of the reference (io,
   visualisation, encoders, update block, LightningModule, data, CLI) is left alone.

`unpatch_reference()` restores every binding.  Nothing here imports the reference: it only touches modules that are
already in `sys.modules` (or are passed in).
"""
import sys
from typing import Dict, List, Optional, Tuple

import torch

from ofb200.ops import corr as _corr
from ofb200.ops import epe as _epe
from ofb200.ops import f1 as _f1
from ofb200.ops import operator as _op
from ofb200.ops import raft_ops as _raft
from ofb200.ops import sampling as _samp

_OPERATOR_NAMES = ("warp", "warp_grid", "scale", "resize", "normalize", "denormalize", "integrate")
_saved: List[Tuple[object, str, object]] = []


def _is_ours(obj) -> bool:
    mod = getattr(obj, "__module__", "") or ""
    return mod.startswith("ofb200.")


def _bind(owner, name: str, value) -> bool:
    if owner is None or not hasattr(owner, name):
        return False
    old = owner.__dict__.get(name, None) if isinstance(owner, type) else getattr(owner, name)
    if old is value or _is_ours(getattr(old, "__func__", old)):
        return False
    _saved.append((owner, name, old))
    setattr(owner, name, value)
    return True


def _epe_update(self, pred, target, valid=None):
    """AverageEndPointError.update (reference epe.py:25-35) on the K4c kernel; states stay the reference's."""
    pred_d, target_d = _epe._prep(pred, target, self.dim)
    acc = torch.zeros(2, dtype=torch.float64, device=pred_d.device)
    _epe._accumulate(acc, pred_d, target_d, valid)
    self.sum_epe += acc[0].to(device=self.sum_epe.device, dtype=self.sum_epe.dtype)
    self.total += acc[1].to(device=self.total.device, dtype=self.total.dtype)


def _outlier_update(self, pred, target, valid=None):
    """OutlierRatio.update (reference f1.py:33-48) on the K4c kernel in outlier mode."""
    pred_d, target_d = _epe._prep(pred, target, self.dim)
    acc = torch.zeros(2, dtype=torch.float64, device=pred_d.device)
    _f1._accumulate_outliers(acc, pred_d, target_d, valid, self.abs_threshold, self.rel_threshold)
    self.sum_outliers += acc[0].to(device=self.sum_outliers.device, dtype=self.sum_outliers.dtype)
    self.total += acc[1].to(device=self.total.device, dtype=self.total.dtype)


def reference_metric_class(shim_name: str, shim_file: str, filename: str, cls_name: str):
    """Path-overlay helper for optical_flow/metrics/{epe,f1}.py: find the reference's file of the same name in the
    other directories of the `optical_flow.metrics` package path, execute it unmodified as `<shim>_reference`, rebind
    its hot-path pieces (`update`, `end_point_error`) to the kernels and return its Metric class.  None when there is
    no reference on the path or its import fails (torchmetrics missing): the caller keeps the self-contained class."""
    import importlib.util
    import os

    pkg = sys.modules.get(shim_name.rsplit(".", 1)[0])
    here = os.path.dirname(os.path.abspath(shim_file))
    for d in list(getattr(pkg, "__path__", [])):
        cand = os.path.join(d, filename)
        if os.path.abspath(d) == here or not os.path.isfile(cand):
            continue
        alias = shim_name + "_reference"
        spec = importlib.util.spec_from_file_location(alias, cand)
        mod = importlib.util.module_from_spec(spec)
        sys.modules[alias] = mod
        try:
            spec.loader.exec_module(mod)
        except ImportError:
            del sys.modules[alias]
            return None
        cls = getattr(mod, cls_name)
        if cls_name == "AverageEndPointError":
            cls.reference_update, cls.update = cls.update, _epe_update
            mod.end_point_error = _epe.end_point_error
        else:
            cls.reference_update, cls.update = cls.update, _outlier_update
        return cls
    return None


def patch_reference(modules: Optional[Dict[str, object]] = None, metrics: bool = True) -> List[str]:
    """Rebind the hot-path names of the imported reference to the B200 implementations.

    modules: name -> module, defaults to `sys.modules` (the reference's `optical_flow*` / `model*` entries).
    metrics: also patch the `update` methods of the reference's metric classes.
    Returns the list of "module.name" bindings that were changed (empty if the reference is not imported)."""
    mods = modules if modules is not None else sys.modules
    done: List[str] = []

    def bind(mod_name: str, name: str, value, attr_of: Optional[str] = None):
        mod = mods.get(mod_name)
        owner = getattr(mod, attr_of, None) if (mod is not None and attr_of) else mod
        if _bind(owner, name, value):
            done.append(f"{mod_name}.{attr_of + '.' if attr_of else ''}{name}")

    for name in _OPERATOR_NAMES:
        bind("optical_flow.operator.operator", name, getattr(_op, name))
        if name != "warp_grid":
            bind("optical_flow", name, getattr(_op, name))
    bind("model.corr", "CorrBlock", _corr.CorrBlock)
    bind("model", "CorrBlock", _corr.CorrBlock)
    for name in ("bilinear_sampler", "upflow8"):
        bind("model.utils", name, getattr(_samp, name))
    bind("model.corr", "bilinear_sampler", _samp.bilinear_sampler)
    bind("model.raft", "CorrBlock", _corr.CorrBlock)
    bind("model.raft", "upflow8", _samp.upflow8)
    bind("model.raft", "sequence_loss", _raft.sequence_loss)
    bind("model.raft", "upsample_flow", staticmethod(_raft.upsample_flow), attr_of="RAFT")
    if metrics:
        bind("optical_flow.metrics.epe", "end_point_error", _epe.end_point_error)
        bind("optical_flow.metrics.epe", "update", _epe_update, attr_of="AverageEndPointError")
        bind("optical_flow.metrics.f1", "update", _outlier_update, attr_of="OutlierRatio")
    return done


def unpatch_reference() -> int:
    """Undo every binding `patch_reference` made.  Returns how many were restored."""
    n = len(_saved)
    while _saved:
        owner, name, old = _saved.pop()
        setattr(owner, name, old)
    return n
