"""Drop-in `optical_flow` package: the hot-path names of awaelchli/torch-optical-flow's call surface (reference
optical_flow/__init__.py:1-5) bound to the sm_100a kernels of libofb200 (`ofb200.ops`).

It is an OVERLAY, not a replacement.  Put this directory ahead of a reference checkout on `sys.path` /
`PYTHONPATH` and everything the hot path does not cover keeps coming from the reference: the package path is
extended over every other `optical_flow` directory on `sys.path`, so `optical_flow.io`, `optical_flow.visualization`
and the top-level names `read`, `write`, `flow2rgb`, `colorwheel` resolve there (lazily, on first use).  Without a
reference on the path those four names raise an ImportError that says so; the hot-path names never need it.
The other way round -- the reference imported first, then patched in place -- is `ofb200.overlay.patch_reference()`.
"""
import importlib
import pkgutil

from ofb200.ops.epe import AverageEndPointError  # noqa: F401
from ofb200.ops.f1 import OutlierRatio  # noqa: F401
from ofb200.ops.operator import denormalize, integrate, normalize, resize, scale, warp  # noqa: F401

__path__ = pkgutil.extend_path(__path__, __name__)

# reference optical_flow/__init__.py:1,3 -- host-side codecs and visualisation, out of the accelerated path
_FALLTHROUGH = {
    "read": "optical_flow.io.read_write",
    "write": "optical_flow.io.read_write",
    "flow2rgb": "optical_flow.visualization.flow2rgb",
    "colorwheel": "optical_flow.visualization.flow2rgb",
}


def __getattr__(name):
    if name in _FALLTHROUGH:
        global __path__
        __path__ = pkgutil.extend_path(__path__, __name__)      # a reference appended to sys.path after this import
        try:
            mod = importlib.import_module(_FALLTHROUGH[name])
        except ModuleNotFoundError as e:
            raise ImportError(
                f"optical_flow.{name} is not part of the B200 hot path; it falls through to a reference checkout of "
                f"torch-optical-flow on sys.path, and none provides {_FALLTHROUGH[name]} ({e})") from e
        value = getattr(mod, name)
        globals()[name] = value
        return value
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
