"""`model.raft` of the reference (methods/raft/model/raft.py).

When a reference checkout's `methods/raft` is on `sys.path` behind this package, its raft.py is executed UNMODIFIED
as `model._reference_raft`: its `from model.corr import CorrBlock` / `from model.utils import ...` lines
(raft.py:6-9) bind to the B200 implementations, `update.py` / `extractor.py` are its own files, and the two hot-path
pieces defined inside the module itself -- `RAFT.upsample_flow` (raft.py:73-85) and `sequence_loss` (raft.py:231-260) --
are rebound to the kernels.  Otherwise `RAFT` is a namespace with the static hot-path methods."""
import importlib.util
import os
import sys

import model as _pkg
from ofb200.ops.raft_ops import RAFT as _HotPathRAFT
from ofb200.ops.raft_ops import sequence_loss, upsample_flow  # noqa: F401


def _load_reference_raft():
    here = os.path.dirname(os.path.abspath(__file__))
    for d in list(_pkg.__path__):
        cand = os.path.join(d, "raft.py")
        if os.path.abspath(d) == here or not os.path.isfile(cand):
            continue
        if not all(os.path.isfile(os.path.join(d, f)) for f in ("update.py", "extractor.py")):
            continue
        spec = importlib.util.spec_from_file_location("model._reference_raft", cand)
        mod = importlib.util.module_from_spec(spec)
        sys.modules[spec.name] = mod
        try:
            spec.loader.exec_module(mod)
        except ImportError:
            del sys.modules[spec.name]          # e.g. pytorch_lightning missing: stay with the namespace class
            return None
        return mod
    return None


_ref = _load_reference_raft()
if _ref is not None:
    RAFT = _ref.RAFT
    RAFT.reference_upsample_flow = staticmethod(RAFT.__dict__["upsample_flow"].__func__)
    RAFT.upsample_flow = staticmethod(upsample_flow)
    _ref.reference_sequence_loss = _ref.sequence_loss
    _ref.sequence_loss = sequence_loss
else:
    RAFT = _HotPathRAFT
