"""`model.corr` of the reference (methods/raft/model/corr.py); implementation: `ofb200.ops.corr` (K2 + K3)."""
from ofb200.ops.corr import CorrBlock, prepare_operands  # noqa: F401
