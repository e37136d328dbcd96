"""`model.utils` of the reference (methods/raft/model/utils.py); implementation: `ofb200.ops.sampling`."""
from ofb200.ops.sampling import InputPadder, bilinear_sampler, coords_grid, upflow8  # noqa: F401
