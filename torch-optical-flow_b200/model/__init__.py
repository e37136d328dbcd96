"""Drop-in for the hot-path part of the reference's `model` package (methods/raft/model): the
correlation block, the sampling / upsampling utilities and RAFT's convex upsampling.  The encoders,
the update block and the LightningModule stay in the reference (out of scope)."""
from model.corr import CorrBlock  # noqa: F401
from model.raft import RAFT, sequence_loss, upsample_flow  # noqa: F401
from model.utils import InputPadder, bilinear_sampler, coords_grid, upflow8  # noqa: F401
