"""Drop-in for the hot-path part of the reference's `model` package (methods/raft/model): the correlation block, the
sampling / upsampling utilities and RAFT's convex upsampling, bound to libofb200 (`ofb200.ops`).

An OVERLAY like `optical_flow`: the package path is extended over every other `model` directory on `sys.path`, so with
a reference checkout's `methods/raft` behind this directory `model.update` and `model.extractor` are the reference's
files, and `model.RAFT` is the reference's own LightningModule (methods/raft/model/raft.py), executed unmodified
against this package's `model.corr` / `model.utils` and with `RAFT.upsample_flow` rebound to the K4b kernel.  Without a
reference on the path `model.RAFT` is a namespace holding the static hot-path methods only."""
import pkgutil

__path__ = pkgutil.extend_path(__path__, __name__)

from model.corr import CorrBlock  # noqa: E402,F401
from model.raft import RAFT, sequence_loss, upsample_flow  # noqa: E402,F401
from model.utils import InputPadder, bilinear_sampler, coords_grid, upflow8  # noqa: E402,F401
