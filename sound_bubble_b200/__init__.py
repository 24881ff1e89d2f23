"""sound_bubble_b200 — B200-native (sm_100a) forward pass of the Sound Bubble separator behind the reference's
``src/models/*/net.py::Net`` API.  See DESIGN.md / INTEGRATION.md."""
import os as _os

# The pipelined sessions keep 16-32 streams busy; with the CUDA default of 8 hardware work queues streams alias onto the
# same queue and serialise (measured: 381k -> 800k frames/s at G = 8, depth 32; profiles/r02_group_sweep_tcp*.txt).
# The variable is read when the CUDA context is created, so it only takes effect when this package is imported before
# the first CUDA call of the process; a value set by the user wins.
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

from ._abi import SoundBubbleError  # noqa: E402,F401
from .packing import ModelConfig  # noqa: E402,F401

__all__ = ["SoundBubbleError", "ModelConfig", "Net", "NetOptim", "StreamingSession", "PipelinedSession"]


def __getattr__(name):          # lazy: importing the package must not need torch.cuda or the built library
    if name == "Net":
        from .tfgridnet_realtime_clean_dis_embd3.net import Net
        return Net
    if name == "NetOptim":
        from .tfgridnet_realtime_clean_optim.net import Net
        return Net
    if name == "StreamingSession":
        from .streaming import StreamingSession
        return StreamingSession
    if name == "PipelinedSession":
        from .streaming import PipelinedSession
        return PipelinedSession
    raise AttributeError(name)
