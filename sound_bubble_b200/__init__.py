"""sound_bubble_b200 — B200-native (sm_100a) forward pass of the Sound Bubble separator behind the reference's
``src/models/*/net.py::Net`` API.  See DESIGN.md / INTEGRATION.md."""
from ._abi import SoundBubbleError  # noqa: F401
from .packing import ModelConfig  # noqa: F401

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
