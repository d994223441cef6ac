"""speechcatcher_b200: B200-native (sm_100a) multi-stream streaming decode path for speechcatcher.

Public surface mirrors the reference: `Speech2TextStreaming` (drop-in facade) plus the batched
`StreamGroup` engine underneath.  The CUDA extension (libscb200.so) is required; there is no CPU path.
"""
__all__ = ["Speech2TextStreaming", "StreamGroup", "create_streaming_interface"]


def __getattr__(name):
    if name in ("Speech2TextStreaming", "create_streaming_interface"):
        from . import speech2text_streaming as m
        return getattr(m, name)
    if name == "StreamGroup":
        from .stream_group import StreamGroup
        return StreamGroup
    raise AttributeError(name)
