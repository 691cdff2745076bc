from .preprocess import frame_signal, frame_batch

__all__ = ["frame_signal", "frame_batch"]
