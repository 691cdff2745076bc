from .ingest import load_recording_csv, recording_to_frames
from .preprocess import frame_batch, frame_signal

__all__ = ["frame_signal", "frame_batch", "load_recording_csv", "recording_to_frames"]
