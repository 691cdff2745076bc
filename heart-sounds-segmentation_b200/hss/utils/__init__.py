from .ingest import load_recording_csv, load_recordings_csv, recording_to_frames, scan_recordings_csv, stream_recordings
from .preprocess import frame_batch, frame_signal

__all__ = ["frame_signal", "frame_batch", "load_recording_csv", "load_recordings_csv", "scan_recordings_csv", "recording_to_frames",
           "stream_recordings"]
