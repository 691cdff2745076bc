from .segmenter import HeartSoundSegmenter

__all__ = ["HeartSoundSegmenter"]
