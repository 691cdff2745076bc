from .synchrosqueeze import FSST

__all__ = ["FSST"]
