from .resample import Resample
from .synchrosqueeze import FSST

__all__ = ["Resample", "FSST"]
