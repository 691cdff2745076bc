"""Signal transforms of the B200 build.

``FSST`` -- Fourier synchrosqueezed transform features, computed by the CUDA kernels of ``libhssb.so``
(drop-in for reference ``hss/transforms/synchrosqueeze.py``); ``Resample`` -- Fourier-method resampling with
``torch.fft`` (reference ``hss/transforms/resample.py``).
"""
from . import resample as _resample
from . import synchrosqueeze as _synchrosqueeze

FSST = _synchrosqueeze.FSST
Resample = _resample.Resample

__all__ = ("FSST", "Resample")
