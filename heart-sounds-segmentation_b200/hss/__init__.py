"""B200-native drop-in for the hot path of ``hss`` (alvgaona/heart-sounds-segmentation).

Mirrors the reference package layout for the three subsystems on the FSST -> BiLSTM path
(reference ``hss/__init__.py:1-9``): ``hss.transforms.FSST``, ``hss.moments`` and
``hss.model.segmenter.HeartSoundSegmenter`` (plus the framing helper of ``hss.utils.preprocess``, the caller side of FSST).  All arithmetic runs in ``libhssb.so`` (hand-written
sm_100a CUDA behind the C-ABI of ``include/hssb.h``); there is no CPU fallback.  Datasets, file
walking, training utilities and plotting are out of scope.
"""
from . import moments, transforms, model, utils  # noqa: F401

__all__ = ["moments", "transforms", "model", "utils"]
