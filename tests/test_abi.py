"""The C-ABI boundary without a GPU: libhssb.so loads, exports every symbol include/hssb.h
declares, and its argument checking follows the stated error convention (no compute calls)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib(built):
    from hss import _lib

    return _lib.lib()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "hssb.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hssb_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    from hss import _lib

    names = declared_symbols()
    assert len(names) >= 14
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/hssb.h but not exported"
    assert sorted(_lib.EXPORTED_SYMBOLS) == names
    assert lib.hssb_version() == 1


def test_header_cites_reference_lines():
    text = open(os.path.join(ROOT, "include", "hssb.h")).read()
    for cite in ("synchrosqueeze.py:48", "segmenter.py:70-87", "hss/moments/__init__.py", "main.py:36-62"):
        assert cite in text


def test_argument_errors_do_not_need_a_gpu(lib):
    from hss import _lib

    assert lib.hssb_fsst_stft(None, 1, 1, None, None, 128, None, None, None) == -1          # HSSB_E_NULL
    assert b"null" in lib.hssb_last_error()
    buf = (ctypes.c_float * 8)()
    p = ctypes.addressof(buf)
    assert lib.hssb_fsst_stft(p, 1, 8, p, p, 3, p, p, None) == -3                            # HSSB_E_NWIN (4 .. 1024)
    assert lib.hssb_fsst_stft(p, -1, 8, p, p, 128, p, p, None) == -2                         # HSSB_E_SHAPE
    assert lib.hssb_fsst_reassign(p, p, 1, 8, 128, 1000.0, 30, 20, p, None, None) == -4      # HSSB_E_BAND
    assert lib.hssb_fsst_finish(p, None, 1, 8, 22, 7, p, None) == -5                         # HSSB_E_MODE
    assert lib.hssb_fsst_forward(p, 1, 8, p, p, 128, 1000.0, 4, 25, 2, p, None, 0, None) == -6   # HSSB_E_WORKSPACE
    assert lib.hssb_fsst_stft(p, 0, 0, p, p, 128, p, p, None) == 0                           # empty input: no-op
    with pytest.raises(ValueError):
        _lib.check(-4, "x")
    with pytest.raises(RuntimeError):
        _lib.check(700, "x")
    assert lib.hssb_fsst_workspace_bytes(2, 2000, 128, 4, 25, 2) >= 2 * 2 * 65 * 2000 * 8
    assert lib.hssb_fsst_stats_words(2, 2000) == 2 * 16 * 6 + 4
    assert lib.hssb_lstm_forward(None, p, 1, 8, 44, p, p, p, None, None, 0, None) == -1      # HSSB_E_NULL (SURVEY 8b names)
    assert lib.hssb_lstm_workspace_bytes(None, 1, 8) == 0
    assert lib.hssb_auroc_hist(p, p, 8, 5000, p, None) == -2                                 # nbins outside [2, 4096]
    assert lib.hssb_lstm_train_forward(None, p, p, p, p, 1, 8, 240, p, p, p, p, None) == -1
    # tensor-core training entry points and the in-place re-pack: null handles / pointers, bad layer, bad shapes
    assert lib.hssb_lstm_train_forward_tc(None, 0, p, 1, 8, p, p, p, p, p, p, p, None, 0, None) == -1
    assert lib.hssb_lstm_train_backward_tc(None, p, None, None, None, p, p, p, p, p, None, None, 1, 8, p, p, None, 0, None) == -1
    assert lib.hssb_lstm_train_backward_tc(p, None, p, None, None, p, p, p, p, p, None, None, 1, 8, p, p, None, 0, None) == -1   # half a pair
    assert lib.hssb_lstm_train_backward_tc(p, p, None, None, None, p, p, p, p, p, None, None, -1, 8, p, p, None, 0, None) == -2
    assert lib.hssb_lstm_train_backward_tc_workspace_bytes() >= 2 * 8 * 2 * 256 * 128 * 2
    assert lib.hssb_model_update(None, None, None) == -1
    assert lib.hssb_model_split_bytes(None, 1, 8) == 0
    assert lib.hssb_model_split_input(None, p, 1, 8, p, 0, None) == -1
    assert lib.hssb_model_forward_split(None, p, None, 1, 8, p, p, p, None, None, 0, None) == -1
    assert lib.hssb_model_side_gate(None, 1, 8, None, None) == -1
    assert lib.hssb_model_uses_tensor_cores(None) == 0
    assert lib.hssb_split_tf32(p, -1, p, p, None) == -2
    assert lib.hssb_split_tf32(p, 0, p, p, None) == 0
    assert lib.hssb_split_tf32(None, 8, p, p, None) == -1


def test_no_cuda_means_loud_failure(lib):
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from hss.transforms import FSST
    from hss.model.segmenter import HeartSoundSegmenter

    f = FSST(1000, window=np.kaiser(128, 0.5), truncate_freq=(25, 200), stack=True)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        f(torch.zeros(2000))
    m = HeartSoundSegmenter(input_size=44, batch_size=1).eval()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 10, 44))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.prepare(torch.zeros(1, 10, 44))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "heart-sounds-segmentation_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src, f"{f} mentions the oracle"
