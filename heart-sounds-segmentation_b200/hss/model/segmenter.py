"""``hss.model.segmenter.HeartSoundSegmenter`` with the inference forward on sm_100a kernels.

Constructor signature, parameter names (the 18 ``state_dict`` keys), the random fixed ``h0``/``c0``
and their RNG draw order, the batch-size check and the ``forward(x[B,T,F]) -> logp[B,T,4]``
contract follow reference ``hss/model/segmenter.py:20-87``.  ``nn.LSTM`` / ``nn.Linear`` modules are
kept purely as parameter containers (so checkpoints load unchanged); the arithmetic of the eval-mode
forward runs in ``libhssb.so`` (``hssb_model_forward``).  In training mode ``forward`` is differentiable
(SURVEY.md 8f-4): see ``hss/model/_train.py``.
"""
from __future__ import annotations

import ctypes
import os

import torch
from torch import nn

from .. import _lib


class PreparedInput:
    """Features of a batch together with their fp16 operand split (``HeartSoundSegmenter.prepare``)."""

    def __init__(self, x: torch.Tensor, planes: torch.Tensor, key):
        self.x, self.planes, self.key = x, planes, key

    @property
    def shape(self):
        return self.x.shape

    def record_stream(self, stream) -> None:
        self.x.record_stream(stream)
        self.planes.record_stream(stream)


class HeartSoundSegmenter(nn.Module):
    """Two-layer bidirectional LSTM + linear head, 4 heart-sound states per time step."""

    def __init__(
        self,
        *,
        input_size: int,
        batch_size: int = 1,
        hidden_size: int = 240,
        bidirectional: bool = True,
        device: torch.device | None = None,
        dtype: torch.dtype = torch.float32,
    ) -> None:
        super().__init__()
        self.device = device if device is not None else torch.device("cpu")
        self.batch_size = batch_size
        self.bidirectional = bidirectional
        D = 2 if bidirectional else 1
        # same draw order as the reference ctor (segmenter.py:38-67): h0, c0, lstm_1, lstm_2, linear
        self.h0, self.c0 = (
            torch.randn(D, batch_size, hidden_size, device=self.device, dtype=dtype),
            torch.randn(D, batch_size, hidden_size, device=self.device, dtype=dtype),
        )
        self.lstm_1 = nn.LSTM(input_size=input_size, hidden_size=hidden_size, bidirectional=bidirectional,
                              batch_first=True, device=self.device, dtype=dtype)
        self.lstm_2 = nn.LSTM(input_size=hidden_size * 2, hidden_size=hidden_size, bidirectional=bidirectional,
                              batch_first=True, device=self.device, dtype=dtype)
        self.dropout = nn.Dropout(0.2)
        self.relu = nn.ReLU()
        self.linear = nn.Linear(in_features=hidden_size * 2, out_features=4, bias=True, device=self.device, dtype=dtype)
        self.softmax = nn.LogSoftmax(dim=2)

        self._handle = None           # hssb_model* (ctypes.c_void_p)
        self._handle_key = None
        self._state_dev: dict = {}
        self._workspace: dict = {}

    # ------------------------------------------------------------------------------------------
    def _params_in_abi_order(self):
        out = []
        for layer in (self.lstm_1, self.lstm_2):
            for suffix in ("", "_reverse"):
                out.append([getattr(layer, f"{k}_l0{suffix}") for k in ("weight_ih", "weight_hh", "bias_ih", "bias_hh")])
        return out

    def _release(self):
        if self._handle is not None:
            _lib.lib().hssb_model_destroy(self._handle)
            self._handle = None
            self._handle_key = None

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    # The packed-weights handle, the device copies of h0 / c0 and the workspaces belong to this object and this process:
    # copies and pickles (copy.deepcopy, torch.save(model), spawn-based strategies) start without them and repack lazily,
    # like the plain reference module, which can be copied and pickled freely.
    def __getstate__(self):
        state = self.__dict__.copy()
        state["_handle"], state["_handle_key"], state["_state_dev"], state["_workspace"] = None, None, {}, {}
        state.pop("_last_forward", None)
        return state

    def __setstate__(self, state):
        super().__setstate__(state)
        self._handle, self._handle_key, self._state_dev, self._workspace = None, None, {}, {}

    def _packed(self, dev: torch.device):
        """(Re)pack the parameters into HBM operands when they changed (load_state_dict, optimiser step)."""
        tensors = [t for group in self._params_in_abi_order() for t in group] + [self.linear.weight, self.linear.bias]
        key = (dev.index,) + tuple((t.data_ptr(), t._version) for t in tensors)
        if self._handle is not None and key == self._handle_key:
            return self._handle
        same_device = self._handle is not None and self._handle_key[0] == dev.index
        if not same_device:
            self._release()
        if not self.bidirectional:
            raise NotImplementedError("the B200 kernels implement the bidirectional segmenter only")
        if any(t.dtype != torch.float32 for t in tensors):
            raise NotImplementedError("the B200 kernels take float32 parameters")
        keep = [t.detach().contiguous() for t in tensors]   # host or device pointers are both accepted
        p = _lib.ModelParams()
        p.input_size = self.lstm_1.input_size
        p.hidden_size = self.lstm_1.hidden_size
        groups = self._params_in_abi_order()
        i = 0
        for layer in range(2):
            for d in range(2):
                p.w_ih[layer][d] = keep[i + 0].data_ptr()
                p.w_hh[layer][d] = keep[i + 1].data_ptr()
                p.b_ih[layer][d] = keep[i + 2].data_ptr()
                p.b_hh[layer][d] = keep[i + 3].data_ptr()
                i += 4
        del groups
        p.lin_w = keep[-2].data_ptr()
        p.lin_b = keep[-1].data_ptr()
        if same_device:
            # the parameters changed in place (optimiser step, load_state_dict): re-pack into the operands the handle owns
            self._handle_key = None
            with torch.cuda.device(dev):
                rc = _lib.lib().hssb_model_update(self._handle, ctypes.byref(p), _lib.stream_ptr())
            if rc:
                self._release()
            _lib.check(rc, "hssb_model_update")
            self._handle_key = key
            return self._handle
        handle = ctypes.c_void_p()
        with torch.cuda.device(dev):
            rc = _lib.lib().hssb_model_create(ctypes.byref(p), ctypes.byref(handle), _lib.stream_ptr())
        _lib.check(rc, "hssb_model_create")
        self._handle, self._handle_key = handle, key
        return handle

    def _initial_state(self, dev: torch.device):
        key = (dev.index, self.h0.data_ptr(), self.h0._version, self.c0.data_ptr(), self.c0._version)
        if self._state_dev.get("key") != key:
            self._state_dev = {
                "key": key,
                "h0": self.h0.detach().to(device=dev, dtype=torch.float32).contiguous(),
                "c0": self.c0.detach().to(device=dev, dtype=torch.float32).contiguous(),
            }
        return self._state_dev["h0"], self._state_dev["c0"]

    # ------------------------------------------------------------------------------------------
    def _check_input(self, x: torch.Tensor) -> None:
        if x.dim() != 3:
            raise ValueError(f"expected input of shape (batch, seq, feature), got {tuple(x.shape)}")
        if x.shape[0] != self.batch_size:
            # the reference fails inside nn.LSTM because h0/c0 are pre-shaped (segmenter.py:38-41)
            raise RuntimeError(
                f"Expected hidden[0] size {(self.h0.shape[0], x.shape[0], self.h0.shape[2])}, got {list(self.h0.shape)}"
            )
        if x.shape[2] != self.lstm_1.input_size:
            raise RuntimeError(f"input.size(-1) must be equal to input_size. Expected {self.lstm_1.input_size}, got {x.shape[2]}")

    def prepare(self, x: torch.Tensor) -> "PreparedInput":
        """Run the first kernel of the inference forward (the fp16 operand split of ``x[B, T, F]``) ahead of time, on the current
        stream: ``hssb_model_split_input``.  ``predict`` / ``forward_with_labels`` / ``forward`` accept the result in place of
        ``x`` (``hssb_model_forward_split``) with identical results; ``hss.pipeline`` uses it to take the split off the critical path."""
        if self.training:
            raise RuntimeError("prepare() belongs to the inference calls: switch the module to .eval()")
        if not x.is_cuda:
            raise RuntimeError("prepare() takes the features where the kernels run: a CUDA tensor (no CPU fallback)")
        self._check_input(x)
        dev = x.device
        with torch.cuda.device(dev):
            xd = x.detach().to(dtype=torch.float32).contiguous()
            B, T, _ = xd.shape
            handle = self._packed(dev)
            need = _lib.lib().hssb_model_split_bytes(handle, B, T) if B and T else 0
            planes = torch.empty(max(need, 1), dtype=torch.uint8, device=dev)
            if need:
                rc = _lib.lib().hssb_model_split_input(handle, xd.data_ptr(), B, T, planes.data_ptr(), planes.numel(), _lib.stream_ptr())
                _lib.check(rc, "hssb_model_split_input")
        return PreparedInput(xd, planes, self._handle_key)

    def _run(self, x, want_logp: bool, want_labels: bool):
        if self.training:
            raise RuntimeError("predict() / forward_with_labels() are inference calls: switch the module to .eval()")
        prepared = x if isinstance(x, PreparedInput) else None
        if prepared is not None:
            x = prepared.x
        self._check_input(x)
        lib = _lib.lib()
        was_cpu = not x.is_cuda
        dev = _lib.require_cuda() if was_cpu else x.device
        impl = {"auto": 0, "tc": 0, "simt": 1}[os.environ.get("HSSB_LSTM_IMPL", "auto")]
        with torch.cuda.device(dev):
            xd = x.detach().to(device=dev, dtype=torch.float32, non_blocking=True).contiguous()
            B, T, _ = xd.shape
            handle = self._packed(dev)
            h0, c0 = self._initial_state(dev)
            logp = torch.empty((B, T, 4), dtype=torch.float32, device=dev) if want_logp else None
            labels = torch.empty((B, T), dtype=torch.int32, device=dev) if want_labels else None
            if B and T:
                need = lib.hssb_model_workspace_bytes(handle, B, T)
                ws = _lib.cached_workspace(self._workspace, dev, need)
                self._last_forward = (handle, B, T, ws, dev)
                p_logp, p_labels = logp.data_ptr() if want_logp else None, labels.data_ptr() if want_labels else None
                if prepared is not None and impl == 0 and prepared.key == self._handle_key:
                    # (a split made for other weights / another device is simply not used: the planes depend on neither, the check
                    #  only guards against a PreparedInput that outlived a move of the module)
                    rc = lib.hssb_model_forward_split(handle, xd.data_ptr(), prepared.planes.data_ptr(), B, T, h0.data_ptr(), c0.data_ptr(),
                                                      p_logp, p_labels, ws.data_ptr(), ws.numel(), _lib.stream_ptr())
                else:
                    rc = lib.hssb_model_forward(handle, xd.data_ptr(), B, T, h0.data_ptr(), c0.data_ptr(), p_logp, p_labels,
                                                ws.data_ptr(), ws.numel(), impl, _lib.stream_ptr())
                _lib.check(rc, "hssb_model_forward")
        if was_cpu:
            logp = logp.cpu() if logp is not None else None
            labels = labels.cpu() if labels is not None else None
        return logp, labels

    def side_gate(self, side_stream: "torch.cuda.Stream") -> None:
        """Hold ``side_stream`` back until the layer-1 recurrence of the inference forward enqueued LAST holds its SMs
        (``hssb_model_side_gate``): work queued on it afterwards runs on the SMs the recurrences leave idle (``hss.pipeline``)."""
        last = getattr(self, "_last_forward", None)
        if last is None:
            return
        handle, B, T, ws, dev = last
        with torch.cuda.device(dev):
            rc = _lib.lib().hssb_model_side_gate(handle, B, T, ws.data_ptr(), side_stream.cuda_stream)
        _lib.check(rc, "hssb_model_side_gate")

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """``x[B, T, F]`` -> log-probabilities ``[B, T, 4]`` (reference segmenter.py:70-87).

        Eval mode: the sm_100a inference kernels (no autograd graph).  Training mode: dropout active and a
        differentiable result, the recurrences and their back-propagation through time in ``hss/model/_train.py``."""
        if self.training:
            from ._train import training_forward

            self._check_input(x)
            return training_forward(self, x)
        return self._run(x, True, False)[0]

    def training_loss(self, x: torch.Tensor, y: torch.Tensor):
        """``(loss, logp)`` of a training batch: what ``loss_fn(self(x).permute(0, 2, 1), y)`` computes in reference
        main.py:67-70, with linear + log-softmax + cross-entropy (and their backward) fused into one kernel each way."""
        from ._train import training_loss

        return training_loss(self, x, y)

    @torch.no_grad()
    def predict(self, x: torch.Tensor) -> torch.Tensor:
        """argmax labels ``[B, T]`` (int32), computed in the head kernel without materialising logp."""
        return self._run(x, False, True)[1]

    @torch.no_grad()
    def forward_with_labels(self, x: torch.Tensor):
        return self._run(x, True, True)
