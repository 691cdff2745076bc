"""Optimiser step of the reference training loop on one fused CUDA kernel pair.

Reference: ``main.py:130-135`` (``Adam(lr=0.01)`` + ``LambdaLR(lambda epoch: 0.9 ** epoch)``) and ``main.py:226``
(``pl.Trainer(gradient_clip_val=1)``, i.e. ``clip_grad_norm_(parameters, 1.0)`` before every optimiser step).  ``ClipAdam.step``
does the clip and the Adam update of ALL parameter tensors in two launches of ``libhssb.so`` (``hssb_clip_adam_step``) instead of
the ~10 small kernels per tensor of the eager ops; same arithmetic as ``torch.optim.Adam`` (default betas / eps, no weight decay).
"""
from __future__ import annotations

import ctypes
from typing import Iterable

import torch

from . import _lib


class ClipAdam:
    """``clip_grad_norm_(params, max_norm)`` + ``Adam.step()`` + the ``0.9 ** epoch`` learning-rate schedule."""

    def __init__(self, params: Iterable[torch.nn.Parameter], lr: float = 0.01, betas: tuple[float, float] = (0.9, 0.999), eps: float = 1e-8,
                 max_norm: float = 1.0, lr_decay: float = 0.9):
        self.params = [p for p in params if p.requires_grad]
        if not self.params or len(self.params) > 32:
            raise ValueError(f"ClipAdam takes 1..32 parameter tensors, got {len(self.params)}")
        for p in self.params:
            if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                raise ValueError("ClipAdam needs contiguous float32 CUDA parameters (no CPU fallback)")
        self.lr, self.betas, self.eps, self.max_norm, self.lr_decay = lr, betas, eps, max_norm, lr_decay
        self.exp_avg = [torch.zeros_like(p) for p in self.params]
        self.exp_avg_sq = [torch.zeros_like(p) for p in self.params]
        dev = self.params[0].device
        self._scratch = torch.zeros(1, dtype=torch.float64, device=dev)
        self.grad_norm = torch.zeros(1, dtype=torch.float32, device=dev)      # un-clipped global norm of the last step
        self.steps = 0
        self.epoch = 0

    def zero_grad(self, set_to_none: bool = True) -> None:
        for p in self.params:
            if p.grad is not None:
                if set_to_none:
                    p.grad = None
                else:
                    p.grad.zero_()

    def set_epoch(self, epoch: int) -> None:
        """LambdaLR(lr_lambda=lambda epoch: 0.9 ** epoch) of reference main.py:133-134."""
        self.epoch = epoch

    @property
    def current_lr(self) -> float:
        return self.lr * self.lr_decay ** self.epoch

    @torch.no_grad()
    def step(self) -> None:
        live = [(p, m, v) for p, m, v in zip(self.params, self.exp_avg, self.exp_avg_sq) if p.grad is not None]
        if not live:
            return
        n = len(live)
        grads = [p.grad if (p.grad.is_contiguous() and p.grad.dtype == torch.float32) else p.grad.float().contiguous() for p, _, _ in live]
        arr = lambda ts: (ctypes.c_void_p * n)(*[t.data_ptr() for t in ts])                    # noqa: E731
        numel = (ctypes.c_int64 * n)(*[p.numel() for p, _, _ in live])
        self.steps += 1
        dev = live[0][0].device
        with torch.cuda.device(dev):
            rc = _lib.lib().hssb_clip_adam_step(n, arr([p for p, _, _ in live]), arr(grads), arr([m for _, m, _ in live]), arr([v for _, _, v in live]),
                                                numel, self.current_lr, self.betas[0], self.betas[1], self.eps, self.steps, self.max_norm,
                                                self._scratch.data_ptr(), self.grad_norm.data_ptr(), _lib.stream_ptr())
        _lib.check(rc, "hssb_clip_adam_step")
        torch._foreach_add_([p for p, _, _ in live], 0.0)      # one in-place no-op that bumps ._version: the kernel wrote the
        #                                                        parameters behind autograd's back and packed-weight caches key on it
