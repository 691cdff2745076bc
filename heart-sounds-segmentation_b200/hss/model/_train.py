"""Training-mode forward of the segmenter: one autograd function per bidirectional LSTM layer.

The reference trains ``nn.LSTM`` under autograd (``main.py:67-82`` over ``hss/model/segmenter.py:80-87``).
Here a layer's forward (input projection + the recurrence that keeps the activated gates, cell states and raw h) and its
back-propagation through time are CUDA kernels of ``libhssb.so``:

* reference geometry (hidden 240, <= 64 inputs, weights inside the fp16-split range): ``hssb_lstm_train_forward_tc`` (K4 +
  K5m TRAIN variant on tcgen05, operands re-packed after every optimiser step by ``hssb_model_update``) and
  ``hssb_lstm_train_backward_tc`` (K5b), which also hands back dG split for the three-pass TF32 GEMMs and the bias gradients;
* anything else: cuBLAS projection + ``hssb_lstm_train_forward`` / ``hssb_lstm_train_backward`` (fp32 cluster / generic kernels).

The weight / input gradient GEMMs (``dG^T x``, ``dG^T h_prev``, ``dG W_ih``) are library GEMMs through torch -- as three TF32
tensor-core passes on operands split by ``hssb_split_tf32`` (``HSSB_TRAIN_GEMM=fp32``: torch's plain fp32 ones); ReLU and dropout
stay torch ops, so torch's RNG drives dropout exactly as in the reference; the head + loss is ``HeadLossFunction``.
"""
from __future__ import annotations

import os

import torch
import torch.nn.functional as F
from torch.autograd.function import once_differentiable

from .. import _lib


def _split_tf32(t: torch.Tensor):
    """``t = hi + lo`` with ``hi`` exactly representable in TF32 (``hssb_split_tf32``)."""
    t = t.contiguous()
    hi, lo = torch.empty_like(t), torch.empty_like(t)
    with torch.cuda.device(t.device):
        rc = _lib.lib().hssb_split_tf32(t.data_ptr(), t.numel(), hi.data_ptr(), lo.data_ptr(), _lib.stream_ptr())
    _lib.check(rc, "hssb_split_tf32")
    return hi, lo


class _Tf32Matmul:
    """Scope in which torch's fp32 matmuls run on the TF32 tensor cores (operands pre-split, so nothing is lost to the mode)."""

    def __enter__(self):
        self.prev = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = True

    def __exit__(self, *exc):
        torch.backends.cuda.matmul.allow_tf32 = self.prev


def _mm3(a, b, out=None):
    """``a @ b`` for pre-split operands ``a = (hi, lo)``, ``b = (hi, lo)``: ``lo.hi + hi.lo + hi.hi`` (small terms first),
    accumulated into ``out`` when given.  Call inside ``_Tf32Matmul``."""
    (ah, al), (bh, bl) = a, b
    out = al @ bh if out is None else out.addmm_(al, bh)
    out.addmm_(ah, bl)
    return out.addmm_(ah, bh)


def shifted_rows(M: int, H: int, d: int):
    """Row / column slices that pair dG with h_prev WITHOUT materialising h_prev: in the flattened ``[B*T]`` row order the
    previous hidden state of row ``r`` is row ``r - 1`` of ``out[:, :H]`` (forward, ``d = 0``) or row ``r + 1`` of ``out[:, H:]``
    (reverse, ``d = 1``).  Returns ``(rows of dG, rows of out, columns of out)``."""
    if d == 0:
        return slice(1, M), slice(0, M - 1), slice(0, H)
    return slice(0, M - 1), slice(1, M), slice(H, 2 * H)


def edge_fixup(g_edge: torch.Tensor, out2: torch.Tensor, h0_d: torch.Tensor, B: int, T: int, H: int, d: int) -> torch.Tensor:
    """What the row-shifted product gets wrong: at each window's first (forward) / last (reverse) step the partner of dG is
    ``h0``, not the neighbouring window's last / first output.  ``g_edge[B, 4H]``: dG at those steps; ``out2[B*T, 2H]``.
    Returns the ``[4H, H]`` correction: ``+ g_edge^T h0  - (the wrongly paired rows)``."""
    idx = torch.arange(B, device=out2.device) * T
    if d == 0:          # rows b*T (t = 0) were paired with row b*T - 1 (window b - 1, last step); window 0's row was left out
        return g_edge.t() @ h0_d - g_edge[1:].t() @ out2[idx[1:] - 1, :H]
    idx = idx + (T - 1)  # rows b*T + T - 1 were paired with row (b + 1)*T (window b + 1, first step); the last window's row was left out
    return g_edge.t() @ h0_d - g_edge[:-1].t() @ out2[idx[:-1] + 1, H:]


class BiLSTMLayerFunction(torch.autograd.Function):
    """``(x[B,T,F], h0[2,B,H], c0[2,B,H], 8 parameters) -> (out[B,T,2H], hn[2,B,H], cn[2,B,H])`` like
    ``nn.LSTM(bidirectional=True, batch_first=True)``."""

    @staticmethod
    def forward(ctx, x, h0, c0, w_ih, w_hh, b_ih, b_hh, w_ih_r, w_hh_r, b_ih_r, b_hh_r, packed=None):
        """``packed``: ``(hssb_model handle, layer index, workspace cache)`` of a model whose geometry the tcgen05 kernels cover --
        projection and recurrence then run on them (``hssb_lstm_train_forward_tc``); ``None``: cuBLAS projection + generic recurrence."""
        if not x.is_cuda:
            raise RuntimeError("the training recurrences run on the GPU (no CPU fallback)")
        tensors = [t.detach().to(torch.float32).contiguous() for t in (x, h0, c0, w_ih, w_hh, b_ih, b_hh, w_ih_r, w_hh_r, b_ih_r, b_hh_r)]
        x, h0, c0, w_ih, w_hh, b_ih, b_hh, w_ih_r, w_hh_r, b_ih_r, b_hh_r = tensors
        B, T, Fin = x.shape
        H = w_hh.shape[1]
        dev = x.device
        gates = torch.empty((2, B * T, 4 * H), dtype=torch.float32, device=dev)
        out = torch.empty((B, T, 2 * H), dtype=torch.float32, device=dev)
        cells = torch.empty((2, B * T, H), dtype=torch.float32, device=dev)
        hn = torch.empty((2, B, H), dtype=torch.float32, device=dev)
        cn = torch.empty((2, B, H), dtype=torch.float32, device=dev)
        lib = _lib.lib()
        if packed is not None and B * T:
            handle, layer, ws_cache = packed
            with torch.cuda.device(dev):
                ws = _lib.cached_workspace(ws_cache, dev, lib.hssb_model_workspace_bytes(handle, B, T))
                rc = lib.hssb_lstm_train_forward_tc(handle, layer, x.data_ptr(), B, T, h0.data_ptr(), c0.data_ptr(), gates.data_ptr(),
                                                    out.data_ptr(), cells.data_ptr(), hn.data_ptr(), cn.data_ptr(),
                                                    ws.data_ptr(), ws.numel(), _lib.stream_ptr())
            _lib.check(rc, "hssb_lstm_train_forward_tc")
        else:
            x2 = x.reshape(B * T, Fin)
            if B * T:
                torch.addmm(b_ih + b_hh, x2, w_ih.t(), out=gates[0])
                torch.addmm(b_ih_r + b_hh_r, x2, w_ih_r.t(), out=gates[1])
            whT, whT_r = w_hh.t().contiguous(), w_hh_r.t().contiguous()
            with torch.cuda.device(dev):
                rc = lib.hssb_lstm_train_forward(gates.data_ptr(), whT.data_ptr(), whT_r.data_ptr(), h0.data_ptr(), c0.data_ptr(),
                                                 B, T, H, out.data_ptr(), cells.data_ptr(), hn.data_ptr(), cn.data_ptr(), _lib.stream_ptr())
            _lib.check(rc, "hssb_lstm_train_forward")
        ctx.save_for_backward(x, h0, c0, w_ih, w_hh, w_ih_r, w_hh_r, gates, cells, out)
        ctx.tensor_cores = packed is not None
        return out, hn, cn

    @staticmethod
    @once_differentiable
    def backward(ctx, d_out, d_hn, d_cn):
        x, h0, c0, w_ih, w_hh, w_ih_r, w_hh_r, gates, cells, out = ctx.saved_tensors
        B, T, Fin = x.shape
        H = w_hh.shape[1]
        dev = x.device
        d_out = torch.zeros_like(out) if d_out is None else d_out.to(torch.float32).contiguous()
        d_hn = None if d_hn is None else d_hn.to(torch.float32).contiguous()
        d_cn = None if d_cn is None else d_cn.to(torch.float32).contiguous()
        dh0 = torch.empty_like(h0)
        dc0 = torch.empty_like(c0)
        lib = _lib.lib()
        p_hn = d_hn.data_ptr() if d_hn is not None else None
        p_cn = d_cn.data_ptr() if d_cn is not None else None
        M = B * T
        split_gemms = M > 0 and os.environ.get("HSSB_TRAIN_GEMM", "tf32x3") != "fp32"
        dG = g_hi = g_lo = db2 = None
        with torch.cuda.device(dev):
            if ctx.tensor_cores and M and os.environ.get("HSSB_TRAIN_BWD", "tc") == "tc":
                # the forward ran on the tcgen05 kernels (reference geometry, weights in the fp16-split range): so does this.  The gate
                # gradients come back in their own buffers (the saved activations stay intact) -- already split for the TF32 GEMMs
                if split_gemms:
                    g_hi = torch.empty((M, 2, 4 * H), dtype=torch.float32, device=dev)      # [row][direction][gate]: see below
                    g_lo = torch.empty_like(g_hi)
                else:
                    dG = torch.empty_like(gates)
                ws = torch.empty(lib.hssb_lstm_train_backward_tc_workspace_bytes(), dtype=torch.uint8, device=dev)
                db2 = torch.empty((2, 4 * H), dtype=torch.float32, device=dev)        # bias gradients, summed inside the kernel
                rc = lib.hssb_lstm_train_backward_tc(gates.data_ptr(), dG.data_ptr() if dG is not None else None,
                                                     g_hi.data_ptr() if g_hi is not None else None, g_lo.data_ptr() if g_lo is not None else None,
                                                     db2.data_ptr(), cells.data_ptr(), w_hh.data_ptr(), w_hh_r.data_ptr(), c0.data_ptr(), d_out.data_ptr(),
                                                     p_hn, p_cn, B, T, dh0.data_ptr(), dc0.data_ptr(), ws.data_ptr(), ws.numel(), _lib.stream_ptr())
                _lib.check(rc, "hssb_lstm_train_backward_tc")
            else:
                dG = gates.clone()               # this kernel turns activations into dG in place; keep the saved tensor intact
                rc = lib.hssb_lstm_train_backward(dG.data_ptr(), cells.data_ptr(), w_hh.data_ptr(), w_hh_r.data_ptr(), c0.data_ptr(),
                                                  d_out.data_ptr(), p_hn, p_cn, B, T, H, dh0.data_ptr(), dc0.data_ptr(), _lib.stream_ptr())
                _lib.check(rc, "hssb_lstm_train_backward")
        x2 = x.reshape(M, Fin)
        grads = []
        if not split_gemms:
            # the mm kernels autograd itself would run for nn.LSTM (SIMT fp32)
            # h_{prev}: forward direction = out[:, t-1, :H] (h0 at t = 0); reverse direction = out[:, t+1, H:] (h0 at t = T-1)
            hp_f = torch.cat([h0[0].unsqueeze(1), out[:, :-1, :H]], dim=1).reshape(M, H)
            hp_r = torch.cat([out[:, 1:, H:], h0[1].unsqueeze(1)], dim=1).reshape(M, H)
            for d, hp in enumerate((hp_f, hp_r)):
                g = dG[d]
                db = db2[d] if db2 is not None else g.sum(dim=0)
                grads.append((g.t() @ x2, g.t() @ hp, db, db.clone()))
            dx = (dG[0] @ w_ih + dG[1] @ w_ih_r).reshape(B, T, Fin) if ctx.needs_input_grad[0] else None
        else:
            # The same products as three TF32 tensor-core GEMMs each: hi.hi + hi.lo + lo.hi, fp32 accumulation (product error 2^-21).
            # dG^T h_prev without materialising h_prev: in the flattened [B*T] row order h_prev of row r is row r - 1 of out[:, :H]
            # (forward) / row r + 1 of out[:, H:] (reverse), i.e. one GEMM on views shifted by a row -- except at each window's first
            # (last) step, where the partner is h0 instead of the neighbouring window's last (first) step: B rows, fixed up in fp32.
            if g_hi is None:            # dG [2][M][4H] from the fp32 backward: per-direction operands
                a_hi, a_lo = _split_tf32(dG)
                gd_hi, gd_lo = (a_hi[0], a_hi[1]), (a_lo[0], a_lo[1])
                merged = None
            else:                       # K5b's [M][2][4H]: the two directions of a row side by side
                gd_hi, gd_lo = (g_hi[:, 0], g_hi[:, 1]), (g_lo[:, 0], g_lo[:, 1])
                merged = (g_hi.view(M, 8 * H), g_lo.view(M, 8 * H))
            xs = _split_tf32(x2)
            o_hi, o_lo = (t.view(M, 2 * H) for t in _split_tf32(out))
            o2 = out.view(M, 2 * H)
            first = torch.arange(B, device=dev) * T          # rows of t = 0
            last = first + (T - 1)                           # rows of t = T - 1
            sl = [shifted_rows(M, H, d) for d in range(2)]
            dx = None
            with _Tf32Matmul():
                dwh = tuple(_mm3((gd_hi[d][rg].t(), gd_lo[d][rg].t()), (o_hi[ro, co], o_lo[ro, co])) for d, (rg, ro, co) in enumerate(sl))
                if merged is not None:
                    # dG^T x for both directions in one GEMM ([8H, M] x [M, F]: twice the tiles), likewise dG [W_ih; W_ih_r]
                    both = _mm3((merged[0].t(), merged[1].t()), xs)
                    dwi = (both[:4 * H], both[4 * H:])
                    if ctx.needs_input_grad[0]:
                        dx = _mm3(merged, _split_tf32(torch.cat([w_ih, w_ih_r], dim=0)))
                else:
                    dwi = tuple(_mm3((gd_hi[d].t(), gd_lo[d].t()), xs) for d in range(2))
                    if ctx.needs_input_grad[0]:
                        for d, wi in enumerate((w_ih, w_ih_r)):
                            dx = _mm3((gd_hi[d], gd_lo[d]), _split_tf32(wi), out=dx)
            # the B edge rows, in plain fp32 (outside the TF32 switch: with T = 1 they are the whole gradient)
            edges = (gd_hi[0][first] + gd_lo[0][first], gd_hi[1][last] + gd_lo[1][last])      # hi + lo is the value itself
            for d in range(2):
                fix = edge_fixup(edges[d], o2, h0[d], B, T, H, d)
                db = db2[d] if db2 is not None else dG[d].sum(dim=0)
                grads.append((dwi[d], dwh[d] + fix, db, db.clone()))
            dx = dx.reshape(B, T, Fin) if dx is not None else None
        (dwi, dwh, dbi, dbh), (dwi_r, dwh_r, dbi_r, dbh_r) = grads
        return dx, dh0, dc0, dwi, dwh, dbi, dbh, dwi_r, dwh_r, dbi_r, dbh_r, None


def _layer(lstm: torch.nn.LSTM, x, h0, c0, packed=None):
    return BiLSTMLayerFunction.apply(
        x, h0, c0, lstm.weight_ih_l0, lstm.weight_hh_l0, lstm.bias_ih_l0, lstm.bias_hh_l0,
        lstm.weight_ih_l0_reverse, lstm.weight_hh_l0_reverse, lstm.bias_ih_l0_reverse, lstm.bias_hh_l0_reverse, packed)


def _packed_layers(model, dev):
    """``(packed_l0, packed_l1)`` when the model's geometry runs on the tcgen05 kernels, else ``(None, None)``.
    ``HSSB_TRAIN_IMPL`` = ``cluster`` / ``gather`` / ``stream`` selects one of the generic fp32 recurrences instead (validation)."""
    if os.environ.get("HSSB_TRAIN_IMPL", "tc") != "tc" or not getattr(model, "bidirectional", True):
        return None, None
    if model.lstm_1.hidden_size != 240 or model.lstm_1.input_size > 64 or any(p.dtype != torch.float32 for p in model.parameters()):
        return None, None
    handle = model._packed(dev)
    if not _lib.lib().hssb_model_uses_tensor_cores(handle):
        return None, None
    return (handle, 0, model._workspace), (handle, 1, model._workspace)


def training_forward(model, x: torch.Tensor) -> torch.Tensor:
    """``HeartSoundSegmenter.forward`` in training mode: reference segmenter.py:80-87 with dropout active."""
    p = model.linear.weight
    if not (x.is_cuda and p.is_cuda and x.device == p.device):
        raise RuntimeError("training mode needs the module and its input on the same CUDA device "
                           "(model.to('cuda')); there is no CPU fallback")
    h0 = model.h0.to(device=x.device, dtype=torch.float32)
    c0 = model.c0.to(device=x.device, dtype=torch.float32)
    p0, p1 = _packed_layers(model, x.device)
    out, hn, cn = _layer(model.lstm_1, x, h0, c0, p0)
    out = model.dropout(F.relu(out))
    out, _, _ = _layer(model.lstm_2, out, hn, cn, p1)
    out = model.dropout(F.relu(out))
    return F.log_softmax(model.linear(out), dim=2)


class HeadLossFunction(torch.autograd.Function):
    """``(act[B,T,2H], weight[4,2H], bias[4], target[B,T]) -> (loss, logp[B,T,4])``: linear + log_softmax (segmenter.py:86-87)
    and ``nn.CrossEntropyLoss`` on the permuted output (main.py:69-70) in one kernel each way (``hssb_ce_head_forward`` /
    ``hssb_ce_head_backward``).  ``logp`` is returned for the metrics and carries no gradient."""

    @staticmethod
    def forward(ctx, act, weight, bias, target):
        if not act.is_cuda:
            raise RuntimeError("the fused head + loss runs on the GPU (no CPU fallback)")
        act = act.detach().to(torch.float32).contiguous()
        weight = weight.detach().to(torch.float32).contiguous()
        bias = bias.detach().to(torch.float32).contiguous()
        target = target.to(device=act.device, dtype=torch.int64).contiguous()
        B, T, K = act.shape
        M = B * T
        logp = torch.empty((B, T, 4), dtype=torch.float32, device=act.device)
        loss_sum = torch.zeros(1, dtype=torch.float64, device=act.device)
        with torch.cuda.device(act.device):
            rc = _lib.lib().hssb_ce_head_forward(act.data_ptr(), M, K, weight.data_ptr(), bias.data_ptr(), target.data_ptr(),
                                                 logp.data_ptr(), loss_sum.data_ptr(), _lib.stream_ptr())
        _lib.check(rc, "hssb_ce_head_forward")
        ctx.save_for_backward(act, weight, target, logp)
        ctx.mark_non_differentiable(logp)
        return (loss_sum / max(M, 1)).to(torch.float32).reshape(()), logp

    @staticmethod
    @once_differentiable
    def backward(ctx, d_loss, _d_logp):
        act, weight, target, logp = ctx.saved_tensors
        B, T, K = act.shape
        M = B * T
        d_act = torch.empty_like(act)
        d_w = torch.zeros_like(weight)
        d_b = torch.zeros(4, dtype=torch.float32, device=act.device)
        # dL/d(loss) stays on the device (a host read here would stall the enqueue of the whole backward pass behind the forward)
        up = d_loss.detach().to(device=act.device, dtype=torch.float32).reshape(1).contiguous()
        with torch.cuda.device(act.device):
            rc = _lib.lib().hssb_ce_head_backward(act.data_ptr(), logp.data_ptr(), M, K, weight.data_ptr(), target.data_ptr(), 1.0 / max(M, 1),
                                                  up.data_ptr(), d_act.data_ptr(), d_w.data_ptr(), d_b.data_ptr(), _lib.stream_ptr())
        _lib.check(rc, "hssb_ce_head_backward")
        return d_act, d_w, d_b, None


def training_loss(model, x: torch.Tensor, y: torch.Tensor):
    """``(loss, logp)`` of one training batch: ``loss_fn(model(x).permute(0, 2, 1), y)`` of reference main.py:67-70 with the head
    and the loss fused (``HeadLossFunction``); ``logp`` feeds the metrics exactly as ``model(x)`` would."""
    p = model.linear.weight
    if not (x.is_cuda and p.is_cuda and x.device == p.device):
        raise RuntimeError("training mode needs the module and its input on the same CUDA device (model.to('cuda'))")
    model._check_input(x)
    h0 = model.h0.to(device=x.device, dtype=torch.float32)
    c0 = model.c0.to(device=x.device, dtype=torch.float32)
    p0, p1 = _packed_layers(model, x.device)
    out, hn, cn = _layer(model.lstm_1, x, h0, c0, p0)
    out = model.dropout(F.relu(out))
    out, _, _ = _layer(model.lstm_2, out, hn, cn, p1)
    out = model.dropout(F.relu(out))
    return HeadLossFunction.apply(out, model.linear.weight, model.linear.bias, y)
