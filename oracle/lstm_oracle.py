"""CPU oracle for the BiLSTM segmenter -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import this module.

Restates reference ``hss/model/segmenter.py`` (which cannot travel to the GPU box):
  * ``reference_params``     replays the constructor's RNG draw order (segmenter.py:38-67): after
                             ``torch.manual_seed(seed)`` the ctor draws h0, c0, then builds lstm_1,
                             lstm_2, linear.
  * ``forward_torch``        segmenter.py:80-87 with torch's own CPU ``nn.LSTM`` (MKL) -- the same
                             library call the reference makes on a CPU host.
  * ``forward_manual``       the recurrence written out step by step (gate order i,f,g,o; biases
                             b_ih + b_hh; layer 2 seeded with layer 1's final (hn, cn); ReLU between),
                             any dtype -- float64 gives the "truth" used to adjudicate label flips.

Pinned: ``tests/golden/lstm_*.npz`` hold outputs of the REAL reference module imported from
/root/reference (generator: ``tests/golden/make_golden.py``); ``tests/test_oracle_lstm.py`` checks
this restatement against them.
"""
from __future__ import annotations

import torch
from torch import nn

PARAM_NAMES = [
    f"{layer}.{kind}_l0{suffix}"
    for layer in ("lstm_1", "lstm_2")
    for suffix in ("", "_reverse")
    for kind in ("weight_ih", "weight_hh", "bias_ih", "bias_hh")
] + ["linear.weight", "linear.bias"]


def _modules(input_size: int, hidden_size: int):
    lstm_1 = nn.LSTM(input_size=input_size, hidden_size=hidden_size, bidirectional=True, batch_first=True)
    lstm_2 = nn.LSTM(input_size=hidden_size * 2, hidden_size=hidden_size, bidirectional=True, batch_first=True)
    linear = nn.Linear(in_features=hidden_size * 2, out_features=4, bias=True)
    return lstm_1, lstm_2, linear


def reference_params(seed: int, input_size: int = 44, batch_size: int = 1, hidden_size: int = 240):
    """Weights + (h0, c0) exactly as ``torch.manual_seed(seed); HeartSoundSegmenter(...)`` makes them."""
    torch.manual_seed(seed)
    h0 = torch.randn(2, batch_size, hidden_size)
    c0 = torch.randn(2, batch_size, hidden_size)
    lstm_1, lstm_2, linear = _modules(input_size, hidden_size)
    params = {}
    for prefix, mod in (("lstm_1", lstm_1), ("lstm_2", lstm_2), ("linear", linear)):
        for k, v in mod.state_dict().items():
            params[f"{prefix}.{k}"] = v.detach().clone()
    return params, h0, c0


def forward_torch(params: dict, h0: torch.Tensor, c0: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    """segmenter.py:80-87, eval mode (dropout = identity), torch CPU kernels."""
    hidden = h0.shape[2]
    lstm_1, lstm_2, linear = _modules(x.shape[2], hidden)
    lstm_1.load_state_dict({k.split(".", 1)[1]: v for k, v in params.items() if k.startswith("lstm_1.")})
    lstm_2.load_state_dict({k.split(".", 1)[1]: v for k, v in params.items() if k.startswith("lstm_2.")})
    linear.load_state_dict({k.split(".", 1)[1]: v for k, v in params.items() if k.startswith("linear.")})
    with torch.no_grad():
        out, (hn, cn) = lstm_1(x, (h0, c0))
        out = torch.relu(out)
        out, _ = lstm_2(out, (hn, cn))
        out = torch.relu(out)
        out = linear(out)
        return torch.log_softmax(out, dim=2)


def _lstm_dir(x, w_ih, w_hh, b_ih, b_hh, h, c, reverse: bool):
    B, T, _ = x.shape
    H = h.shape[1]
    out = x.new_empty(B, T, H)
    steps = range(T - 1, -1, -1) if reverse else range(T)
    bias = b_ih + b_hh
    for t in steps:
        gates = x[:, t] @ w_ih.T + h @ w_hh.T + bias
        i = torch.sigmoid(gates[:, 0 * H:1 * H])
        f = torch.sigmoid(gates[:, 1 * H:2 * H])
        g = torch.tanh(gates[:, 2 * H:3 * H])
        o = torch.sigmoid(gates[:, 3 * H:4 * H])
        c = f * c + i * g
        h = o * torch.tanh(c)
        out[:, t] = h
    return out, h, c


def forward_manual(params: dict, h0, c0, x, dtype=torch.float64, return_logits: bool = False):
    """Step-by-step restatement in ``dtype``."""
    p = {k: v.to(dtype) for k, v in params.items()}
    h0, c0, x = h0.to(dtype), c0.to(dtype), x.to(dtype)
    inp = x
    h_init, c_init = h0, c0
    for layer in ("lstm_1", "lstm_2"):
        outs, hs, cs = [], [], []
        for d, suffix in enumerate(("", "_reverse")):
            o, h, c = _lstm_dir(
                inp, p[f"{layer}.weight_ih_l0{suffix}"], p[f"{layer}.weight_hh_l0{suffix}"],
                p[f"{layer}.bias_ih_l0{suffix}"], p[f"{layer}.bias_hh_l0{suffix}"],
                h_init[d], c_init[d], reverse=(d == 1),
            )
            outs.append(o); hs.append(h); cs.append(c)
        h_init, c_init = torch.stack(hs), torch.stack(cs)
        inp = torch.relu(torch.cat(outs, dim=2))
    logits = inp @ p["linear.weight"].T + p["linear.bias"]
    if return_logits:
        return logits
    return torch.log_softmax(logits, dim=2)


def label_report(logp_test: torch.Tensor, logp_ref: torch.Tensor, logp_truth: torch.Tensor | None = None) -> dict:
    """Label flips of ``test`` vs ``ref`` with the reference's top-2 margin at the flipped positions."""
    lab_t = logp_test.argmax(dim=-1)
    lab_r = logp_ref.argmax(dim=-1)
    flips = lab_t != lab_r
    top2 = logp_ref.topk(2, dim=-1).values
    margin = (top2[..., 0] - top2[..., 1])
    rep = {
        "labels": int(lab_r.numel()),
        "flips": int(flips.sum()),
        "max_margin_flipped": float(margin[flips].max()) if flips.any() else 0.0,
        "min_margin": float(margin.min()),
        "max_abs_dlogp": float((logp_test.double() - logp_ref.double()).abs().max()),
    }
    if logp_truth is not None:
        lab_truth = logp_truth.argmax(dim=-1)
        rep["flips_test_vs_truth"] = int((lab_t != lab_truth).sum())
        rep["flips_ref_vs_truth"] = int((lab_r != lab_truth).sum())
    return rep


def training_reference(params: dict, h0: torch.Tensor, c0: torch.Tensor, x: torch.Tensor, y: torch.Tensor,
                       masks: tuple[torch.Tensor, torch.Tensor] | None = None, dtype=torch.float32):
    """One training-step forward + backward of the reference on the CPU: segmenter.py:80-87 in training mode
    (``nn.LSTM`` under autograd), ``CrossEntropyLoss`` on the permuted output as in main.py:67-70.

    Dropout (p = 0.2) cannot replay another device's RNG stream, so the two masks (already scaled by 1 / 0.8)
    are passed in; ``None`` = no dropout.  Returns ``(loss, logp, {name: grad}, grad_x)``.
    """
    hidden = h0.shape[2]
    lstm_1, lstm_2, linear = _modules(x.shape[2], hidden)
    lstm_1.load_state_dict({k.split(".", 1)[1]: v for k, v in params.items() if k.startswith("lstm_1.")})
    lstm_2.load_state_dict({k.split(".", 1)[1]: v for k, v in params.items() if k.startswith("lstm_2.")})
    linear.load_state_dict({k.split(".", 1)[1]: v for k, v in params.items() if k.startswith("linear.")})
    mods = {"lstm_1": lstm_1.to(dtype), "lstm_2": lstm_2.to(dtype), "linear": linear.to(dtype)}
    x = x.detach().to(dtype).requires_grad_(True)
    out, (hn, cn) = mods["lstm_1"](x, (h0.to(dtype), c0.to(dtype)))
    out = torch.relu(out)
    if masks is not None:
        out = out * masks[0].to(dtype)
    out, _ = mods["lstm_2"](out, (hn, cn))
    out = torch.relu(out)
    if masks is not None:
        out = out * masks[1].to(dtype)
    logp = torch.log_softmax(mods["linear"](out), dim=2)
    loss = nn.functional.cross_entropy(logp.permute(0, 2, 1), y)
    loss.backward()
    grads = {f"{prefix}.{k}": p.grad.detach().clone() for prefix, mod in mods.items() for k, p in mod.named_parameters()}
    return loss.detach(), logp.detach(), grads, x.grad.detach().clone()
