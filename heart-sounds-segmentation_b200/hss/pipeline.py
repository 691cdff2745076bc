"""Batch pipeline of the hot path: FSST of batch i+1 overlapped with the BiLSTM of batch i.

The reference transforms and segments strictly one after the other (``hss/datasets/heart_sounds.py:199-201`` then
``main.py:64-65``).  On the GPU the BiLSTM is two latency-bound recurrences that leave about a third of the SMs idle, while the
FSST kernels are short and wide -- so a serving / evaluation loop that has the NEXT batch at hand runs its transform on a second
CUDA stream underneath the current batch's model.  Per-batch results are identical to ``model(fsst.batch(x))``: the same kernels
run on the same data, only their placement in time changes.  The transform is enqueued AFTER the model's kernels and behind
``HeartSoundSegmenter.side_gate``: it starts once the layer-1 recurrence's clusters are on the machine (started earlier, its
short CTAs would scatter over all SMs and delay the placement of those clusters by the length of the transform) and has the
first half of that recurrence to itself -- the model's own middle-out projection launch only arrives once its first tile exists.

    pipe = SegmentationPipeline(fsst, model)
    for x in batches:                      # x[B, N] on the device (or pinned host memory)
        logp, labels = pipe(x)             # results of THIS batch, valid on the current stream
    pipe.close()

``pipe(x, prefetch=x_next)`` transforms ``x_next`` while ``x`` is segmented; a loop that knows its next batch passes it, a
loop that does not simply pays the un-overlapped transform (first call or missing prefetch).
"""
from __future__ import annotations

import torch


class SegmentationPipeline:
    def __init__(self, fsst, model, device: torch.device | str | None = None):
        self.fsst, self.model = fsst, model
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.side = torch.cuda.Stream(self.device)
        self._ready: dict[int, tuple[torch.Tensor, torch.cuda.Event, torch.Tensor]] = {}     # id(x) -> (features, event, x)

    def _transform_async(self, x: torch.Tensor, produced: torch.cuda.Event | None = None) -> None:
        """Enqueue ``fsst.batch(x)`` on the side stream, behind whatever produced ``x`` on the current stream (``produced``: an
        event recorded there before the model's kernels were enqueued) and behind the model's side gate."""
        main = torch.cuda.current_stream(self.device)
        if produced is None:
            self.side.wait_stream(main)
        else:
            self.side.wait_event(produced)
            self.model.side_gate(self.side)
        with torch.cuda.stream(self.side):
            feats = self.fsst.batch(x.to(self.device, non_blocking=True))
            if hasattr(self.model, "prepare") and not self.model.training:
                feats = self.model.prepare(feats)          # the forward's operand split, off the critical path as well
            ev = torch.cuda.Event()
            ev.record(self.side)
        self._ready[id(x)] = (feats, ev, x)

    def __call__(self, x: torch.Tensor, prefetch: torch.Tensor | None = None, labels_only: bool = False):
        """``(logp[B, N, 4], labels[B, N])`` (or only the labels) of ``x``; ``prefetch``: the next batch, transformed meanwhile."""
        main = torch.cuda.current_stream(self.device)
        entry = self._ready.pop(id(x), None)
        if entry is None:
            feats = self.fsst.batch(x.to(self.device, non_blocking=True))
        else:
            feats, ev, _ = entry
            main.wait_event(ev)
            feats.record_stream(main)
        want_prefetch = prefetch is not None and id(prefetch) not in self._ready
        produced = None
        if want_prefetch:
            produced = torch.cuda.Event()
            produced.record(main)
        out = (None, self.model.predict(feats)) if labels_only else self.model.forward_with_labels(feats)
        # the next transform is enqueued BEHIND the model's kernels and its side gate (see the module docstring)
        if want_prefetch:
            self._transform_async(prefetch, produced)
        return out

    def close(self) -> None:
        torch.cuda.current_stream(self.device).wait_stream(self.side)
        self._ready.clear()
