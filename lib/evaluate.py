"""Per-talk inference (drop-in for the reference's lib.evaluate.infer, lib/evaluate.py:9-127).

Same signature and the same 4-tuple; same per-batch semantics (the +-1 frame fix-up, masking,
silent windows, NaN fill). Mechanism: one fused CUDA forward per batch, probabilities stay on
the device until the talk is complete, then ONE device->host copy.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np
import torch


def infer(model, dataloader, main_device, autoregression, loss_tag, vocab=None, loss_fn=None) -> Tuple[np.ndarray, np.ndarray, np.ndarray, None]:
    """Does inference for a single wav file"""
    if autoregression:
        raise NotImplementedError()
    if loss_tag != "bce" or vocab is not None:
        raise NotImplementedError("only the binary (bce) frame classifier is on the accelerated path")
    n = int(dataloader.dataset.duration_outframes)
    talk_probs = np.full(n, np.nan)
    talk_logits = np.full(n, np.nan)
    talk_targets = np.zeros(n)

    pending = []  # (probs_dev, logits_dev, starts, ends, included)
    for batch in iter(dataloader):
        audio = batch["audio"].to(main_device, non_blocking=True)
        in_mask = batch["in_mask"].to(main_device, non_blocking=True)
        out_mask = batch["out_mask"].to(main_device, non_blocking=True)
        starts, ends, included = batch["starts"], batch["ends"], batch["included"]
        with torch.no_grad():
            _, hidden = model.wav2vec_model(audio, in_mask)
            size1, size2 = hidden.shape[1], out_mask.shape[1]
            if size1 != size2:
                if size1 < size2:
                    out_mask = out_mask[:, :-1]
                    ends = [e - 1 for e in ends]
                else:
                    hidden = hidden[:, :-1, :]
            logits = model.seg_model(hidden, out_mask)
            probs = torch.sigmoid(logits)
            probs[~out_mask] = 0
            logits[~out_mask] = 0
        pending.append((probs, logits, starts, ends, included))

    for probs, logits, starts, ends, included in pending:
        probs = probs.detach().cpu().numpy()
        logits = logits.detach().cpu().numpy()
        for i in range(len(probs)):
            start, end = starts[i], ends[i]
            if included[i] and end > start:
                talk_probs[start:end] = probs[i, : end - start]
                talk_logits[start:end] = logits[i, : end - start]
            elif not included[i]:
                talk_probs[start:end] = 0
                talk_logits[start:end] = 0

    for j in np.where(np.isnan(talk_probs))[0]:
        lo, hi = max(0, j - 2), min(n, j + 3)
        talk_probs[j] = np.nanmean(talk_probs[lo:hi])
        talk_logits[j] = np.nanmean(talk_logits[lo:hi])
    return talk_probs, talk_logits, talk_targets, None
