"""Per-talk inference and dev-set scoring (drop-in for the reference's lib.evaluate.infer /
evaluate, lib/evaluate.py:9-127 and :130-214).

Same signatures and return values; same per-batch semantics (the +-1 frame fix-up, masking, silent
windows, NaN fill, targets, per-batch loss). Mechanism: one fused CUDA forward per batch,
probabilities stay on the device until the talk is complete, then ONE device->host copy.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np
import torch


def infer(model, dataloader, main_device, autoregression, loss_tag, vocab=None, loss_fn=None) -> Tuple[np.ndarray, np.ndarray, np.ndarray, float | None]:
    """Does inference for a single wav file"""
    if autoregression:
        raise NotImplementedError()
    if loss_tag != "bce" or vocab is not None:
        raise NotImplementedError("only the binary (bce) frame classifier is on the accelerated path")
    n = int(dataloader.dataset.duration_outframes)
    talk_probs = np.full(n, np.nan)
    talk_logits = np.full(n, np.nan)
    talk_targets = np.zeros(n)

    pending = []  # (probs_dev, logits_dev, starts, ends, included, targets, loss_dev)
    for batch in iter(dataloader):
        targets = batch["target"]
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
            loss = None
            if loss_fn:
                # reference :76-82: per-point loss on the common length, masked, summed per window,
                # averaged over the batch (logits / targets stay truncated afterwards, as there)
                dev_targets = targets.to(main_device)
                size = min(logits.shape[1], dev_targets.shape[1])
                logits, dev_targets, targets = logits[:, :size], dev_targets[:, :size], targets[:, :size]
                per_point = loss_fn(logits, dev_targets)
                per_point[~out_mask] = 0
                loss = per_point.sum(dim=1).mean()
            probs = torch.sigmoid(logits)
            probs[~out_mask] = 0
            logits[~out_mask] = 0
        pending.append((probs, logits, starts, ends, included, targets, loss))

    all_losses = []
    for probs, logits, starts, ends, included, targets, loss in pending:
        if loss is not None and float(loss) != 0.0:    # the reference tests `if loss:` (:93)
            all_losses.append(loss.detach().cpu().numpy().item())
        probs = probs.detach().cpu().numpy()
        logits = logits.detach().cpu().numpy()
        for i in range(len(probs)):
            start, end = starts[i], ends[i]
            if included[i] and end > start:
                talk_probs[start:end] = probs[i, : end - start]
                talk_logits[start:end] = logits[i, : end - start]
                if targets is not None:
                    talk_targets[start:end] = targets[i, : end - start].numpy()
            elif not included[i]:
                talk_probs[start:end] = 0
                talk_logits[start:end] = 0

    for j in np.where(np.isnan(talk_probs))[0]:
        lo, hi = max(0, j - 2), min(n, j + 3)
        talk_probs[j] = np.nanmean(talk_probs[lo:hi])
        talk_logits[j] = np.nanmean(talk_logits[lo:hi])
    # reference :112-113: the average is only formed when the LAST batch loss is truthy
    avg_loss = float(np.mean(all_losses)) if pending and pending[-1][6] is not None and float(pending[-1][6]) != 0.0 else None
    return talk_probs, talk_logits, talk_targets, avg_loss


def evaluate(dataloader_generator, model, main_device, autoregression, loss_tag, vocab, loss_fn=None) -> dict:
    """Does inference and evaluation for a dev/test set (reference lib/evaluate.py:130-214).

    Kept bug-compatible with the reference on purpose: the averaged probabilities are
    divided by `inference_times` a second time before thresholding (:179,:185), and `eval_loss` is
    the loss of the LAST talk only (`all_losses` is re-created per talk, :147). One deviation: with
    `loss_fn=None` the reference dies on an unbound `eval_loss` (:211); here the metrics are returned
    without that key."""
    from sklearn.metrics import f1_score, precision_score, recall_score

    if loss_tag != "bce":
        raise NotImplementedError("only the binary (bce) frame classifier is on the accelerated path")
    all_preds, all_targets = np.array([]), np.array([])
    all_losses = []
    for talk_id in dataloader_generator.get_talk_ids():
        inference_times = dataloader_generator.dataset.inference_times
        probs, targets, losses = None, None, None
        all_losses = []
        for iteration in range(inference_times):
            dataloader = dataloader_generator.generate(talk_id, iteration)
            p, _, t, loss = infer(model, dataloader, main_device, autoregression, loss_tag, vocab, loss_fn)
            if probs is None:
                probs, targets = p.copy(), t.copy()
                losses = loss if loss else None
            else:
                probs += p
                if loss:
                    losses += loss
        probs /= inference_times
        if losses:
            losses /= inference_times
        preds = probs / inference_times > 0.5
        all_preds = np.append(all_preds, preds)
        all_targets = np.append(all_targets, targets)
        if loss_fn:
            all_losses.append(losses)
    all_targets, all_preds = all_targets.astype(bool), all_preds.astype(bool)
    results = {
        "eval_accuracy": round(f1_score(all_targets, all_preds, average="micro"), 4),
        "eval_f1": round(f1_score(all_targets, all_preds, average="binary"), 4),
        "eval_precision": round(precision_score(all_targets, all_preds), 4),
        "eval_recall": round(recall_score(all_targets, all_preds), 4),
    }
    if loss_fn and all_losses and all_losses[0]:
        results["eval_loss"] = np.mean(all_losses)
    return results
