"""Batch collation for SFC inference (drop-in for the reference's lib.datautils.CollateFn,
lib/datautils.py:57-142). Same dict, same values; no tokenizer download at import time."""
from __future__ import annotations

import torch


class CollateFn:
    def __init__(self, pad_token_id) -> None:
        self.pad_token_id = pad_token_id

    def __call__(self, batch: list) -> dict:
        waves = [ex[0] for ex in batch]
        starts = [ex[2] for ex in batch]
        ends = [ex[3] for ex in batch]
        included = [bool(w.sum()) for w in waves]
        in_len = [len(w) for w in waves]
        out_len = [e - s for s, e in zip(starts, ends)]
        bs, lmax, tmax = len(batch), max(in_len), max(out_len)

        audio = torch.zeros(bs, lmax, dtype=waves[0].dtype)
        for i, w in enumerate(waves):
            audio[i, : len(w)] = w
        target = None
        if batch[0][1] is not None:
            target = torch.full((bs, tmax), self.pad_token_id, dtype=batch[0][1].dtype)
            for i, ex in enumerate(batch):
                target[i, : len(ex[1])] = ex[1]

        # per-row normalisation over the PADDED row, unbiased std, silent rows untouched
        keep = torch.tensor(included, dtype=torch.bool)
        if keep.any():
            rows = audio[keep]
            audio[keep] = (rows - rows.mean(dim=1, keepdim=True)) / rows.std(dim=1, keepdim=True)

        in_mask = (torch.arange(lmax)[None, :] < torch.tensor(in_len)[:, None]).long()
        out_mask = torch.arange(tmax)[None, :] < torch.tensor(out_len)[:, None]
        return {"audio": audio, "target": target, "in_mask": in_mask, "out_mask": out_mask,
                "included": included, "starts": starts, "ends": ends}
