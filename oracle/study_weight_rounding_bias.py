"""TEST / ANALYSIS INFRASTRUCTURE (oracle side, never imported by the product path). CPU study (oracle, fp32 arithmetic): which weight groups' bf16 ROUNDING produces the frame-independent
logit offset of the large (24/24) configuration (profiles/parity_r01.md: +0.033)?
Rounds one group of matrices at a time to bf16 and reports mean / std of the logit change."""
import re
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent   # oracle/ -> repo root
sys.path.insert(0, str(ROOT))
from oracle import sfc_oracle  # noqa: E402
from wav2vecsegmenter_b200 import synth  # noqa: E402

torch.set_num_threads(8)
spec = synth.LARGE_ALL
sd = synth.random_state_dict(spec, 0)
n = 160_000
audio = sfc_oracle.normalize_rows(synth.synthetic_audio(n, 40)[None, :], [True])
T = sfc_oracle.conv_out_frames(n)
mask = torch.ones(1, T, dtype=torch.bool)


def logits(state):
    with torch.no_grad():
        _, lg, _, _ = sfc_oracle.batch_probs(state, audio, [n], mask, spec.keep_layers, spec.head_heads)
    return lg[0]


base = logits(sd)
GROUPS = {
    "all matrices": r"(conv\.weight|projection\.weight|_proj\.weight|dense\.weight|in_proj_weight|linear\d\.weight|original1)$",
    "head (seg_model) matrices": r"^seg_model\..*(in_proj_weight|out_proj\.weight|linear\d\.weight)$",
    "encoder FFN + adapter": r"encoder\.layers\.\d+\.(feed_forward|ffn_adapter)\..*weight$",
    "encoder attention q,k,v,o": r"encoder\.layers\.\d+\.attention\..*weight$",
    "conv extractor + projection + pos conv": r"(conv_layers\.\d\.conv\.weight|projection\.weight|original1)$",
    "last 4 encoder layers (all matrices)": r"encoder\.layers\.(20|21|22|23)\..*(_proj|dense)\.weight$",
}
print(f"reference logits: mean {base.mean():+.4f} std {base.std():.4f}")
for name, pat in GROUPS.items():
    st = {k: (v.to(torch.bfloat16).to(torch.float32) if re.search(pat, k) and v.dim() >= 2 else v) for k, v in sd.items()}
    nr = sum(1 for k, v in sd.items() if re.search(pat, k) and v.dim() >= 2)
    d = logits(st) - base
    print(f"{name:42s} ({nr:3d} tensors): logit change mean {d.mean():+.4f} std {d.std():.4f}")
