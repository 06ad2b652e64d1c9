"""Per-kernel histogram of the SASS opcodes that prove the Blackwell-native claims (tcgen05 MMA, TMEM
load/store, TMA load/store/reduce, cluster barriers) — so the claim is checkable without the .so.

    python scripts/sass_histogram.py > profiles/sass_histogram_r02.md
"""
import re
import subprocess
import sys
from collections import Counter, OrderedDict
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "wav2vecsegmenter_b200" / "csrc" / "libw2vseg.so"
KEYS = OrderedDict([
    ("UTCHMMA", r"^UTCHMMA(?!.*2CTA)"), ("UTCHMMA.2CTA", r"^UTCHMMA.*2CTA"), ("UTCBAR", r"^UTCBAR"),
    ("LDTM", r"^LDTM"), ("STTM", r"^STTM"), ("UTMALDG", r"^UTMALDG"), ("UTMASTG", r"^UTMASTG"),
    ("UTMAREDG", r"^UTMAREDG"), ("SYNCS (mbarrier)", r"^SYNCS"), ("HMMA (mma.sync)", r"^HMMA"),
    ("MUFU.EX2", r"^MUFU\.EX2"), ("MUFU.TANH", r"^MUFU\.TANH"), ("STL/LDL (local)", r"^(STL|LDL)"),
])


def main():
    out = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    kernels = OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["cu++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            name = name.replace("w2v::(anonymous namespace)::", "").replace("w2v::<unnamed>::", "").replace("void ", "")
            depth = 0
            for i, ch in enumerate(name):      # cut the parameter list: first "(" outside template brackets
                depth += (ch == "<") - (ch == ">")
                if ch == "(" and depth == 0:
                    name = name[:i]
                    break
            cur = kernels.setdefault(name, Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and cur is not None:
            cur["_total"] += 1
            for key, pat in KEYS.items():
                if re.match(pat, m.group(1)):
                    cur[key] += 1
    print("# SASS opcode histogram per kernel of libw2vseg.so (sm_100a)\n")
    print("`cuobjdump -sass wav2vecsegmenter_b200/csrc/libw2vseg.so`, counted by `scripts/sass_histogram.py`. "
          "UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st (TMEM), UTMALDG/UTMASTG/UTMAREDG = TMA load / store / "
          "reduce-add, UTCBAR = tcgen05.commit, SYNCS = mbarrier ops. HMMA only appears in the mma.sync attention "
          "kept as a second implementation for tests.\n")
    cols = list(KEYS)
    print("| kernel | instr | " + " | ".join(cols) + " |")
    print("|---|---|" + "---|" * len(cols))
    for name, c in sorted(kernels.items(), key=lambda kv: -sum(v for k, v in kv[1].items() if k in ("UTCHMMA", "UTCHMMA.2CTA", "UTMALDG", "LDTM"))):
        if c["_total"] < 20:
            continue
        print(f"| `{name[:70]}` | {c['_total']} | " + " | ".join(str(c[k]) if c[k] else "" for k in cols) + " |")


if __name__ == "__main__":
    main()
