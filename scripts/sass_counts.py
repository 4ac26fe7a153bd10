"""Per-kernel counts of the SASS mnemonics that prove a Blackwell-native kernel (B200_PROFILING.md):
UTCHMMA (tcgen05.mma; .2CTA = cta_group::2), UTMALDG (TMA loads), LDTM (tcgen05.ld), UTCBAR (tcgen05.commit),
HMMA (legacy mma.sync path -- only the LSTM step kernel uses it).

    python scripts/sass_counts.py [path/to/libb2ocr.so] > profiles/r2_sass_counts.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "keras-ocr_b200", "libb2ocr.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
demangle = lambda s: subprocess.run(["c++filt", s], capture_output=True, text=True).stdout.strip()
PAT = {"UTCHMMA": r"\bUTCHMMA", "UTCHMMA.2CTA": r"\bUTCHMMA\.2CTA", "UTMALDG": r"\bUTMALDG", "LDTM": r"\bLDTM", "UTCBAR": r"\bUTCBAR",
       "HMMA": r"\bHMMA", "STG.E.256": r"STG\.E\.(ENL2\.)?256", "SYNCS": r"\bSYNCS"}
total = collections.Counter()
rows = []
for block in sass.split("Function : ")[1:]:
    name = block.split("\n", 1)[0].strip()
    counts = {k: len(re.findall(p, block)) for k, p in PAT.items()}
    total.update(counts)
    short = re.sub(r"\(.*", "", demangle(name).replace("(anonymous namespace)::", "").replace("void ", ""))
    rows.append((short, counts))
print(f"# {os.path.relpath(so, ROOT)}: {len(rows)} kernels; totals: " + ", ".join(f"{k} {v}" for k, v in total.items()))
print(f"{'kernel':58s} " + " ".join(f"{k:>12s}" for k in PAT))
for short, counts in sorted(rows, key=lambda r: r[0]):
    if any(counts.values()):
        print(f"{short[:58]:58s} " + " ".join(f"{counts[k]:12d}" for k in PAT))
