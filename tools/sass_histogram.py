"""Per-kernel SASS opcode histogram of libmrag.so (CPU box: cuobjdump only).

    python tools/sass_histogram.py > profiles/<round>_sass_opcode_histogram.txt

Counts, per kernel, the mnemonics that prove the Blackwell paths (B200_PROFILING.md): UTCHMMA (tcgen05.mma; .2CTA = cta_group::2),
UTMALDG (TMA tensor loads), UBLKCP (bulk copies, incl. shared::cta -> shared::cluster), LDTM / STTM (tcgen05.ld / st), UTCBAR
(tcgen05.commit), SYNCS (mbarrier), plus the total instruction count.
"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "mobius-rag_b200", "libmrag.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
WATCH = ["UTCHMMA", "UTCHMMA.2CTA", "UTMALDG", "UBLKCP", "LDTM", "STTM", "UTCBAR", "SYNCS", "UTCATOMSWS", "MEMBAR", "CCTL", "HMMA", "LDGSTS"]
kernels, cur = {}, None
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = kernels.setdefault(m.group(1), {"total": 0})
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and cur is not None:
        op = m.group(1)
        cur["total"] += 1
        base = op.split(".")[0]
        if base in WATCH:
            cur[base] = cur.get(base, 0) + 1
        if op.startswith("UTCHMMA") and ".2CTA" in op:
            cur["UTCHMMA.2CTA"] = cur.get("UTCHMMA.2CTA", 0) + 1
arch = subprocess.run(["cuobjdump", "-lelf", lib], capture_output=True, text=True).stdout.strip().splitlines()
print("library:", os.path.relpath(lib, ROOT), "|", "; ".join(a.strip() for a in arch))
print(f"{'kernel':<78} {'instrs':>7}  " + " ".join(f"{w:>8}" for w in WATCH[:10]))
names = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
tot = {w: 0 for w in WATCH}
for mangled, name in sorted(zip(kernels, names), key=lambda x: x[1]):
    k = kernels[mangled]
    short = re.sub(r"\(.*", "", name).replace("mrag::", "")
    print(f"{short[:78]:<78} {k['total']:>7}  " + " ".join(f"{k.get(w, 0):>8}" for w in WATCH[:10]))
    for w in WATCH:
        tot[w] += k.get(w, 0)
print(f"{'TOTAL':<78} {sum(k['total'] for k in kernels.values()):>7}  " + " ".join(f"{tot[w]:>8}" for w in WATCH[:10]))
