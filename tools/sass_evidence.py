#!/usr/bin/env python3
"""Per-kernel counts of the SASS mnemonics that prove which hardware paths libb200vslam.so uses (B200_PROFILING.md):
UTMALDG (TMA tensor loads), UTCIMMA / UTCHMMA (tcgen05.mma), LDTM (tcgen05.ld), UTCBAR (tcgen05.commit), SYNCS (mbarrier),
UCGABAR (cluster barrier), DMMA (fp64 tensor core), POPC, IDP (dp4a/dp2a), VABSDIFF4, VIMNMX3, LDG.E.*.256 (256-bit loads).
usage: python tools/sass_evidence.py [path/to/libb200vslam.so] > profiles/r2_sass_evidence.txt"""
import collections
import re
import subprocess
import sys

so = sys.argv[1] if len(sys.argv) > 1 else "stella_vslam_b200/libb200vslam.so"
KEYS = ["UTMALDG", "UTCIMMA", "UTCHMMA", "LDTM", "UTCBAR", "SYNCS", "UCGABAR", "DMMA", "POPC", "IDP", "VABSDIFF4", "VIMNMX3", "LDG.E.ENL2.256",
        "STG.E.ENL2.256", "ATOMG", "RED", "MEMBAR", "DFMA", "BAR.SYNC"]
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
demangle = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n
kern, counts, total = None, collections.OrderedDict(), {}
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = m.group(1)
        counts[kern] = collections.Counter()
        total[kern] = 0
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and kern:
        total[kern] += 1
        op = m.group(1)
        for k in KEYS:
            if op.startswith(k) or (k.startswith("LDG") and op == k) or (k.startswith("STG") and op == k):
                counts[kern][k] += 1
print(f"# {so}: SASS mnemonic counts per kernel (cuobjdump -sass, sm_100a)")
for k, c in counts.items():
    name = re.sub(r"\(.*", "", demangle(k))
    tags = "  ".join(f"{kk}={v}" for kk, v in c.items())
    print(f"{name[:70]:70s} instr={total[k]:5d}  {tags}")
