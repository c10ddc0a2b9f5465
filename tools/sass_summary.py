#!/usr/bin/env python
"""Per-kernel SASS evidence for the built library: counts of the mnemonics that show which hardware path a kernel uses
(UTCIMMA = tcgen05.mma kind::i8, LDTM/STTM = tensor-memory loads/stores, UBLKCP = bulk-async (TMA) copy, SYNCS = mbarrier,
FFMA2/FADD2/FMUL2 = packed two-lane fp32, IMMA = mma.sync int8, IMAD/PRMT/SHF = the integer FFT's mix).
  python tools/sass_summary.py > profiles/r02_sass_summary.txt      (runs on the build host: cuobjdump, no GPU)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "pothoscomms_b200", "libb200comms.so")
KEYS = ["UTCIMMA", "LDTM", "STTM", "UBLKCP", "SYNCS", "FFMA2", "FADD2", "FMUL2", "FFMA", "IMMA", "IMAD", "PRMT", "SHF", "LDS", "STS", "LDG", "STG", "SHFL", "BAR"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            kernels[cur][m.group(1)] += 1
            kernels[cur]["_total"] += 1
    names = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    print(f"# {os.path.relpath(LIB, ROOT)}: {len(kernels)} kernels, {os.path.getsize(LIB) / 1e6:.1f} MB")
    print("kernel | total | " + " | ".join(KEYS))
    agg = collections.OrderedDict()
    for name, c in zip(names, kernels.values()):
        short = re.sub(r"\(.*", "", re.sub(r"^void b200c::|^void ", "", name))
        print(short + " | " + str(c["_total"]) + " | " + " | ".join(str(c[k]) for k in KEYS))
        fam = re.sub(r"<.*", "", short)
        agg.setdefault(fam, collections.Counter()).update(c)
    print("\n# per family (all instantiations summed)")
    print("family | total | " + " | ".join(KEYS))
    for fam, c in agg.items():
        print(fam + " | " + str(c["_total"]) + " | " + " | ".join(str(c[k]) for k in KEYS))


if __name__ == "__main__":
    sys.exit(main())
