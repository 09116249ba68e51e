#!/usr/bin/env python
"""Mnemonic counts per kernel of librandblas_b200.so (cuobjdump -sass): the evidence that the tensor-core / TMA
instructions the design names are what the binary contains. Writes profiles/r02_sass_summary.txt.

    UTCHMMA / UTCQMMA = tcgen05.mma      LDTM / STTM = tcgen05.ld / st     UTMALDG = cp.async.bulk.tensor (TMA load)
    UBLKCP = cp.async.bulk                UTMAPF = TMA prefetch            DMMA = mma.sync f64       LDGSTS = cp.async
    RED / ATOMG = global reductions       SYNCS = mbarrier ops             UTCBAR = tcgen05.commit
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "randblas_b200", "librandblas_b200.so")
KEYS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMAPF", "UBLKCP", "UBLKRED", "DMMA", "HMMA", "LDGSTS",
        "SYNCS", "RED", "ATOMG", "ATOMS", "MUFU", "F2F", "DFMA", "IMAD.WIDE", "LOP3", "STG", "LDG", "LDS", "STS", "SHFL"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            op = m.group(1)
            kernels[cur]["_total"] += 1
            for k in KEYS:
                if op == k or op.startswith(k + ".") or (k == "IMAD.WIDE" and op.startswith("IMAD.WIDE")):
                    kernels[cur][k] += 1
    names = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    lines = [__doc__.strip(), "", f"library: randblas_b200/librandblas_b200.so ({os.path.getsize(LIB)} bytes), {len(kernels)} kernels", ""]
    tot = collections.Counter()
    for (mangled, c), name in zip(kernels.items(), names):
        short = name.replace("(anonymous namespace)::", "")          # before the argument list is cut at the first "("
        short = re.sub(r"\(.*", "", short)
        short = short.replace("void rb::", "")
        hits = "  ".join(f"{k}={c[k]}" for k in KEYS if c[k])
        lines.append(f"{short[:110]:110s} total={c['_total']:6d}  {hits}")
        tot.update(c)
    lines += ["", "whole library: " + "  ".join(f"{k}={tot[k]}" for k in KEYS if tot[k])]
    path = os.path.join(ROOT, "profiles", "r02_sass_summary.txt")
    open(path, "w").write("\n".join(lines) + "\n")
    print(lines[-1])


if __name__ == "__main__":
    main()
