#!/usr/bin/env python
"""Per-instruction digest of the source page of an .ncu-rep (ncu --set full --import-source on): share of warp-state samples
and of executed warp instructions per SASS instruction, with the two top stall reasons.
python tools/ncu_hot.py rep kernel-regex [min-share-percent]"""
import collections
import csv
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    thr = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[1]
    iS, iN, iE, iA = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Address")
    iT = hdr.index("Avg. Threads Executed")
    seen, sass = set(), []
    for r in rows[2:]:
        if len(r) != len(hdr) or r[iA] in seen or not r[iS]:
            continue
        try:
            int(r[iA], 16)
        except ValueError:
            continue
        if rows[0][1] and False:
            pass
        seen.add(r[iA])
        sass.append(r)
    sass.sort(key=lambda r: int(r[iA], 16))
    base = int(sass[0][iA], 16)
    st = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[iN] or 0) for r in sass) or 1
    tote = sum(int(r[iE] or 0) for r in sass) or 1
    print(f"samples {tot}, warp instructions {tote}, {len(sass)} SASS instructions")
    ops = collections.Counter()
    for r in sass:
        t = r[iS].split()
        ops[(t[1] if t[0].startswith("@") else t[0]).split(".")[0]] += int(r[iN] or 0)
    print("samples by opcode:", ", ".join(f"{k} {100 * v / tot:.1f}%" for k, v in ops.most_common(12)))
    for r in sass:
        n, e = int(r[iN] or 0), int(r[iE] or 0)
        if 100 * n / tot < thr and 100 * e / tote < thr:
            continue
        s = sorted(((int(r[hdr.index(h)] or 0), h[6:]) for h in st), reverse=True)[:2]
        print(f"{int(r[iA], 16) - base:6x} {r[iS][:64]:64s} smp {100 * n / tot:5.2f}% exe {100 * e / tote:5.2f}% thr {r[iT][:4]:4s} "
              + " ".join(f"{k}={v}" for v, k in s if v))


if __name__ == "__main__":
    main()
