#!/usr/bin/env python3
"""Per-instruction view of one kernel of an .ncu-rep: opcode mix, stall totals and the hottest SASS lines.
    python tools/ncu_source.py rep.ncu-rep <kernel regex> [top_n]"""
import collections
import csv
import subprocess
import sys


def main():
    rep, pat = sys.argv[1], sys.argv[2]
    top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 20
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + pat],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[1]
    data = [r for r in rows[2:] if len(r) == len(hdr) and r[hdr.index("# Samples")].isdigit()]
    ia, ie, isamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    stalls = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    print(rows[0][1][:90])
    print("samples", sum(int(r[isamp]) for r in data), "instructions", sum(int(r[ie]) for r in data), "sass lines", len(data))
    agg = collections.Counter()
    for r in data:
        for i in stalls:
            agg[hdr[i]] += int(r[i])
    print("stalls:", agg.most_common(10))
    h, hs = collections.Counter(), collections.Counter()
    for r in data:
        t = r[ia].split()
        op = t[1] if t[0].startswith('@') else t[0]
        h[op] += int(r[ie]); hs[op] += int(r[isamp])
    for op, c in h.most_common(14):
        print(f"  {op:34s} {c:12d}  samples {hs[op]}")
    iw = hdr.index("L1 Wavefronts Shared") if "L1 Wavefronts Shared" in hdr else None
    if iw is not None:
        print("shared wavefronts", sum(int(r[iw]) for r in data), "ideal", sum(int(r[hdr.index('L1 Wavefronts Shared Ideal')]) for r in data))
    top = sorted(range(len(data)), key=lambda k: -int(data[k][isamp]))[:top_n]
    for k in sorted(top):
        r = data[k]
        st = {hdr[i][6:]: int(r[i]) for i in stalls if int(r[i]) > 0}
        print(k, r[ia][:60].ljust(60), r[isamp], r[ie], dict(sorted(st.items(), key=lambda x: -x[1])[:3]))


if __name__ == '__main__':
    main()
