#!/usr/bin/env python
"""Top stalled SASS instructions of a profiled kernel: tools/ncu_source_top.py rep.ncu-rep [N] [kernel-id]"""
import csv, io, subprocess, sys
path = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"] + (["--launch-skip", sys.argv[3], "--launch-count", "1"] if len(sys.argv) > 3 else []),
                     capture_output=True, text=True).stdout
lines = out.splitlines()
# first line: Kernel Name; then header
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
print(lines[0][:150])
rows = list(csv.reader(io.StringIO("\n".join(lines[start:]))))
hdr = rows[0]
isamp = hdr.index("# Samples"); isrc = hdr.index("Source"); iex = hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
body = []
for k, r in enumerate(rows[1:]):
    if len(r) <= isamp: break
    if r[0].startswith('"Kernel') or not r[0].startswith("0x"): break
    body.append((k, r))
tot = sum(int(r[isamp]) for _, r in body)
print(f"{len(body)} instructions, {tot} samples")
for k, r in sorted(body, key=lambda kr: -int(kr[1][isamp]))[:N]:
    st = sorted(((int(r[i]), hdr[i][6:]) for i in stall_cols if r[i].isdigit() and int(r[i]) > 0), reverse=True)[:3]
    print(f"{k:5d} {int(r[isamp]):6d} {100*int(r[isamp])/tot:5.1f}% ex={r[iex]:>8s} {r[isrc].strip()[:70]:70s} {st}")
if len(sys.argv) > 4:
    B = int(sys.argv[4])
    print(f"--- buckets of {B} instructions")
    for s in range(0, len(body), B):
        seg = body[s:s + B]
        n = sum(int(r[isamp]) for _, r in seg)
        agg = {}
        for _, r in seg:
            for i in stall_cols:
                if r[i].isdigit() and int(r[i]):
                    agg[hdr[i][6:]] = agg.get(hdr[i][6:], 0) + int(r[i])
        top = sorted(agg.items(), key=lambda kv: -kv[1])[:4]
        ops = {}
        for _, r in seg:
            op = r[isrc].strip().split()[0] if not r[isrc].strip().startswith("@") else r[isrc].strip().split()[1]
            ops[op] = ops.get(op, 0) + 1
        topops = sorted(ops.items(), key=lambda kv: -kv[1])[:4]
        print(f"{s:5d} {n:6d} {100*n/tot:5.1f}% {top} {topops}")
