#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that prove a Blackwell-native path (B200_PROFILING.md):
UTC*MMA (tcgen05.mma), LDTM/STTM (tcgen05.ld/st), UTMALDG/UTMASTG/UTMAREDG (TMA), UTCBAR (tcgen05.commit), SYNCS
(mbarrier), MUFU.  Runs in the build container:   python tools/sass_summary.py > profiles/sass_summary.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "crossscore_b200", "libcrossscore_sm100a.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
names = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
pats = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UTCBAR", "SYNCS", "MUFU", "HMMA", "FFMA2", "USETMAXREG"]
rows, cur, it = [], None, iter(names)
for line in sass.split("\n"):
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = collections.Counter()
        rows.append((next(it), cur))
        continue
    if cur is None:
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", line)
    if m:
        op = m.group(1)
        cur["total"] += 1
        for p in pats:
            if op.startswith(p):
                cur[p] += 1
                if p == "UTCHMMA" and ".2CTA" in line:
                    cur["UTCHMMA.2CTA"] += 1
print(f"# {os.path.relpath(lib, ROOT)}: SASS mnemonic counts per kernel (cuobjdump -sass, sm_100a)")
print("# kernel | instructions | " + " | ".join(pats + ["UTCHMMA.2CTA"]))
tot = collections.Counter()
for name, c in rows:
    short = name.replace("void xs::", "").replace("xs::", "")
    cut = short.find(">(")
    short = short[:cut + 1] if cut >= 0 else re.sub(r"\(.*", "", short)
    short = short.replace("(int)", "").replace("(bool)", "")
    if c["total"] == 0:
        continue
    tot.update(c)
    print(f"{short:90s} {c['total']:6d} " + " ".join(f"{c[p]:5d}" for p in pats + ["UTCHMMA.2CTA"]))
print(f"{'TOTAL':90s} {tot['total']:6d} " + " ".join(f"{tot[p]:5d}" for p in pats + ["UTCHMMA.2CTA"]))
