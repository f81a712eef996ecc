"""Blackwell/Hopper async-copy evidence in the SHIPPED library, without rebuilding: counts of the TMA / bulk-copy SASS
mnemonics per kernel (cuobjdump -sass vren_b200/libvrenb200.so).  UBLKCP = cp.async.bulk (1-D TMA, both directions),
UBLKPF = cp.async.bulk.prefetch.L2, UTMALDG = cp.async.bulk.tensor (2-D TMA load), SYNCS = mbarrier operations.
usage (repo root): python tools/sass_summary.py > profiles/r2_sass_tma_counts.txt"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "vren_b200/libvrenb200.so"
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
names = {}
counts = collections.defaultdict(collections.Counter)
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    m = re.search(r"\b(UBLKCP|UBLKPF|UTMALDG|UTMASTG|UTMAPF|SYNCS|ATOMS|RED|ATOMG)\b[.\w]*", line)
    if m and cur:
        counts[cur][m.group(1)] += 1
mangled = list(counts)
demangled = subprocess.run(["c++filt"], input="\n".join(mangled), capture_output=True, text=True).stdout.splitlines()
rows = collections.defaultdict(collections.Counter)
inst = collections.Counter()
for m, d in zip(mangled, demangled):
    fam = re.sub(r"^void ", "", d)
    fam = fam.replace("vrenb200::(anonymous namespace)::", "").replace("vrenb200::", "")
    fam = re.sub(r"\(.*$", "", fam)
    fam_short = re.sub(r"<.*$", "", fam)
    for k, v in counts[m].items():
        rows[fam_short][k] += v
    inst[fam_short] += 1
cols = ["UBLKCP", "UBLKPF", "UTMALDG", "UTMASTG", "SYNCS", "ATOMS"]
print(f"# {lib}: SASS mnemonic counts summed over the template instances of every kernel that uses an async copy or an mbarrier")
print(f"{'kernel':44s} {'instances':>9s} " + " ".join(f"{c:>8s}" for c in cols))
tot = collections.Counter()
for fam in sorted(rows):
    if not any(rows[fam][c] for c in cols[:5]):
        continue
    print(f"{fam:44s} {inst[fam]:9d} " + " ".join(f"{rows[fam][c]:8d}" for c in cols))
    tot.update(rows[fam])
print(f"{'total':44s} {'':9s} " + " ".join(f"{tot[c]:8d}" for c in cols))
