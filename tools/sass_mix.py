"""Static instruction mix of the sort pass kernel's instances (cuobjdump -sass, no GPU needed).  The ranking / regroup / write-out
loops of onesweep_pass_kernel are fully unrolled over the ITEMS rows of a thread, so (static instruction count) / ITEMS is close
to the dynamic thread-instructions per pair that ncu reports (30 per pair for the unchecked 256x48 kernel, profiles/r1y_ncu_sort_v10.md);
the table shows what the in-kernel ranking check and the ballot match add, by pipe.
usage (repo root): python tools/sass_mix.py > profiles/r2_sass_instruction_mix.txt"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "vren_b200/libvrenb200.so"
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
CLASSES = [
    ("shared ld/st", r"^(LDS|STS|LDSM)"), ("shared atomic", r"^ATOMS"), ("global ld/st", r"^(LDG|STG|LD|ST)\b|^(LDG|STG)\."), ("global atomic/red", r"^(ATOMG|RED|ATOM)\b"),
    ("TMA / mbarrier", r"^(UBLKCP|UBLKPF|UTMA|SYNCS)"), ("vote / match", r"^(VOTE|MATCH|VOTEU)"), ("shuffle", r"^SHFL"), ("barrier", r"^(BAR|WARPSYNC|BSYNC|BSSY)"),
    ("popc / flo / prmt", r"^(POPC|FLO|PRMT|BREV)"), ("logic / shift", r"^(LOP3|SHF|LEA|SEL|SGXT|PLOP3|ULOP3|USHF)"), ("int add / mul", r"^(IADD3|IMAD|IADD|UIADD3|UIMAD|VIADD|IABS)"),
    ("compare / predicate", r"^(ISETP|UISETP|PSETP|R2P|P2R|R2UR|S2R|S2UR|CS2R)"), ("branch / control", r"^(BRA|EXIT|CALL|RET|NANOSLEEP|YIELD|BRX|JMP|WARPSYNC)"),
    ("move", r"^(MOV|UMOV|IMAD\.MOV)"),
]
FLAGS = {1: "LB_INTERLEAVED", 2: "LB_STEP8", 3: "EARLY_TMA", 4: "PREFETCH_L2", 5: "RANK_LEADER", 6: "RANK_ATOMIC", 7: "VERIFY_ALL",
         8: "VERIFY_SAMPLED", 9: "REG_COUNTS", 10: "KEYS_CHUNKED", 11: "MATCH_SPLIT4", 12: "SEGMENTED", 13: "REDO"}
funcs, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        funcs[cur] = []
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        funcs[cur].append(m.group(1))
names = subprocess.run(["c++filt"], input="\n".join(funcs), capture_output=True, text=True).stdout.splitlines()
rows = []
for mangled, name in zip(funcs, names):
    m = re.search(r"onesweep_pass_kernel<(\d+), (\d+), (\d+), (\d+)u, (\d+)>", name)
    if not m:
        continue
    t, items, layout, f, occ = (int(x) for x in m.groups())
    if (t, items, layout) != (256, 48, 1) and (t, items, layout) != (256, 46, 1):
        continue                                        # the key / value tiles of the large-input sort
    if f >> 12 & 1 or f >> 13 & 1:
        continue                                        # not the segmented / repeat forms
    counts = collections.Counter()
    for op in funcs[mangled]:
        for label, pat in CLASSES:
            if re.match(pat, op):
                counts[label] += 1
                break
        else:
            counts["other"] += 1
    rows.append((items, f, len(funcs[mangled]), counts))
print(f"# {lib}: static SASS instruction mix of onesweep_pass_kernel<256, ITEMS, key/value arrays>, per pair (count / ITEMS)")
labels = [c[0] for c in CLASSES] + ["other"]
for items, f, total, counts in sorted(rows, key=lambda r: (r[1] >> 5 & 1, r[1])):
    bits = " | ".join(n for b, n in FLAGS.items() if f >> b & 1 and n in ("RANK_LEADER", "RANK_ATOMIC", "VERIFY_ALL", "VERIFY_SAMPLED"))
    print(f"\n256x{items}  {bits}: {total} instructions = {total / items:.1f} per pair")
    for label in labels:
        if counts[label]:
            print(f"    {label:22s} {counts[label]:6d}   {counts[label] / items:6.2f} per pair")
