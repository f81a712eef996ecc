"""Registers, static shared memory, stack and local memory per kernel of the SHIPPED library, without a GPU
(cuobjdump --dump-resource-usage vren_b200/libvrenb200.so).  Template instances of one kernel are folded into one row
(min-max over the instances); STACK / LOCAL > 0 means spills or local arrays.
usage (repo root): python tools/resource_summary.py > profiles/r2_resource_usage.txt"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "vren_b200/libvrenb200.so"
text = subprocess.run(["cuobjdump", "--dump-resource-usage", lib], capture_output=True, text=True, check=True).stdout
funcs = re.findall(r"Function (\S+):\n\s*REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", text)
demangled = subprocess.run(["c++filt"], input="\n".join(f[0] for f in funcs), capture_output=True, text=True).stdout.splitlines()
rows = collections.defaultdict(list)
instances = []          # of the sort pass kernel: (template arguments, REG, STACK)
for (_, reg, stack, shared, local), d in zip(funcs, demangled):
    full = re.sub(r"^void ", "", d).replace("vrenb200::(anonymous namespace)::", "").replace("vrenb200::", "")
    fam = re.sub(r"[<(].*$", "", full)
    rows[fam].append((int(reg), int(stack), int(shared), int(local)))
    m = re.match(r"onesweep_pass_kernel<(\d+), (\d+), (\d+), (\d+)u, (\d+)>", full)
    if m:
        instances.append((tuple(int(x) for x in m.groups()), int(reg), int(stack)))


def span(vals):
    return str(min(vals)) if min(vals) == max(vals) else f"{min(vals)}-{max(vals)}"


print(f"# {lib}: per-kernel resources (sm_100a), template instances folded (min-max); dynamic shared memory is not listed here")
print(f"{'kernel':46s} {'instances':>9s} {'REG':>9s} {'STACK':>9s} {'SHARED':>12s} {'LOCAL':>7s}")
for fam in sorted(rows):
    r = rows[fam]
    print(f"{fam:46s} {len(r):9d} {span([x[0] for x in r]):>9s} {span([x[1] for x in r]):>9s} {span([x[2] for x in r]):>12s} {span([x[3] for x in r]):>7s}")

# the sort pass kernel, instance by instance: <THREADS, ITEMS, LAYOUT (0 keys, 1 key/value arrays, 2 interleaved uvec2), F, MIN_BLOCKS>
FLAGS = {1: "LB_INTERLEAVED", 2: "LB_STEP8", 3: "EARLY_TMA", 4: "PREFETCH_L2", 5: "RANK_LEADER", 6: "RANK_ATOMIC", 7: "VERIFY_ALL",
         8: "VERIFY_SAMPLED", 9: "REG_COUNTS", 10: "KEYS_CHUNKED", 11: "MATCH_SPLIT4", 12: "SEGMENTED", 13: "REDO"}
print()
print("# onesweep_pass_kernel<THREADS, ITEMS, LAYOUT, F, MIN_BLOCKS>: LAYOUT 0 = keys, 1 = key / value arrays, 2 = interleaved uvec2")
print(f"{'threads x items':>15s} {'layout':>6s} {'occ':>3s} {'REG':>4s} {'STACK':>5s}  option bits")
for (t, i, layout, f, occ), reg, stack in sorted(instances):
    bits = " | ".join(name for b, name in FLAGS.items() if f >> b & 1)
    print(f"{t:>9d} x {i:<3d} {layout:6d} {occ:3d} {reg:4d} {stack:5d}  {bits}")
