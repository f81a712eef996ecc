"""Is the device code of the current library the device code a given commit built?  Builds `commit`'s CUDA sources in a
temporary directory (nvcc cross-compiles without a GPU) and compares the SASS instruction stream of every kernel with the
current vren_b200/libvrenb200.so.  Used at the end of round 2, after the GPU budget was spent, to show that everything changed
since the last GPU run left the verified kernels byte-identical.
usage (repo root): python tools/sass_identity.py 1fa3cdc > profiles/r2_sass_identity_vs_gpu_verified.txt"""
import collections
import re
import subprocess
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
commit = sys.argv[1]


def kernels(lib):
    sass = subprocess.run(["cuobjdump", "-sass", str(lib)], capture_output=True, text=True, check=True).stdout
    funcs, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = []
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(.*?);", line)
        if m and cur:
            funcs[cur].append(re.sub(r"\s+", " ", m.group(1)))
    names = subprocess.run(["c++filt"], input="\n".join(funcs), capture_output=True, text=True).stdout.splitlines()
    out = {}
    for mangled, name in zip(funcs, names):
        name = re.sub(r"\(anonymous namespace\)::", "", name)
        name = re.sub(r"\(.*", "", name).replace("void vrenb200::", "").replace("vrenb200::", "")
        # a template argument added since (default behaviour = false): the same kernel
        name = re.sub(r"exclusive_scan_u32_kernel<(\d+)>", r"exclusive_scan_u32_kernel<\1, false>", name)
        out[name] = funcs[mangled]
    return out


with tempfile.TemporaryDirectory() as tmp:
    tar = subprocess.run(["git", "archive", commit, "vren_b200", "include"], cwd=ROOT, capture_output=True, check=True).stdout
    subprocess.run(["tar", "-x", "-C", tmp], input=tar, check=True)
    subprocess.run([sys.executable, "-c", "from vren_b200 import build; build.build_cuda(force=True)"], cwd=tmp, check=True, capture_output=True)
    old = kernels(Path(tmp) / "vren_b200" / "libvrenb200.so")
new = kernels(ROOT / "vren_b200" / "libvrenb200.so")
head = subprocess.run(["git", "rev-parse", "--short", "HEAD"], cwd=ROOT, capture_output=True, text=True).stdout.strip()
same = [k for k in old if k in new and old[k] == new[k]]
different = [k for k in old if k in new and old[k] != new[k]]
missing = [k for k in old if k not in new]
added = [k for k in new if k not in old]
print(f"# device code of the working tree (HEAD {head}) against commit {commit}, kernel by kernel (SASS instruction streams, sm_100a)")
print(f"kernels built by {commit}: {len(old)}")
print(f"  byte-identical now:     {len(same)}")
print(f"  different now:          {len(different)} {different}")
print(f"  no longer built:        {len(missing)} {missing}")
print(f"kernels added since:      {len(added)}")
for k in added:
    print(f"  + {k}   ({len(new[k])} instructions)")
sys.exit(0 if not different and not missing else 1)
