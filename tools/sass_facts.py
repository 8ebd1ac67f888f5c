"""Static SASS facts of the in-tree library: python tools/sass_facts.py <tag>  ->  profiles/<tag>_sass_facts.md
(cuobjdump -sass on torchdriveenv_b200/libtde_b200.so; no GPU needed)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
lib = os.path.join(ROOT, "torchdriveenv_b200", "libtde_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
usage = {}
cur = None
for line in res.splitlines():
    m = re.search(r"Function (\S+):", line)
    if m:
        cur = m.group(1)
    m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+)", line)
    if m and cur:
        usage[cur] = (int(m.group(1)), int(m.group(2)), int(m.group(3)))
funcs = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); funcs[cur] = collections.Counter(); continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        funcs[cur]["total"] += 1
        for key, pat in (("STG.E.128", r"^STG\.E\.128"), ("LDG.128", r"^LDG\.E\.128"), ("LDG.CONSTANT", r"^LDG.*CONSTANT"), ("ATOMS", r"^ATOMS"),
                         ("UBLKCP", r"^UBLKCP"), ("SYNCS", r"^SYNCS"), ("BAR", r"^BAR"), ("VOTE/REDUX/SHFL", r"^(VOTE|REDUX|CREDUX|SHFL)"),
                         ("FP64 (DFMA/DMUL/F2I.S64)", r"^(DFMA|DMUL|DADD|F2I\.S64\.F64)"), ("LDS", r"^LDS"), ("LDL/STL", r"^(LDL|STL)"),
                         ("tensor (MMA)", r"^(HMMA|IMMA|DMMA|UTC.*MMA|HGMMA)")):
            if re.match(pat, op):
                funcs[cur][key] += 1
cols = ["total", "STG.E.128", "LDG.128", "LDG.CONSTANT", "LDS", "ATOMS", "VOTE/REDUX/SHFL", "BAR", "UBLKCP", "SYNCS", "FP64 (DFMA/DMUL/F2I.S64)", "LDL/STL", "tensor (MMA)"]
out = [f"# {tag}: SASS facts of the in-tree library (cuobjdump -sass / -res-usage, sm_100a)\n",
       "Static instruction counts per kernel (device functions included).  What the north star asks for and where it shows: 128-bit stores of the "
       "observation tensor (`STG.E.128`), read-only tables through the non-coherent path (`LDG...CONSTANT` = `ld.global.nc`), shared-memory staging "
       "by the TMA engine in the staged physics launch (`UBLKCP` = `cp.async.bulk`, `SYNCS` = mbarrier arrive / try_wait), warp votes / "
       "reductions in the SAT, mesh and coverage code, FP64 only in the rasteriser's edge setup, no tensor-core instructions (nothing here is a "
       "dense contraction).\n",
       "| kernel | registers | stack B | " + " | ".join(cols) + " |", "|---|---|---|" + "---|" * len(cols)]
for f, c in funcs.items():
    u = usage.get(f, ("?", "?", "?"))
    out.append(f"| `{f}` | {u[0]} | {u[1]} | " + " | ".join(str(c.get(k, 0)) for k in cols) + " |")
path = os.path.join(ROOT, "profiles", f"{tag}_sass_facts.md")
open(path, "w").write("\n".join(out) + "\n")
print(path)
