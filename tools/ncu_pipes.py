"""Per source function: executed warp instructions split by issue pipe (FMA / ALU / LSU / other) from an ncu
report with source correlation.  On Blackwell both the FMA and the ALU pipe take one warp instruction every two
cycles per scheduler, so a function whose mix is far from 50/50 is bound by the busier pipe, not by issue slots.
Usage: python tools/ncu_pipes.py report.ncu-rep [envs]"""
import csv, subprocess, sys, collections, re, os
rep = sys.argv[1]
E = float(sys.argv[2]) if len(sys.argv) > 2 else 16384.0
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FMA = {"IMAD", "FMUL", "FADD", "FFMA", "HFMA2", "HADD2", "HMUL2", "IDP"}
ALU = {"ISETP", "LOP3", "PRMT", "SEL", "SHF", "IADD3", "VIMNMX", "FMNMX", "VIADD", "FSETP", "MOV", "VIMNMX3", "PLOP3", "LEA", "FSEL", "IABS", "POPC", "FLO", "BREV", "IMNMX", "BMSK", "SGXT", "FCHK", "VABSDIFF", "P2R", "R2P", "CS2R", "FSET", "LOP", "VOTE"}
LSU = {"LDG", "STG", "LDS", "STS", "LDL", "STL", "ATOMS", "ATOMG", "RED", "LD", "ST", "LDC", "LDSM", "SHFL", "ATOM"}
def func_map(path):
    m = {}; cur = "?"
    for i, line in enumerate(open(path), 1):
        g = re.match(r"^(?:__device__|__global__|static|inline).*?\b([A-Za-z_0-9]+)\s*\(", line)
        if g and not line.startswith(" "): cur = g.group(1)
        m[i] = cur
    return m
maps = {f: func_map(os.path.join(ROOT, "torchdriveenv_b200", "csrc", f)) for f in ("tde_kernels.cuh", "tde_device.cuh")}
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
def num(x):
    try: return float(x)
    except Exception: return 0.0
fname = kern = sec = None
agg = collections.defaultdict(lambda: collections.defaultdict(lambda: collections.Counter()))
for r in csv.reader(out.splitlines()):
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": kern = "render" if "render" in r[1] else ("physics" if "physics" in r[1] else r[1][:20]); continue
    if r[0] == "Line No" or len(r) < 10: continue
    if r[0] != "":
        sec = maps.get(fname, {}).get(int(r[0]), fname); continue
    t = r[3].strip().split()
    if not t: continue
    op = (t[1] if t[0].startswith("@") and len(t) > 1 else t[0]).split(".")[0]
    pipe = "fma" if op in FMA else "alu" if op in ALU else "lsu" if op in LSU else "other"
    agg[kern][sec][pipe] += num(r[7])
    agg[kern]["TOTAL"][pipe] += num(r[7])
for k, d in agg.items():
    print(f"===== {k} (warp instructions per env; pipe-bound cycles = 2 x max(fma, alu))")
    for s, c in sorted(d.items(), key=lambda kv: -sum(kv[1].values())):
        tot = sum(c.values())
        if tot / E < 20: continue
        print(f"  {s:24s} total={tot/E:7.0f} fma={c['fma']/E:6.0f} alu={c['alu']/E:6.0f} lsu={c['lsu']/E:6.0f} other={c['other']/E:6.0f}  2*max={2*max(c['fma'],c['alu'])/E:7.0f}")
