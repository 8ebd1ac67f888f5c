"""Per-source-line warp instructions of one kernel: python tools/ncu_kernel_lines.py report.ncu-rep render|physics [min_per_env] [E]"""
import csv, subprocess, sys
rep, which = sys.argv[1], sys.argv[2]
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 8.0
E = float(sys.argv[4]) if len(sys.argv) > 4 else 16384.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
fname = func = hdr = None
agg = {}
def num(x):
    try: return float(x)
    except Exception: return 0.0
for r in csv.reader(out.splitlines()):
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": func = r[1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < 10 or r[0] == "": continue
    if which in (func or ""):
        k = (fname, int(r[0]))
        a = agg.setdefault(k, [0.0, 0.0, 0.0, r[1].strip()])
        a[0] += num(r[7]); a[1] += num(r[6]); a[2] += num(r[8])
tot = sum(a[0] for a in agg.values())
print(f"{which}: {tot / E:.0f} warp-inst per env")
for (f, ln), a in sorted(agg.items()):
    if a[0] / E >= thr:
        print(f"{f[-15:]:15s} L{ln:4d} {a[0] / E:7.1f} lanes={a[2] / max(a[0], 1):4.1f} samp={a[1]:6.0f}  {a[3][:100]}")
