"""Section-level instruction breakdown of an ncu report (functions via // SECTION markers are not needed:
sections are inferred from the enclosing function names found by scanning the source files)."""
import csv, subprocess, sys, collections, re, os
rep = sys.argv[1]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# optional third argument: the csrc directory the profiled library was built from (default: the working tree)
CSRC = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "torchdriveenv_b200", "csrc")
def func_map(path):
    """line -> name of the enclosing top-level function/kernel (crude: last line matching a definition)."""
    m = {}; cur = "?"
    for i, line in enumerate(open(path), 1):
        g = re.match(r"^(?:template.*\n)?(?:__device__|__global__|static|inline).*?\b([A-Za-z_0-9]+)\s*\(", line)
        if g and not line.startswith(" "): cur = g.group(1)
        m[i] = cur
    return m
maps = {f: func_map(os.path.join(CSRC, f)) for f in ("tde_kernels.cuh", "tde_device.cuh", "tde_render.cuh") if os.path.exists(os.path.join(CSRC, f))}
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
def num(x):
    try: return float(x)
    except Exception: return 0.0
fname = func = None; agg = collections.defaultdict(lambda: collections.defaultdict(lambda: [0, 0, 0]))
for r in csv.reader(out.splitlines()):
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": func = "render" if "render" in r[1] else ("physics" if "physics" in r[1] else r[1][:20]); continue
    if r[0] == "Line No" or len(r) < 10 or r[0] == "": continue
    sec = maps.get(fname, {}).get(int(r[0]), fname)
    if fname == "tde_kernels.cuh" and sec in ("raster_band",):
        ln = int(r[0]); src = r[1]
        sec = "raster_band"
    a = agg[func][sec]; a[0] += num(r[7]); a[1] += num(r[6]); a[2] += num(r[8])
E = float(sys.argv[2]) if len(sys.argv) > 2 else 16384.0
for k, d in agg.items():
    tot = sum(v[0] for v in d.values()); ts = sum(v[1] for v in d.values())
    print(f"===== {k}: warp-inst {tot:.3e} ({tot/E:.0f}/env) samples {ts:.0f}")
    for s, v in sorted(d.items(), key=lambda kv: -kv[1][0]):
        print(f"  {s:24s} inst={v[0]/tot*100:5.1f}% samp={v[1]/max(ts,1)*100:5.1f}% lanes={v[2]/max(v[0],1):4.1f} per-env={v[0]/E:7.0f}")
