"""Aggregate an ncu report by CUDA source line: python tools/ncu_lines.py report.ncu-rep [top]"""
import csv, subprocess, sys, collections
def num(x):
    try: return float(x)
    except Exception: return 0.0
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fname = None; hdr = None; lines = []; kernel_seen = 0
pipes = collections.Counter()
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = r[1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < 10: continue
    if r[0] != "":   # a source line with aggregated metrics
        d = dict(file=fname.split("/")[-1], line=r[0], src=r[1].strip(), samples=num(r[6]), inst=num(r[7]), tinst=num(r[8]))
        lines.append(d)
    else:
        op = r[3].strip().split()
        if op:
            name = op[0] if not op[0].startswith("@") else (op[1] if len(op) > 1 else op[0])
            pipes[name.split(".")[0]] += num(r[7])
ti = sum(d["inst"] for d in lines); ts = sum(d["samples"] for d in lines)
print(f"total warp-inst {ti:.3e} samples {ts:.0f}")
for d in sorted(lines, key=lambda d: -d["samples"])[:top]:
    print(f"{d['file'][-16:]:16s} L{d['line']:>4s} inst={d['inst']/ti*100:5.1f}% samp={d['samples']/ts*100:5.1f}% lanes={d['tinst']/max(d['inst'],1):4.1f}  {d['src'][:100]}")
print("--- opcode mix (warp-inst %)")
tp = sum(pipes.values())
print(", ".join(f"{k}:{v/tp*100:.1f}" for k, v in pipes.most_common(40)))
