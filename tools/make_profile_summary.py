"""Summarise gpurun_out/ (bench JSON, ncu launch list, ncu full capture) into profiles/<tag>_*.  Usage:
python tools/make_profile_summary.py r01c"""
import csv, json, os, subprocess, sys, collections, shutil
tag = sys.argv[1]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)
out = [f"# {tag}: ncu summary (command: tools/gpu_round.sh)\n"]
# bench lines
for name in ("bench.json", "bench_ref.json"):
    fp = os.path.join(G, name)
    if os.path.exists(fp) and os.path.getsize(fp):
        shutil.copy(fp, os.path.join(P, f"{tag}_{name}"))
        d = json.loads(open(fp).read().strip().splitlines()[-1])
        out.append(f"## {name}\n\n```\nvalue {d.get('value'):.4g} {d.get('unit')}  ms/step {d.get('ms_per_step'):.4g}  e2e {d.get('e2e', {}).get('value'):.4g}\n"
                   f"roofline {json.dumps(d.get('roofline'))}\ncpu_baseline {json.dumps(d.get('cpu_baseline'))}\nclocks {json.dumps(d.get('clocks'))}\n```\n")
# launch list
fp = os.path.join(G, "launches.csv")
if os.path.exists(fp):
    shutil.copy(fp, os.path.join(P, f"{tag}_launches.csv"))
    rows = [r for r in csv.reader(open(fp)) if len(r) > 10]
    hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value")
    d = collections.defaultdict(list)
    for r in rows[1:]:
        d[r[ki]].append(float(r[vi].replace(",", "")))
    tot = sum(sum(v) for v in d.values())
    out.append("## launch list (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised: compare shares)\n\n| kernel | launches | avg us | share |\n|---|---|---|---|")
    for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
        out.append(f"| `{k[:70]}` | {len(v)} | {sum(v)/len(v)/1e3:.1f} | {sum(v)/tot:.3f} |")
    out.append("")
# full capture
rep = os.path.join(G, "prof_step.ncu-rep")
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines())); hdr = rows[0]
    want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "smsp__issue_active.avg.per_cycle_active",
            "smsp__warps_active.avg.per_cycle_active", "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__registers_per_thread",
            "launch__grid_size", "launch__block_size", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
    out.append("## ncu --set full (one launch of each kernel, C3 workload)\n\n| metric | " + " | ".join(r[hdr.index("Kernel Name")].replace("void ", "")[:28] for r in rows[2:]) + " |\n|---|" + "---|" * len(rows[2:]))
    units = rows[1]
    for w in want[1:]:
        if w in hdr:
            i = hdr.index(w)
            out.append(f"| {w} [{units[i]}] | " + " | ".join(r[i] for r in rows[2:]) + " |")
    out.append("\nwarp stall reasons (pc sampling, % of samples):\n")
    for r in rows[2:]:
        v = []
        for i, h in enumerate(hdr):
            if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h:
                try: v.append((float(r[i]), h.replace("smsp__pcsamp_warps_issue_stalled_", "")))
                except ValueError: pass
        t = sum(x for x, _ in v) or 1
        out.append(f"* `{r[hdr.index('Kernel Name')][:40]}`: " + ", ".join(f"{h} {x/t*100:.1f}" for x, h in sorted(v, reverse=True)[:8]))
    sec = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_sections.py"), rep], capture_output=True, text=True).stdout
    out.append("\nwarp instructions by source function (from the SASS/source correlation):\n\n```\n" + sec + "```\n")
    traffic = {}
    tp = os.path.join(P, "traffic.json")
    if os.path.exists(tp): traffic = json.load(open(tp))
    i_r, i_w, i_k = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("Kernel Name")
    u_r, u_w = units[i_r], units[i_w]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tot = 0.0
    for r in rows[2:]:
        tot += float(r[i_r]) * scale.get(u_r, 1) + float(r[i_w]) * scale.get(u_w, 1)
    traffic["c3"] = tot
    traffic["c3_note"] = f"{tag}: dram__bytes_read.sum + dram__bytes_write.sum summed over one physics + one render launch (one env step of 16384 envs)"
    json.dump(traffic, open(tp, "w"), indent=1)
open(os.path.join(P, f"{tag}_summary.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out)[:3000])
