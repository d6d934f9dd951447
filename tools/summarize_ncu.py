"""Summarises one profiling round (tools/profile_round.sh <tag>) from gpurun_out/ into profiles/<tag>_*.{md,csv,json}.
Run here (no GPU needed): python tools/summarize_ncu.py <tag> [frame_launch_offset]"""
import csv
import json
import os
import shutil
import subprocess
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)

lines = [l for l in open(os.path.join(G, f"launches_{tag}.csv")) if not l.startswith("==")]
rows = list(csv.DictReader(lines))
shutil.copy(os.path.join(G, f"launches_{tag}.csv"), os.path.join(P, f"{tag}_launches.csv"))
# steady-state frames: everything after the first cells_kernel (the stats frame) up to the last complete frame
names = [r["Kernel Name"].split("(")[0].replace("void ", "") for r in rows]
ends = [i for i, n in enumerate(names) if n.startswith("cells_kernel")]
frames = [(ends[k] + 1, ends[k + 1] + 1) for k in range(len(ends) - 1)]
agg = OrderedDict()
for a, b in frames[1:]:
    for r, n in zip(rows[a:b], names[a:b]):
        d = agg.setdefault(n, [0, 0.0])
        d[0] += 1
        d[1] += float(r["Metric Value"])
nf = max(1, len(frames) - 1)
tot = sum(v[1] for v in agg.values())
md = [f"# {tag}: ncu launch list (`--metrics gpu__time_duration.sum --clock-control none`), {nf} steady-state frames of bench.py", "",
      "Per-launch times under ncu are cold-cache and serialised: read the SHARES.", "",
      "| kernel | launches/frame | ns/frame | share |", "|---|---|---|---|"]
for n, (cnt, ns) in agg.items():
    md.append(f"| {n} | {cnt / nf:.1f} | {ns / nf:.0f} | {100 * ns / tot:.1f}% |")
md.append(f"| total | {sum(v[0] for v in agg.values()) / nf:.1f} | {tot / nf:.0f} | 100% |")

rep = os.path.join(G, f"prof_{tag}.ncu-rep")
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    hdr, units = rr[0], rr[1]
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "launch__registers_per_thread",
            "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "smsp__issue_active.avg.pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
            "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__cycles_elapsed.max"]
    want += [h for h in hdr if "issue_stalled" in h and h.endswith("per_warp_active.pct")]
    idx = [(w, hdr.index(w)) for w in want if w in hdr]
    seen = OrderedDict()
    for r in rr[2:]:
        k = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "")
        seen.setdefault(k, []).append({w: r[i] + " " + units[i] for w, i in idx})
    md += ["", f"## ncu --set full (one capture per kernel, `--clock-control none`)", ""]
    js = {}
    for k, caps in seen.items():
        c = caps[-1]
        js[k] = caps
        md.append(f"### {k}  ({len(caps)} captures; last shown)")
        md += [f"- {w}: {v}" for w, v in c.items()]
        md.append("")
    json.dump(js, open(os.path.join(P, f"{tag}_ncu_full.json"), "w"), indent=1)
    # roofline.traffic of bench.py: DRAM bytes per launch, valid for exactly this build of the library and this workload
    import hashlib
    to_b = lambda v: float(v.split()[0]) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[v.split()[1]]
    lib = os.path.join(ROOT, "yetanotherconsolegameengine_b200", "libycge.so")
    sys.path.insert(0, ROOT)
    from yetanotherconsolegameengine_b200 import api as _api
    traffic = {"lib_sha256": hashlib.sha256(open(lib, "rb").read()).hexdigest(), "source_sha256": _api.library_source_digest(), "workload": "dragon 480x135 ss=4", "captured": f"profiles/{tag}_ncu_full.json, ncu --set full --clock-control none of `python bench.py --steps 2 --warmup 3`",
               "kernels": {k: {"dram_bytes": to_b(c[-1]["dram__bytes_read.sum"]) + to_b(c[-1]["dram__bytes_write.sum"])} for k, c in js.items() if "dram__bytes_read.sum" in c[-1]}}
    json.dump(traffic, open(os.path.join(P, "ncu_traffic.json"), "w"), indent=1)
for f in (f"bench_{tag}.json", f"clocks_{tag}.csv"):
    if os.path.exists(os.path.join(G, f)):
        shutil.copy(os.path.join(G, f), os.path.join(P, f"{tag}_{f.split('_')[0]}{os.path.splitext(f)[1]}"))
open(os.path.join(P, f"{tag}_summary.md"), "w").write("\n".join(md) + "\n")
print("\n".join(md[:30]))
