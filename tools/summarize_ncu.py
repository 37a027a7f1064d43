#!/usr/bin/env python
"""Turn an ncu report of the render kernel into the committed summaries under profiles/ (runs here, no GPU).

    python tools/summarize_ncu.py gpurun_out/prof_r02_render_4k.ncu-rep <steps> <tag> [<summary.json>]
"""
import collections
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
rep, steps, tag = sys.argv[1], int(sys.argv[2]), sys.argv[3]
# optional 4th argument: where the summary goes (default profiles/ncu_summary.json = the render kernel's, read by bench.py)
summary_path = sys.argv[4] if len(sys.argv) > 4 else os.path.join(ROOT, "profiles", "ncu_summary.json")


def page(name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


raw = page("raw")
d = {h: (u, v) for h, u, v in zip(raw[0], raw[1], raw[2])}


def num(k):
    return float(d[k][1].replace(",", ""))


def nbytes(k):
    u, v = d[k]
    return float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]


src = page("source")
hdr, data = src[1], src[2:]
ia, ie, it = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
tot = thr = 0
byop, thr_by = collections.Counter(), collections.Counter()
for r in data:
    try:
        n, t = int(r[ie]), int(r[it])
    except (ValueError, IndexError):
        continue
    op = re.sub(r"^@!?U?P\d+\s+", "", r[ia].strip()).split()[0].split(".")[0]
    tot += n; thr += t; byop[op] += n; thr_by[op] += t
fp = sum(byop[k] for k in ("DFMA", "DMUL", "DADD", "DSETP"))
flops = 2 * thr_by["DFMA"] + thr_by["DMUL"] + thr_by["DADD"]
stalls = {h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""): float(d[h][1])
          for h in raw[0] if "issue_stalled" in h and h.endswith("per_issue_active.ratio")}
import bench  # noqa: E402  (csrc_sha: the kernel sources this capture was made from -- run right after the capture)
summ = {
    "csrc_sha": bench.csrc_sha(),
    "dram_bytes_read": nbytes("dram__bytes_read.sum"), "dram_bytes_written": nbytes("dram__bytes_write.sum"),
    "source": f"profiles/{tag}_details.csv (ncu --set full --clock-control none, {d['Kernel Name'][1]}, {steps} geodesic steps)",
    "dram_bytes_per_launch": nbytes("dram__bytes_read.sum") + nbytes("dram__bytes_write.sum"),
    "fp64_pipe_pct": num("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
    "issue_active_pct": num("smsp__issue_active.avg.pct_of_peak_sustained_active"),
    "achieved_occupancy_pct": num("sm__warps_active.avg.pct_of_peak_sustained_active"),
    "registers_per_thread": num("launch__registers_per_thread"),
    "grid_size": num("launch__grid_size"), "block_size": num("launch__block_size"),
    "warp_execution_efficiency_threads_per_inst": num("smsp__thread_inst_executed_per_inst_executed.ratio"),
    "executed_warp_instructions_per_warp_step": tot / (steps / 32),
    "executed_fp64_warp_instructions_per_warp_step": fp / (steps / 32),
    "executed_flop_per_step": flops / steps,
    "kernel_ms_under_ncu": num("gpu__time_duration.sum"),
    "sm_ghz_under_ncu": num("sm__cycles_elapsed.avg.per_second"),
    "top_stalls_per_issue": dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:5]),
}
json.dump(summ, open(summary_path, "w"), indent=1)
with open(os.path.join(ROOT, "profiles", f"{tag}_opcode_mix.txt"), "w") as f:
    f.write(f"{d['Kernel Name'][1]}, {steps} geodesic steps (ncu source page of {os.path.basename(rep)})\n")
    f.write(f"executed warp instructions {tot} = {tot / (steps / 32):.1f} per warp-step; FP64-pipe {fp / (steps / 32):.1f} per warp-step; "
            f"avg active threads/instr {thr / tot:.2f}; executed flops/step {flops / steps:.1f}\n\n")
    for k, v in byop.most_common():
        f.write(f"{k:<10} {v:>14} {100 * v / tot:6.2f}%\n")
det = subprocess.run(["ncu", "-i", rep, "--page", "details", "--csv"], capture_output=True, text=True).stdout
open(os.path.join(ROOT, "profiles", f"{tag}_details.csv"), "w").write(det)
print(json.dumps(summ, indent=1))
