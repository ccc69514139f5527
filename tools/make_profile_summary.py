"""Developer tool: turn gpurun_out/prof_<tag>.ncu-rep + launches_<tag>.csv into profiles/<name>_*.{md,csv,json}."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, name, workload = sys.argv[1], sys.argv[2], sys.argv[3]
rep = os.path.join(ROOT, "gpurun_out", f"prof_{tag}.ncu-rep")
launches = os.path.join(ROOT, "gpurun_out", f"launches_{tag}.csv")
out_md = os.path.join(ROOT, "profiles", f"{name}_ncu_summary.md")
os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)

METRICS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
           'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
           'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
           'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
           'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
           'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
           'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
           'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
           'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'smsp__inst_executed.sum',
           'l1tex__m_xbar2l1tex_read_bytes.sum', 'l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct', 'lts__t_sector_op_read_hit_rate.pct',
           'launch__shared_mem_config_size']
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
lines = [f"# ncu summary `{name}` ({workload}, one B200, `ncu --set full --clock-control none`)", "",
         "Source: `tools/gpu_prof2.sh` / `tools/gpu_r3_prof.sh` -> `gpurun_out/prof_%s.ncu-rep`; values per launch.  Times under ncu are cold-cache and" % tag,
         "serialised: compare SHARES with bench.py's CUDA-event timings, not absolutes.", ""]
traffic = {}
seen = set()
for r in rows[2:]:
    kn = r[idx['Kernel Name']]
    short = [k for k in ("P1FBody", "P3FBody", "P5FBody", "PCFBody", "TanChain", "CotChain", "P1MBody", "P1Body", "P3Body", "P5Body", "PCBody", "SegSum",
                         "ScanApplyBody<double, Vjp", "ScanApplyBody<double, Jvp", "ScanAggBody<double, Vjp", "ScanAggBody<double, Jvp") if k in kn]
    if not short or short[0] in seen:
        continue
    seen.add(short[0])
    short[0] = short[0].replace("Body<double, ", "_")
    lines += [f"## {kn}", "", "| metric | value | unit |", "|---|---|---|"]
    for m in METRICS:
        if m in idx:
            lines.append(f"| `{m}` | {r[idx[m]]} | {units[idx[m]]} |")
    lines.append("")
    def tobytes(v, u):
        v = float(v.replace(",", ""))
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    traffic[short[0]] = tobytes(r[idx['dram__bytes_read.sum']], units[idx['dram__bytes_read.sum']]) + \
        tobytes(r[idx['dram__bytes_write.sum']], units[idx['dram__bytes_write.sum']])
if os.path.exists(launches):
    lines += ["## launch list (gpu__time_duration.sum, one product = the kernels between two first-pass launches)", "", "```"]
    txt = [l for l in open(launches).read().splitlines() if l and not l.startswith("==")]
    rd = list(csv.reader(txt))
    h = rd[0]
    ki, vi = h.index("Kernel Name"), h.index("Metric Value")
    n = 0
    for r in rd[1:]:
        if len(r) > vi:
            lines.append(f"{r[vi]:>12s} ns  {r[ki][:110]}")
            n += 1
            if n >= 32:
                break
    lines += ["```", ""]
open(out_md, "w").write("\n".join(lines))
tj = os.path.join(ROOT, "profiles", "traffic.json")
allt = json.load(open(tj)) if os.path.exists(tj) else {}
allt[workload] = traffic
allt["_source"] = f"ncu --set full, profiles/{name}_ncu_summary.md (dram__bytes_read.sum + dram__bytes_write.sum per launch)"
json.dump(allt, open(tj, "w"), indent=1)
print("wrote", out_md, traffic)
