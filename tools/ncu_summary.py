"""Developer tool: summarise an .ncu-rep (raw page + SASS-level stall attribution) as text."""
import csv
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'lts__t_bytes.sum', 'l1tex__t_bytes.sum',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__inst_executed.sum', 'sm__inst_executed_pipe_fp64.sum', 'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_lsu.sum']


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print("====", r[idx['Kernel Name']][:110])
        for w in WANT:
            if w in idx:
                print(f"  {w:82s} {r[idx[w]]:>16s} {units[idx[w]]}")


def sass(rep, kernel, top=30):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name-base", "demangled", "-k",
                          f"regex:{kernel}"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[1]
    body = [r for r in rows[2:] if len(r) == len(hdr) and r[0] != 'Address']
    si, ns = hdr.index('Source'), hdr.index('# Samples')
    cols = {k: hdr.index(k) for k in ('stall_long_sb', 'stall_short_sb', 'stall_barrier', 'stall_mio', 'stall_math', 'stall_wait', 'stall_lg')}

    def I(x):
        try:
            return int(x)
        except Exception:
            return 0
    tot = sum(I(r[ns]) for r in body) or 1
    print(f"---- {kernel}: {tot} samples, {len(body)} SASS instructions")
    agg = {k: sum(I(r[c]) for r in body) for k, c in cols.items()}
    print("  stall totals:", {k: f"{100*v/tot:.1f}%" for k, v in agg.items()})
    for i, r in enumerate(body):
        r.append(i)
    for r in sorted(sorted(body, key=lambda r: -I(r[ns]))[:top], key=lambda r: r[-1]):
        st = " ".join(f"{k[6:]}={I(r[c])}" for k, c in cols.items() if I(r[c]) > 0.2 * I(r[ns]))
        print(f"  {r[-1]:5d} {100*I(r[ns])/tot:5.1f}%  {st:32s} | {r[si][:90]}")


if __name__ == "__main__":
    rep = sys.argv[1]
    raw(rep)
    for k in sys.argv[2:]:
        sass(rep, k)
