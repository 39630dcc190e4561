"""Turns gpurun_out/*.ncu-rep (ncu --set full) and the launch-list CSV into the committed summaries.

    python profiles/summarize.py gpurun_out/k2_r1.ncu-rep gpurun_out/k3_r1.ncu-rep > profiles/r01_ncu_summary.md
"""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "sm__cycles_active.avg", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "gcc__cache_requests_type_instruction.sum", "gcc__average_cache_request_hit_rate.pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    return {h: (v, u) for h, u, v in zip(hdr, units, vals)}


def launch_table(path):
    """Per-kernel totals of a `--metrics gpu__time_duration.sum --csv` launch list (the integer-peak
    micro-benchmark launches are listed but excluded from the shares)."""
    import collections
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
    d = collections.defaultdict(list)
    for r in rows:
        d[r[4].split("(")[0].replace("void ", "")].append(float(r[-1]) / 1000)
    tot = sum(sum(v) for k, v in d.items() if k.startswith("evx_") and "peak" not in k)
    print("| kernel | launches | total µs | avg µs | share of encode kernels |\n|---|---|---|---|---|")
    for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
        share = f"{100 * sum(v) / tot:.1f}%" if k.startswith("evx_") and "peak" not in k else "(excluded)"
        print(f"| {k} | {len(v)} | {sum(v):.1f} | {sum(v) / len(v):.1f} | {share} |")
    print()


args = sys.argv[1:]
if args and args[0] == "--launches":
    launch_table(args[1])
    args = args[2:]
for rep in args:
    m = raw(rep)
    print(f"## {m.get('Kernel Name', ('?', ''))[0]}  ({rep})\n")
    print("| metric | value | unit |\n|---|---|---|")
    for k in KEYS:
        if k in m:
            print(f"| {k} | {m[k][0]} | {m[k][1]} |")
    print()
