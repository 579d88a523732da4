#!/usr/bin/env python3
"""Summarise an .ncu-rep into markdown (key counters per kernel + top source lines by stall
samples) for profiles/.   python tools/ncu_summary.py rep.ncu-rep "title" > profiles/x.md"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM written"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput % of peak"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX throughput % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit rate"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput % of peak"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads / instruction"),
    ("smsp__average_warp_latency_per_inst_issued.ratio", "warp cycles / issued instruction"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared-memory bank conflicts"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared-memory wavefronts"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem / block"),
]
STALLS = ["long_scoreboard", "wait", "short_scoreboard", "branch_resolving", "not_selected", "selected",
          "math_pipe_throttle", "no_instruction", "lg_throttle", "mio_throttle", "barrier", "dispatch_stall"]


def main():
    rep, title = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    print(f"# {title}\n")
    print(f"Source: `{rep}` (ncu --set full --clock-control none --import-source on; numbers under the "
          "profiler are cold-cache and serialised: read shares and ratios, not absolute times).\n")
    for r in rows[2:]:
        print(f"## `{r[ix['Kernel Name']]}`\n")
        print("| counter | value |\n|---|---|")
        for k, name in KEYS:
            if k in ix and r[ix[k]]:
                print(f"| {name} (`{k}`) | {r[ix[k]]} {units[ix[k]]} |")
        print("\nStall reasons (warps stalled per issued instruction, `smsp__average_warps_issue_stalled_*_per_issue_active`):\n")
        print("| " + " | ".join(STALLS) + " |\n|" + "---|" * len(STALLS))
        vals = []
        for s in STALLS:
            k = f"smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio"
            vals.append(f"{float(r[ix[k]]):.2f}" if k in ix and r[ix[k]] else "-")
        print("| " + " | ".join(vals) + " |\n")
    for kern in sorted(set(r[ix["Kernel Name"]] for r in rows[2:])):
        short = kern.split("(")[0].split("<")[0].split()[-1].split("::")[-1]
        out = subprocess.run([sys.executable, __file__.replace("ncu_summary.py", "ncu_src_hot.py"), rep, "14", short],
                             capture_output=True, text=True).stdout
        print(f"### top source lines by stall samples — `{short}`\n\n```\n{out}```\n")


if __name__ == "__main__":
    main()
