#!/usr/bin/env python
"""Print the judged subset of an `ncu --page raw --csv` export, one block per launch.
usage: tools/ncu_view.py file.raw.csv [substring of kernel name]"""
import csv
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__occupancy_limit_registers", "occ_lim_regs"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved_occupancy_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_throughput_pct"),
    ("sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "pipe_fmaheavy_pct"),
    ("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active", "fmaheavy_cycles_pct"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "pipe_alu_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_throughput_pct"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
    ("l1tex__t_sector_hit_rate.pct", "l1_hit_pct"),
    ("smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "stall_long_scoreboard"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall_long_sb_per_issue"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall_math_throttle_per_issue"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall_wait_per_issue"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall_barrier_per_issue"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall_lg_throttle_per_issue"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall_no_instr_per_issue"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall_short_sb_per_issue"),
    ("smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "stall_membar_per_issue"),
    ("smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "stall_dispatch_per_issue"),
]


def main():
    path = sys.argv[1]
    pat = sys.argv[2] if len(sys.argv) > 2 else ""
    rows = list(csv.reader(open(path)))
    hdr = rows[0]
    units = rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    kn = idx.get("Kernel Name")
    for r in rows[2:]:
        if len(r) < len(hdr) or (pat and pat not in r[kn]):
            continue
        print("==", r[kn][:100])
        for key, name in KEYS:
            if key in idx:
                print(f"   {name:32s} {r[idx[key]]:>16s} {units[idx[key]]}")


if __name__ == "__main__":
    main()
