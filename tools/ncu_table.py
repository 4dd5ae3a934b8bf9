#!/usr/bin/env python3
"""Markdown table of the metrics this project tracks from an `ncu --set full` report (one column per captured launch):
    python tools/ncu_table.py gpurun_out/r02m/c3s_wavefront.ncu-rep > profiles/r02_ncu_wavefront_before.md"""
import csv
import io
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__registers_per_thread", "registers / thread"),
    ("sm__maximum_warps_per_active_cycle_pct", "theoretical occupancy %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active lanes / instruction"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__t_sectors_srcunit_tex_op_read.sum", "L2 sectors read (from SMs)"),
    ("lts__t_sectors_srcunit_tex_op_write.sum", "L2 sectors written (from SMs)"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit rate %"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "L1 sectors, global loads"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "L1 sectors, global stores"),
    ("l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "L1 sectors, local loads"),
    ("l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum", "L1 sectors, local stores"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall: long_scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall: lg_throttle / issue"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall: no_instruction / issue"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall: wait / issue"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall: short_scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall: mio_throttle / issue"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall: branch_resolving / issue"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall: not_selected / issue"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall: barrier / issue"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64 pipe %"),
]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    names = [r[idx["Kernel Name"]].split("(")[0].replace("void ", "").replace("<unnamed>::", "").replace("B2Variant<0, 4, 1, 3, 2, 0>", "c3").replace("B2Variant<0, 5, 1, 3, 2, 0>", "c4")
             for r in data]
    print(f"`{path}` (`ncu --set full --clock-control none --import-source on`, one column per captured launch)\n")
    print("| metric | " + " | ".join(f"`{n}`" for n in names) + " |")
    print("|---|" + "---|" * len(names))
    for key, label in METRICS:
        if key not in idx:
            continue
        i = idx[key]
        cells = []
        for r in data:
            v = r[i]
            try:
                f = float(v.replace(",", ""))
                v = f"{f:,.0f}" if abs(f) >= 1000 else f"{f:.3g}"
            except ValueError:
                pass
            cells.append(v)
        unit = units[i]
        print(f"| {label}{' (' + unit + ')' if unit and unit not in ('%', 'inst', 'sector') else ''} | " + " | ".join(cells) + " |")


if __name__ == "__main__":
    main(sys.argv[1])
