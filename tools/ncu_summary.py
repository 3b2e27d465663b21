#!/usr/bin/env python
"""Summarise one kernel of an .ncu-rep (ncu -i rep --page raw --csv) into the few lines the roofline discussion needs."""
import csv, subprocess, sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__cycles_elapsed.avg.per_second", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__inst_executed.sum",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "smsp__warps_issue_stalled_long_scoreboard_per_warp_active.pct"]


def main(rep, header=""):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    if header:
        print(header)
    for r in rows[2:]:
        d = dict(zip(hdr, zip(r, units)))
        print("kernel:", d.get("Kernel Name", ("?",))[0][:110])
        for k in WANT:
            hit = k if k in d else next((h for h in d if h.endswith("." + k)), None)      # some sections prefix the metric name
            if hit and d[hit][0] != "":
                print(f"{k:100s} {d[hit][0]:>16s} {d[hit][1]}")
        print()


if __name__ == "__main__":
    main(sys.argv[1], " ".join(sys.argv[2:]))
