#!/usr/bin/env python
"""Side-by-side table of selected metrics from `ncu -i X.ncu-rep --page raw --csv` dumps."""
import csv
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed.sum", "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_drain_per_issue_active.ratio",
        "lts__xbar2lts_cycles_active.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.avg",
        "smsp__warps_eligible.avg.per_cycle_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts.sum",
        "sm__inst_executed_pipe_lsu.sum", "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_active",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_bytes_equiv_l1sectormiss_pipe_lsu_mem_global_op_ldgsts.sum"]


def main():
    tabs, names = {}, []
    for f in sys.argv[1:]:
        rows = list(csv.reader(open(f)))
        hdr, units = rows[0], rows[1]
        for k, vals in enumerate(rows[2:]):
            name = "%s#%d" % (f.split("/")[-1].replace(".csv", "")[-14:], k)
            kn = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else ""
            tabs[name] = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
            tabs[name]["_kernel"] = (kn[:60], "")
            names.append(name)
    print("%-74s" % "metric" + "".join("%18s" % n for n in names))
    for k in ["_kernel"] + KEYS:
        if k in tabs[names[0]]:
            print("%-74s" % k[:74] + "".join("%18s" % tabs[n].get(k, ("", ""))[0][:17] for n in names), tabs[names[0]][k][1])


if __name__ == "__main__":
    main()
