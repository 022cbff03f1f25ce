"""Summarise an `ncu --page raw --csv` export: one block per kernel launch with the counters DESIGN.md argues from.
usage: ncu -i X.ncu-rep --page raw --csv > X_raw.csv ; python tools/ncu_summary.py X_raw.csv"""
import csv
import sys

KEYS = [
    ("gpu__time_duration.sum", "time"),
    ("launch__grid_size", "grid"),
    ("launch__registers_per_thread", "regs"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem"),
    ("sm__warps_active.avg.per_cycle_active", "warps/SM"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe active %"),
    ("sm__ops_path_tensor_src_tf32_dst_fp32.avg.pct_of_peak_sustained_elapsed", "tf32 ops % of peak"),
    ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "LSU data-pipe wavefronts %"),
    ("l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "tensor smem-read wavefronts %"),
    ("l1tex__m_xbar2l1tex_read_bytes.sum", "L2->L1 bytes"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput %"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait / issue"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall lg_throttle / issue"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall mio_throttle / issue"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_tex_throttle_per_issue_active.ratio", "stall tex_throttle / issue"),
    ("smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "stall membar / issue"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "L1 global load requests"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "L1 global load sectors"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "LSU smem wavefronts"),
    ("l1tex__data_pipe_lsu_wavefronts.sum", "LSU wavefronts"),
    ("sm__cycles_elapsed.max", "SM cycles"),
]
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        name = r[idx["Kernel Name"]]
        print("==", name[:110])
        vals = {}
        for k, label in KEYS:
            if k in idx and r[idx[k]] != "":
                vals[k] = (float(r[idx[k]].replace(",", "")), units[idx[k]])
                print(f"   {label:34s} {r[idx[k]]:>16s} {units[idx[k]]}")
        t = vals.get("gpu__time_duration.sum")
        b = vals.get("l1tex__m_xbar2l1tex_read_bytes.sum")
        d0, d1 = vals.get("dram__bytes_read.sum"), vals.get("dram__bytes_write.sum")
        if t and b:
            sec = t[0] * UNIT.get(t[1], 1.0)
            print(f"   {'L2->L1 GB/s':34s} {b[0] * UNIT.get(b[1], 1.0) / sec / 1e9:16.0f}")
            if d0 and d1:
                print(f"   {'dram GB/s':34s} {(d0[0] * UNIT.get(d0[1], 1.0) + d1[0] * UNIT.get(d1[1], 1.0)) / sec / 1e9:16.0f}")


if __name__ == "__main__":
    main(sys.argv[1])
