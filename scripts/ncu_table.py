#!/usr/bin/env python
"""Print a markdown table of selected metrics from `ncu -i x.ncu-rep --page raw --csv` output (first kernel row).
usage: ncu -i x.ncu-rep --page raw --csv > x.csv; ncu_table.py x.csv >> profiles/x.md"""
import csv
import sys

KEYS = """Kernel Name
dram__bytes_read.sum
dram__bytes_write.sum
gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed
gpu__time_duration.sum
launch__block_size
launch__grid_size
launch__registers_per_thread
launch__shared_mem_per_block_dynamic
lts__t_sector_hit_rate.pct
lts__throughput.avg.pct_of_peak_sustained_elapsed
l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum
sm__cycles_elapsed.avg.per_second
sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active
sm__ops_path_tensor_op_utchmma_src_fp16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed
sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active
sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active
sm__throughput.avg.pct_of_peak_sustained_elapsed
sm__warps_active.avg.pct_of_peak_sustained_active
smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio
smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio
smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio
smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio
smsp__average_warps_issue_stalled_wait_per_issue_active.ratio
smsp__inst_executed.sum
smsp__issue_active.avg.pct_of_peak_sustained_active""".split("\n")


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    h, u, v = rows[0], rows[1], rows[2]
    ix = {k: i for i, k in enumerate(h)}
    print("| metric | unit | value |\n|---|---|---|")
    for k in KEYS:
        if k in ix:
            print("| %s | %s | %s |" % (k, u[ix[k]], v[ix[k]][:100]))


if __name__ == "__main__":
    main()
