#!/usr/bin/env python3
"""Summarise an .ncu-rep (raw page) into the handful of metrics DESIGN.md / profiles/ quote."""
import csv, subprocess, sys
WANT = [
 'gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','dram__bytes_read.sum.per_second',
 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__throughput.avg.pct_of_peak_sustained_elapsed',
 'sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__grid_size','launch__block_size',
 'launch__shared_mem_per_block_dynamic','launch__occupancy_limit_shared_mem','launch__occupancy_limit_registers','launch__waves_per_multiprocessor',
 'smsp__issue_active.avg.pct_of_peak_sustained_active','sm__inst_executed.sum','smsp__inst_executed.sum',
 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active',
 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
 'sm__inst_executed_pipe_uniform.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
 'l1tex__data_pipe_lsu_wavefronts_mem_shared.avg.pct_of_peak_sustained_elapsed','sm__cycles_elapsed.max','smsp__cycles_active.avg','lts__t_sector_hit_rate.pct',
 'smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio','smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_membar_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio','smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_selected_per_issue_active.ratio','smsp__average_warps_issue_stalled_imc_miss_per_issue_active.ratio',
]
def main(path):
    out = subprocess.run(['ncu','-i',path,'--page','raw','--csv'],capture_output=True,text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        print('# kernel:', vals[hdr.index('Kernel Name')])
        for w in WANT:
            if w in hdr:
                i = hdr.index(w); print(f'{w:90s} {vals[i]:>16s} {units[i]}')
if __name__ == '__main__':
    main(sys.argv[1])
