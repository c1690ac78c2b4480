"""Per-kernel table from an ncu report (development aid):
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [regex]"""
import csv
import re
import subprocess
import sys

WANT = ['Kernel Name', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_membar_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio',
        'lts__t_sectors_op_red.sum', 'lts__t_sectors_op_atom.sum', 'lts__t_sectors_srcunit_tex_op_read.sum',
        'lts__t_sectors_srcunit_tex_op_write.sum']
SHORT = {'launch__grid_size': 'grid', 'launch__block_size': 'block', 'launch__registers_per_thread': 'regs',
         'gpu__time_duration.sum': 'us', 'dram__bytes_read.sum': 'rdMB', 'dram__bytes_write.sum': 'wr',
         'dram__throughput.avg.pct_of_peak_sustained_elapsed': 'dram%',
         'smsp__issue_active.avg.pct_of_peak_sustained_active': 'issue%',
         'sm__warps_active.avg.pct_of_peak_sustained_active': 'warps%', 'smsp__inst_executed.sum': 'inst',
         'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum': 'bankconf',
         'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum': 'smemwave'}


def main():
    rep = sys.argv[1]
    pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = [(w, hdr.index(w)) for w in WANT if w in hdr]
    for r in rows[2:]:
        name = r[hdr.index('Kernel Name')]
        if pat and not pat.search(name):
            continue
        print(name[:70])
        parts = []
        for w, i in idx[1:]:
            label = SHORT.get(w, w.replace('smsp__average_warps_issue_stalled_', 'st_').replace('_per_issue_active.ratio', ''))
            try:
                v = float(r[i])
                sv = f"{v:.3g}" if abs(v) < 1e6 else f"{v / 1e6:.2f}M"
            except ValueError:
                sv = r[i]
            parts.append(f"{label}={sv}{'' if units[i] in ('', 'inst', '%') else units[i]}")
        print("   " + "  ".join(parts))


if __name__ == '__main__':
    main()
