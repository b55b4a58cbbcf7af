"""Turns a .ncu-rep (read with `ncu -i ... --page raw --csv`) into the short text summaries committed here."""
import csv
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__cycles_active.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.per_cycle_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic',
        'launch__shared_mem_per_block_static', 'launch__grid_size', 'launch__block_size', 'launch__waves_per_multiprocessor',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_warps',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']


def main(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print('kernel:', r[hdr.index('Kernel Name')])
        for k in KEYS:
            if k in hdr:
                print(f'  {k:75s} {r[hdr.index(k)]:>18s} {units[hdr.index(k)]}')
        stalls = sorted(((float(r[i]), h.split('issue_stalled_')[1].split('_per_')[0]) for i, h in enumerate(hdr)
                         if 'smsp__average_warps_issue_stalled' in h and h.endswith('_per_issue_active.ratio') and r[i]), reverse=True)
        print('  warp stall cycles per issued instruction: ' + ', '.join(f'{n} {v:.2f}' for v, n in stalls[:8]))
        print()


if __name__ == '__main__':
    main(sys.argv[1])
