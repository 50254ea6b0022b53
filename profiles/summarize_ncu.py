"""Condense an `ncu --page raw --csv` dump into the handful of columns the design notes cite."""
import csv
import sys

KEEP = ['Kernel Name', 'Block Size', 'Grid Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_registers', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'smsp__inst_executed.sum', 'sm__cycles_elapsed.max',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed']


def main(src, dst=None):
    rows = list(csv.reader(open(src)))
    hdr, units = rows[0], rows[1]
    keep = KEEP + [h for h in hdr if 'issue_stalled' in h and 'not_issued' not in h and 'ratio' in h]
    idx = [hdr.index(k) for k in keep if k in hdr]
    out = [[hdr[i] for i in idx], [units[i] for i in idx]] + [[r[i] for i in idx] for r in rows[2:]]
    if dst:
        with open(dst, 'w', newline='') as f:
            csv.writer(f).writerows(out)
    for r in out[2:]:
        print('-----', r[0][:90])
        stalls = []
        for h, u, v in zip(out[0][1:], out[1][1:], r[1:]):
            if 'issue_stalled' in h:
                try:
                    stalls.append((float(v), h.split('issue_stalled_')[1].split('_per_')[0]))
                except ValueError:
                    pass
            else:
                print(f"   {h:72s} {v} {u}")
        print('   top stalls:', ', '.join(f"{n}={v:.2f}" for v, n in sorted(stalls, reverse=True)[:5]))


if __name__ == '__main__':
    main(*sys.argv[1:])
