"""Key metrics of an `ncu --page raw --csv` export, one block per captured launch:
python scripts/ncu_keymetrics.py report.raw.csv [more.raw.csv ...]"""
import csv, sys
KEYS = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'launch__waves_per_multiprocessor',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'l1tex__m_xbar2l1tex_read_bytes.sum', 'l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum', 'sm__cycles_elapsed.avg',
        'smsp__sass_inst_executed_op_tma_ld.sum']
for path in sys.argv[1:]:
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    stalls = [i for i, h in enumerate(hdr) if 'pcsamp_warps_issue_stalled' in h and 'not_issued' not in h]
    for r in rows[2:]:
        print('== ' + r[hdr.index('Kernel Name')][:110])
        for k in KEYS:
            if k in hdr:
                print('   %-72s %-16s %s' % (k, units[hdr.index(k)], r[hdr.index(k)]))
        top = sorted(((float(r[i].replace(',', '') or 0), hdr[i].split('stalled_')[1]) for i in stalls), reverse=True)[:5]
        print('   top stall samples: ' + ', '.join('%s=%d' % (n, v) for v, n in top))
        print()
