"""Summarise an ncu --page raw --csv dump: one line per kernel with the metrics the roofline needs.
usage: ncu -i X.ncu-rep --page raw --csv | python tools/ncu_summary.py"""
import csv, sys
rows = list(csv.reader(sys.stdin))
hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sector_hit_rate.pct', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor.sum', 'smsp__cycles_active.avg', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'launch__grid_size', 'launch__block_size']
def find(name):
    for i, h in enumerate(hdr):
        if h == name or h.endswith('.' + name) or h.split('.', 2)[-1] == name:
            return i
    for i, h in enumerate(hdr):
        if name in h:
            return i
    return None
idx = {w: find(w) for w in want}
kn = hdr.index('Kernel Name')
for r in rows[2:]:
    if len(r) <= kn: continue
    print(r[kn][:70])
    for w in want:
        i = idx[w]
        if i is not None and i < len(r):
            print('    %-70s %s %s' % (w, r[i], units[i]))
