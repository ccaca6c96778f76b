"""Summarise an ncu --csv launch list: per kernel name, launches, mean duration and the other metrics."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = None; data = collections.OrderedDict()
for r in rows:
    if len(r) > 10 and r[0] == 'ID': hdr = r; continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r)); data.setdefault((d['ID'], d['Kernel Name']), {})[d['Metric Name']] = float(d['Metric Value'].replace(',', ''))
agg = collections.OrderedDict()
for (i, k), m in data.items():
    name = k.split('(')[0][:44]
    a = agg.setdefault(name, collections.defaultdict(float)); a['n'] += 1
    for kk, v in m.items(): a[kk] += v
keys = [('gpu__time_duration.sum', 'us', 1e-3), ('dram__bytes_read.sum', 'rdMB', 1e-6), ('dram__bytes_write.sum', 'wrMB', 1e-6),
        ('lts__t_bytes.sum', 'L2MB', 1e-6), ('sm__warps_active.avg.pct_of_peak_sustained_active', 'occ%', 1), ('smsp__inst_executed.sum', 'Minst', 1e-6),
        ('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'Mconf', 1e-6), ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm%', 1)]
print('%-44s %4s ' % ('kernel', 'n') + ' '.join('%9s' % k[1] for k in keys))
for k, a in agg.items():
    print('%-44s %4d ' % (k, a['n']) + ' '.join('%9.1f' % (a[kk[0]] / a['n'] * kk[2]) for kk in keys))
