"""Per-slice / per-orientation DRAM traffic and kernel shares from an ncu launch list of bench.py.
usage: python scripts/traffic_from_launches.py launches.csv passes [out.json]"""
import collections, csv, json, sys
rows = list(csv.reader(open(sys.argv[1])))
passes = int(sys.argv[2])
hdr = None; data = collections.OrderedDict()
for r in rows:
    if len(r) > 10 and r[0] == 'ID': hdr = r; continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r)); data.setdefault((d['ID'], d['Kernel Name']), {})[d['Metric Name']] = float(d['Metric Value'].replace(',', ''))
tot = collections.defaultdict(lambda: collections.defaultdict(float))
for (i, k), m in data.items():
    name = k.split('(')[0].replace('void ', '')
    for kk, v in m.items(): tot[name][kk] += v
    tot[name]['n'] += 1
f1 = next(v for k, v in tot.items() if k.startswith('slice_rows_fused'))
f2 = next(v for k, v in tot.items() if k.startswith('slice_cols_fused'))
det = next(v for k, v in tot.items() if k.startswith('detector_affine_kernel'))
byt = lambda t: t['dram__bytes_read.sum'] + t['dram__bytes_write.sum']
slices, orient = 1800 * passes, 360 * passes
stage_a = sum(t['gpu__time_duration.sum'] for n, t in tot.items()
              if not n.startswith(('detector', 'affine', 'rotate_points', 'minmax', 'row_hist', 'row_scan', 'row_scatter',
                                   'extreme', 'hull', 'species', 'at::', 'at_cuda')))
print('F1 %.2f us/slice, F2 %.2f us/slice' % (f1['gpu__time_duration.sum'] / 1e3 / slices, f2['gpu__time_duration.sum'] / 1e3 / slices))
print('DRAM per slice: F1 %.1f MB, F2 %.1f MB' % (byt(f1) / slices / 1e6, byt(f2) / slices / 1e6))
print('stage-A kernels %.2f ms/pass, F1 share %.3f, F1+F2 share %.3f' % (stage_a / 1e6 / passes, f1['gpu__time_duration.sum'] / stage_a,
      (f1['gpu__time_duration.sum'] + f2['gpu__time_duration.sum']) / stage_a))
print('detector %.1f us/pass, DRAM %.1f MB/pass' % (det['gpu__time_duration.sum'] / 1e3 / passes, byt(det) / passes / 1e6))
for n, t in sorted(tot.items(), key=lambda kv: -kv[1]['gpu__time_duration.sum'])[:12]:
    print('  %-40s n=%4d  %9.1f us/launch  %8.2f ms/pass' % (n[:40], t['n'], t['gpu__time_duration.sum'] / 1e3 / t['n'], t['gpu__time_duration.sum'] / 1e6 / passes))
if len(sys.argv) > 3:
    src = "profiles/%s (ncu launch list of `bench.py --steps 2 --warmup 1`: dram__bytes_read+write, 64 slices per fused launch pair)" % sys.argv[1].split('/')[-1]
    json.dump({"fused": {"dram_bytes_per_slice": (byt(f1) + byt(f2)) / slices, "source": src},
               "detector_affine": {"dram_bytes_per_orientation": byt(det) / orient, "source": src}}, open(sys.argv[3], 'w'), indent=1)
