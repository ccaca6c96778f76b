"""Per-source-line totals from `ncu --page source --csv --print-source cuda,sass`: samples and warp instructions."""
import csv, sys, subprocess
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file = None; lines = []; tot_s = tot_i = 0
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path': cur_file = r[1].split('/')[-1]; continue
    if len(r) > 8 and r[0] not in ('', 'Line No') and r[2] == '-':
        s = int(r[6] or 0); n = int(r[7] or 0)
        lines.append((s, n, cur_file, r[0], r[1].strip())); tot_s += s; tot_i += n
print('total samples %d, warp-inst %d' % (tot_s, tot_i))
for s, n, f, ln, src in sorted(lines, key=lambda t: -t[0])[:int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    print('%5.1f%% smp %5.1f%% inst  %s:%s  %s' % (100.0 * s / tot_s, 100.0 * n / tot_i, f, ln, src[:100]))
