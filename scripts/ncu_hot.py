"""Hot spots of an `ncu --page source --csv --print-source sass` dump: instructions with the most stall samples,
and totals per opcode class."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; ix = {n: i for i, n in enumerate(hdr)}
tot = 0; items = []; by_op = collections.Counter(); inst_by_op = collections.Counter()
for r in rows[2:]:
    if len(r) < len(hdr): continue
    s = int(r[ix['# Samples']] or 0); n = int(r[ix['Instructions Executed']] or 0)
    src = r[ix['Source']].strip(); op = src.split()[0] if not src.startswith('@') else src.split()[1]
    op = op.split('.')[0]
    tot += s; items.append((s, n, r[ix['Address']], src, r)); by_op[op] += s; inst_by_op[op] += n
print('total samples', tot, 'total warp-inst', sum(inst_by_op.values()))
print('--- by opcode: samples%, inst%')
ti = sum(inst_by_op.values())
for op, s in by_op.most_common(18):
    print('  %-10s %5.1f%%  %5.1f%%' % (op, 100.0 * s / tot, 100.0 * inst_by_op[op] / ti))
print('--- top instructions')
for pos, (s, n, addr, src, r) in enumerate(sorted(items, key=lambda t: -t[0])[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]):
    st = {k: int(r[ix[k]] or 0) for k in ('stall_long_sb', 'stall_short_sb', 'stall_barrier', 'stall_wait', 'stall_mio', 'stall_math', 'stall_lg', 'stall_not_selected')}
    top = max(st, key=st.get)
    print('  %5.2f%% inst=%9d %s  [%s %d]  %s' % (100.0 * s / tot, n, addr[-5:], top, st[top], src[:90]))
