"""Top sampled instructions of an `ncu -i X.ncu-rep --page source --csv` dump (stall samples per SASS line)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
i_s, i_src, i_ex = hdr.index('# Samples'), hdr.index('Source'), hdr.index('Instructions Executed')
stall = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
data = []
for k, r in enumerate(rows[2:]):
    try:
        data.append((int(r[i_s]), k, r))
    except (ValueError, IndexError):
        pass
tot = sum(d[0] for d in data)
print('total samples', tot, 'instructions', len(data), 'executed', sum(int(d[2][i_ex]) for d in data))
for n, k, r in sorted(data, key=lambda x: -x[0])[:int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    top = sorted(((int(r[i]), hdr[i][6:]) for i in stall if r[i] not in ('', '0')), reverse=True)[:2]
    print('%5d %5.1f%% #%4d ex=%7s %-60s %s' % (n, 100.0 * n / tot, k, r[i_ex], r[i_src].strip()[:60], top))
