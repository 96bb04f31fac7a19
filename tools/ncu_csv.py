"""Prints one line per profiled launch of an `ncu --csv --metrics ...` log (kernel, then metric = value)."""
import csv, sys
for f in sys.argv[1:]:
    rows = [r for r in csv.reader(open(f)) if len(r) > 10]
    hdr, by = rows[0], {}
    for r in rows[1:]:
        d = dict(zip(hdr, r))
        by.setdefault((int(d['ID']), d['Kernel Name'][:48]), {})[d['Metric Name']] = d['Metric Value']
    print(f)
    for (i, k), v in sorted(by.items()):
        print('  %2d %-48s' % (i, k), ' '.join('%s=%s' % (a.split('__')[1].split('.')[0][:14], b) for a, b in v.items()))
