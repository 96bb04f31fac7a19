"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list of tools/prof_step.py: the launches of the
LAST step (the list is cut at the last `image_pool_apply` group boundary = launches_per_step from the end), grouped by
kernel, with their share of the summed device time -> JSON for profiles/.
usage: python tools/launch_summary.py launches.csv launches_in_last_step out.json "command line\""""
import csv, json, re, sys
path, per_step, out, cmd = sys.argv[1], int(sys.argv[2]), sys.argv[3], sys.argv[4]
rows = [r for r in csv.reader(open(path)) if len(r) > 10]
hdr = rows[0]
ik, iv, iu = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
launches = []
for r in rows[1:]:
    v = float(r[iv].replace(',', ''))
    if r[iu] in ('ns', 'nsecond'):
        v /= 1e3
    elif r[iu] in ('ms', 'msecond'):
        v *= 1e3
    launches.append((r[ik], v))
last = launches[-per_step:] if per_step > 0 else launches
by = {}
for k, v in last:
    k = re.sub(r'\(.*$', '', k).replace('void ', '')
    e = by.setdefault(k, [0, 0.0])
    e[0] += 1
    e[1] += v
tot = sum(e[1] for e in by.values())
res = {"command": cmd, "launches_in_step": len(last), "sum_of_kernel_us": round(tot, 1),
       "kernels": [{"kernel": k, "launches": e[0], "us": round(e[1], 1), "share": round(e[1] / tot, 4)}
                   for k, e in sorted(by.items(), key=lambda x: -x[1][1])]}
json.dump(res, open(out, 'w'), indent=1)
for k in res["kernels"][:30]:
    print('%6.1f us %5.1f %% %4d  %s' % (k["us"], 100 * k["share"], k["launches"], k["kernel"][:90]))
print('total', round(tot, 1), 'us in', len(last), 'launches')
