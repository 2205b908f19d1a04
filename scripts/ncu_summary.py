"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and, with
--seq, the launch sequence of raft:: kernels (grid, duration)."""
import collections, csv, sys
path = sys.argv[1]
rows = list(csv.reader(open(path, errors='ignore')))
hdr = None; agg = collections.defaultdict(lambda: [0, 0.0]); seq = []
for r in rows:
    if 'Kernel Name' in r: hdr = r; continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        if d.get('Metric Name') != 'gpu__time_duration.sum': continue
        try: v = float(d['Metric Value'].replace(',', ''))
        except ValueError: continue
        unit = d.get('Metric Unit', 'ns')
        us = v / 1e3 if unit in ('ns', 'nsecond') else v if unit in ('us', 'usecond') else v * 1e3
        k = d['Kernel Name'][:64]; agg[k][0] += 1; agg[k][1] += us
        seq.append((k, d.get('Grid Size', ''), us))
tot = sum(v[1] for v in agg.values())
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
    print(f'{t/1e3:9.3f} ms {100*t/tot:5.1f}%  n={n:5d}  avg {t/n:8.1f} us  {k}')
print(f'total {tot/1e3:.3f} ms over {len(seq)} launches')
if '--seq' in sys.argv:
    start = int(sys.argv[sys.argv.index('--seq') + 1]); count = int(sys.argv[sys.argv.index('--seq') + 2])
    for k, g, us in seq[start:start + count]:
        print(f'{us:9.1f} us  grid {g:>16}  {k}')
