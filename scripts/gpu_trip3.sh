#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python scripts/diag_e2e.py > gpurun_out/diag_e2e.log 2>&1; echo "diag exit $?"
cat gpurun_out/diag_e2e.log
timeout 400 python bench.py --steps 5 --warmup 3 --precision f16x2 > gpurun_out/bench_f16x2.json 2> gpurun_out/bench_f16x2.err; echo "bench exit $?"
cat gpurun_out/bench_f16x2.json; tail -n 12 gpurun_out/bench_f16x2.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1a.csv python bench.py --steps 1 --warmup 3 --precision f16x2 > gpurun_out/ncu_bench.log 2>&1; echo "ncu exit $?"
python - <<'PY'
import csv, collections
rows = list(csv.reader(open('gpurun_out/launches_r1a.csv', errors='ignore')))
hdr = None; agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    if 'Kernel Name' in r: hdr = r; continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        try: v = float(d['Metric Value'].replace(',', ''))
        except: continue
        unit = d.get('Metric Unit', 'ns')
        v = v / 1e3 if unit == 'ns' else (v if unit == 'us' else v * 1e3 if unit == 'ms' else v)
        k = d['Kernel Name'][:70]; agg[k][0] += 1; agg[k][1] += v
tot = sum(v[1] for v in agg.values())
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
    print(f'{t/1e3:9.2f} ms {100*t/tot:5.1f}%  n={n:5d}  {k}')
print('total', tot/1e3, 'ms')
PY
