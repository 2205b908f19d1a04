"""DRAM traffic per launch of the profiled kernels, from `ncu --set full` captures -> profiles/r02_traffic.json
(read by bench.py for the `traffic` fields of its roofline objects).

usage: python scripts/ncu_traffic.py gpurun_out/r02_prof_mega.ncu-rep gpurun_out/r02_prof_lookup.ncu-rep ... > profiles/r02_traffic.json
"""
import csv, io, json, subprocess, sys

UNIT = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}


def main():
    out = {}
    for rep in sys.argv[1:]:
        raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units, data = rows[0], rows[1], rows[2:]
        col = {n: i for i, n in enumerate(hdr)}
        for r in data:
            name = r[col['Kernel Name']].replace('void ', '').replace('raft::', '').split('<')[0].split('(')[0]
            rd = float(r[col['dram__bytes_read.sum']]) * UNIT[units[col['dram__bytes_read.sum']]]
            wr = float(r[col['dram__bytes_write.sum']]) * UNIT[units[col['dram__bytes_write.sum']]]
            out[name] = {'bytes_per_launch': rd + wr, 'dram_read': rd, 'dram_write': wr,
                         'duration_us_under_ncu': float(r[col['gpu__time_duration.sum']]),
                         'tensor_pipe_pct': float(r[col['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']]),
                         'source': f'{rep} (dram__bytes_read.sum + dram__bytes_write.sum, ncu --set full, one launch; ncu flushes '
                                   'the caches before each replay, so operands that are L2-resident in the real step count as DRAM reads)'}
    json.dump(out, sys.stdout, indent=1)


if __name__ == '__main__':
    main()
