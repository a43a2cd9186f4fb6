#!/usr/bin/env python3
"""profiles/ncu_traffic.json from a committed ncu summary: per kernel, dram__bytes_read.sum + dram__bytes_write.sum per launch
(mean over the captured launches), which bench.py prints as roofline.traffic.

    python tools/ncu_traffic.py moflex_400x240 profiles/r02e_prof_summary.csv [more workload=csv pairs ...]

The summary CSV is what tools/ncu_summary.py wrote from the `ncu --set full --clock-control none` capture of
`python bench.py --profile --steps 4 --warmup 2` (tools/gpu_ncu.sh); its second row holds the units."""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCALE = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
KERNELS = ['k_inter_chunk', 'k_inter_v3', 'k_mc', 'k_res', 'k_intra_key', 'k_intra', 'k_bgra', 'k_pack_i420']


def kernel_of(name):
    for k in KERNELS:
        if k + '<' in name or k + '(' in name or name.strip().endswith(k) or ('::' + k) in name:
            return k
    return None


def read(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    ir, iw, it = hdr.index('dram__bytes_read.sum'), hdr.index('dram__bytes_write.sum'), hdr.index('gpu__time_duration.sum')
    acc = {}
    for r in rows[2:]:
        k = kernel_of(r[0])
        if k is None:
            continue
        rd = float(r[ir].replace(',', '')) * SCALE[units[ir]]
        wr = float(r[iw].replace(',', '')) * SCALE[units[iw]]
        a = acc.setdefault(k, {'read': 0.0, 'write': 0.0, 'n': 0, 'time': 0.0})
        a['read'] += rd; a['write'] += wr; a['n'] += 1; a['time'] += float(r[it].replace(',', ''))
    return {k: {'bytes_per_launch': (a['read'] + a['write']) / a['n'], 'read': a['read'] / a['n'], 'write': a['write'] / a['n'], 'launches': a['n'],
                'ncu_time_per_launch_%s' % units[it]: a['time'] / a['n']} for k, a in acc.items()}


def main():
    out_path = os.path.join(ROOT, 'profiles', 'ncu_traffic.json')
    out = json.load(open(out_path)) if os.path.exists(out_path) else {}
    args = sys.argv[1:]
    for i in range(0, len(args), 2):
        wl, path = args[i], args[i + 1]
        detail = read(path)
        out[wl] = {k: v['bytes_per_launch'] for k, v in detail.items()}
        out.setdefault('_detail', {})[wl] = {'source': os.path.relpath(path, ROOT), 'kernels': detail}
    json.dump(out, open(out_path, 'w'), indent=1, sort_keys=True)
    print(json.dumps({k: v for k, v in out.items() if not k.startswith('_')}, indent=1))


if __name__ == '__main__':
    main()
