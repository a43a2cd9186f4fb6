#!/usr/bin/env python3
"""Per-source-line dynamic instruction counts and stall samples of one kernel from an `ncu --set full --import-source on`
report, read here without a GPU.  ncu's source page is per SASS instruction; the line table comes from nvdisasm on the
library the report was taken with (same build!), matched by instruction order.

    python tools/ncu_source_lines.py gpurun_out/prof.ncu-rep k_inter_chunkILi9 [units] > profiles/rNN_kinter_lines.txt

`units` (default 1) divides the instruction counts, e.g. the number of macroblocks of the launch."""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, 'mobiclipdecoder_b200', 'lib', 'libmobicuda.so')
CSRC = os.path.join(ROOT, 'mobiclipdecoder_b200', 'csrc')


def line_table(mangled_part):
    with tempfile.TemporaryDirectory() as td:
        subprocess.run('cuobjdump -xelf all %s >/dev/null' % SO, shell=True, cwd=td, check=True)
        cubin = [f for f in os.listdir(td) if f.startswith('mobi_kernels') and f.endswith('.cubin')][0]
        txt = subprocess.run(['nvdisasm', '--print-line-info', cubin], cwd=td, capture_output=True, text=True).stdout.split('\n')
    start = [i for i, l in enumerate(txt) if '.section' in l and '.text' in l and mangled_part in l][0]
    lines, cur = [], None
    for l in txt[start + 1:]:
        if l.strip().startswith('.section'):
            break
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            f = os.path.basename(m.group(1))
            cur = (f, int(m.group(2))) if os.path.exists(os.path.join(CSRC, f)) else ('', -1)
            continue
        if re.match(r'\s*/\*[0-9a-f]{4}\*/', l):
            lines.append(cur)
    return lines


def sass_rows(rep, name_part, n_instr):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdrs = [i for i, r in enumerate(rows) if 'Instructions Executed' in r]
    names = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name']
    for k, hi in enumerate(hdrs):
        title = rows[max(i for i in names if i < hi)][1]
        end = min([i for i in names if i > hi] + [len(rows)])
        body = [r for r in rows[hi + 1:end] if len(r) == len(rows[hi])]
        name = re.split(r'EPK|IL[a-z]', name_part.lstrip('0123456789'))[0]   # mangled fragment -> plain kernel name
        if re.search(r'::%s[(<]' % re.escape(name), title) and len(body) == n_instr:
            return rows[hi], body, title
    raise SystemExit('no launch of %s with %d instructions in %s' % (name_part, n_instr, rep))


def main():
    rep, part = sys.argv[1], sys.argv[2]
    units = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
    lines = line_table(part)
    h, body, title = sass_rows(rep, part, len(lines))
    ci, sa = h.index('Instructions Executed'), h.index('# Samples')
    wf, we = h.index('L1 Wavefronts Shared'), h.index('L1 Wavefronts Shared Excessive')
    stall = {n: h.index(n) for n in h if n.startswith('stall_') and '(Not Issued)' not in n}
    per = collections.defaultdict(lambda: [0.0, 0.0, 0.0, 0.0])
    tot = collections.Counter()
    for ln, r in zip(lines, body):
        p = per[ln]
        p[0] += float(r[ci] or 0); p[1] += float(r[sa] or 0); p[2] += float(r[wf] or 0); p[3] += float(r[we] or 0)
        for n, c in stall.items():
            tot[n] += float(r[c] or 0)
    T, S = sum(p[0] for p in per.values()), sum(p[1] for p in per.values())
    src = {}
    print('# %s' % title)
    print('# warp instructions executed: %.0f (%.1f per unit of %.0f), stall samples: %.0f' % (T, T / units, units, S))
    print('# stall reasons: ' + ', '.join('%s %.1f%%' % (n[6:], 100 * v / S) for n, v in tot.most_common(8)))
    print('# shared-memory wavefronts: %.3g, of which excessive (bank conflicts): %.3g' % (sum(p[2] for p in per.values()), sum(p[3] for p in per.values())))
    print('# line  instr/unit  samples%  smem-wavefronts(excess)  source')
    for ln in sorted(per, key=lambda x: (x is None, x or ('', 0))):
        n, s, w, e = per[ln]
        if n / units < 0.25 and s / S < 0.003:
            continue
        if ln and ln[1] > 0:
            if ln[0] not in src:
                src[ln[0]] = open(os.path.join(CSRC, ln[0])).read().split('\n')
            text, tag = src[ln[0]][ln[1] - 1].strip()[:110], '%s:%d' % ('v3' if 'v3' in ln[0] else 'k', ln[1])
        else:
            text, tag = '(inlined CUDA header code: __ldg, __shfl_sync, __funnelshift, __popc ...)', '-1'
        print('%8s %10.1f %8.1f%% %12.0f(%.0f)  %s' % (tag, n / units, 100 * s / S, w, e, text))


if __name__ == '__main__':
    main()
