#!/usr/bin/env python3
"""Compulsory DRAM read traffic of the inter kernel on the bench workload, computed on the CPU from the parsed streams:
every leaf's luma (w+1) x (h+1) and chroma windows are mapped to 32-byte sectors of the picture they point into, sectors
are de-duplicated per (stream, reference picture) -- i.e. a perfect L2 within a step is assumed -- and summed.  This is
what the ring layout and the sector granularity make unavoidable; the measured dram__bytes_read of k_inter_chunk is
compared with it in DESIGN.md.

    python tools/traffic_floor.py [streams] [step]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np


def main():
    from mobiclipdecoder_b200 import MobiParser
    from mobiclipdecoder_b200.workloads import CONFIGS, make_stream
    name = 'moflex_400x240'
    n_streams = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    step = int(sys.argv[2]) if len(sys.argv) > 2 else 12
    w, h, ver, ov = CONFIGS[name]
    S = 256 if w <= 256 else 512 if w <= 512 else 1024
    mbw = w // 16
    gop = ov['gop']
    tot_sec = tot_alg = tot_mb = tot_parts = 0
    box_sec = 0
    for i in range(n_streams):
        s = make_stream(name, 1000 + i, gop_phase=(i * 37) % gop)
        par = MobiParser(w, h, ver)
        pf = None
        for k in range(step + 1):
            data, key = s.next_frame()
            rc, off, pf = par.parse(data, 0)
            assert rc == 0
        s.close()
        hdr = pf.hdr.contents
        if hdr.flags & 1:
            par.close()
            continue   # an I-picture this step
        secs = {}      # ref -> set of sector ids (luma plane and chroma plane of one picture share the id space)
        boxes = {}
        for m in range(hdr.n_mb):
            mb = pf.mbs[m]
            if mb.info & 3:
                continue
            tot_mb += 1
            mbx, mby = m % mbw, m // mbw
            npart = (mb.info >> 2) & 127
            for q in range(npart):
                p = pf.parts[mb.first_sub + q]
                tot_parts += 1
                x, y = (p.xy & 15) * 2, (p.xy >> 4) * 2
                pw, ph = 2 << (p.shape & 3), 2 << ((p.shape >> 2) & 3)
                ref = p.shape >> 4
                st = secs.setdefault(ref, set())
                # luma window
                x0, y0 = mbx * 16 + x + (p.mvx >> 1), mby * 16 + y + (p.mvy >> 1)
                nx, ny = pw + (p.mvx & 1), ph + (p.mvy & 1)
                for yy in range(y0, y0 + ny):
                    a = yy * S + x0
                    for sec in range(a // 32, (a + nx - 1) // 32 + 1):
                        st.add(sec)
                # chroma windows (U at column c, V at column S/2 + c of the chroma rows, which follow the S*h luma bytes)
                cx, cy = p.mvx >> 1, p.mvy >> 1
                cx0, cy0 = mbx * 8 + x // 2 + (cx >> 1), mby * 8 + y // 2 + (cy >> 1)
                cnx, cny = max(1, pw // 2) + (cx & 1), max(1, ph // 2) + (cy & 1)
                for yy in range(cy0, cy0 + cny):
                    for base in (0, S // 2):
                        a = S * h + yy * S + base + cx0
                        for sec in range(a // 32, (a + cnx - 1) // 32 + 1):
                            st.add(sec)
        tot_sec += sum(len(v) for v in secs.values())
        # what k_inter_chunk asks for: unsplit macroblocks and macroblocks split once fetch, per leaf, one 32 x 17 luma box and one
        # 32 x 2 x 9 chroma box starting at the window column rounded down to 16 (whole-macroblock window at the leaf's vector);
        # the others load exactly their windows
        for m in range(hdr.n_mb):
            mb = pf.mbs[m]
            if mb.info & 3:
                continue
            mbx, mby = m % mbw, m // mbw
            npart = (mb.info >> 2) & 127
            for q in range(npart):
                p = pf.parts[mb.first_sub + q]
                ref = p.shape >> 4
                st = boxes.setdefault(ref, set())
                if npart <= 2:
                    x0, y0 = mbx * 16 + (p.mvx >> 1), mby * 16 + (p.mvy >> 1)
                    cx0, cy0 = mbx * 8 + (p.mvx >> 2), mby * 8 + (p.mvy >> 2)
                    if x0 >= 0 and x0 + 17 <= S and cx0 >= 0 and cx0 + 9 <= S // 2:
                        for yy in range(y0, y0 + 17):
                            a = yy * S + (x0 & ~15)
                            st.add(a // 32); st.add((a + 31) // 32)
                        for yy in range(cy0, cy0 + 9):
                            for base in (0, S // 2):
                                a = S * h + yy * S + base + (cx0 & ~15)
                                st.add(a // 32); st.add((a + 31) // 32)
                        continue
                x, y = (p.xy & 15) * 2, (p.xy >> 4) * 2
                pw, ph = 2 << (p.shape & 3), 2 << ((p.shape >> 2) & 3)
                x0, y0 = mbx * 16 + x + (p.mvx >> 1), mby * 16 + y + (p.mvy >> 1)
                nx, ny = pw + (p.mvx & 1), ph + (p.mvy & 1)
                for yy in range(y0, y0 + ny):
                    a = yy * S + x0
                    for sec in range(a // 32, (a + nx - 1) // 32 + 1):
                        st.add(sec)
                cx, cy = p.mvx >> 1, p.mvy >> 1
                cx0, cy0 = mbx * 8 + x // 2 + (cx >> 1), mby * 8 + y // 2 + (cy >> 1)
                cnx, cny = max(1, pw // 2) + (cx & 1), max(1, ph // 2) + (cy & 1)
                for yy in range(cy0, cy0 + cny):
                    for base in (0, S // 2):
                        a = S * h + yy * S + base + cx0
                        for sec in range(a // 32, (a + cnx - 1) // 32 + 1):
                            st.add(sec)
        box_sec += sum(len(v) for v in boxes.values())
        tot_alg += 384 * sum(1 for m in range(hdr.n_mb) if not (pf.mbs[m].info & 3))
        par.close()
    print('%d P-pictures: %d inter MBs, %.2f leaves/MB' % (n_streams, tot_mb, tot_parts / max(1, tot_mb)))
    print('algorithmic reference read (384 B / MB): %.1f KB per picture' % (tot_alg / n_streams / 1e3))
    print('compulsory sector read (32 B sectors, de-duplicated per picture and reference): %.1f KB per picture = %.2fx algorithmic' % (tot_sec * 32 / n_streams / 1e3, tot_sec * 32 / max(1, tot_alg)))
    print('as fetched by k_inter_chunk (aligned boxes for one- and two-leaf macroblocks): %.1f KB per picture = %.2fx algorithmic = %.0f MB per step' % (box_sec * 32 / n_streams / 1e3, box_sec * 32 / max(1, tot_alg), box_sec * 32 / n_streams * 1013 / 1e6))
    print('=> per 1024-picture step: %.0f MB compulsory reference read (algorithmic %.0f MB)' % (tot_sec * 32 / n_streams * 1013 / 1e6, tot_alg / n_streams * 1013 / 1e6))


if __name__ == '__main__':
    main()
