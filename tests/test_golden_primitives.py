"""The oracle's primitives against digests frozen from the reference's SECOND (encoder-side) copies of the same
primitives (tests/golden/primitives_second_copies.json, made by tools/make_golden_primitives.py where /root/reference
exists): motion compensation at all 16 sizes x 4 phases, 600 coefficient blocks through the inverse transforms, every
directional / DC predictor on luma, U and V, the three plane predictors.  Needs neither the reference nor oracle/_ref."""
import json
import os

import primitive_cases as pc

PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'primitives_second_copies.json')


def test_oracle_primitives_match_the_references_second_copies():
    want = json.load(open(PATH))['digests']
    got = pc.digests(pc.run_oracle)
    assert got == want


def test_partition_code_tables_match_the_encoders_inverse_tables():
    """The decoder reads partition trees through 16 peek tables per version (MobiclipDecoder.cs:458-1746, extracted into
    mobi_tables.h MOBI_PART_CODE); the reference's encoder writes them from inverse (value, bits) tables of its own
    (Analyzer.cs:472-526, Moflex3DS only; frozen in tests/golden/tables_partition_encoder.json).  Every code the encoder can
    write must read back as its symbol with the same length, for every completion of the peeked bits, and the encoder must
    know exactly the symbols the decoder accepts."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    enc = json.load(open(os.path.join(os.path.dirname(PATH), 'tables_partition_encoder.json')))
    txt = open(os.path.join(root, 'mobiclipdecoder_b200', 'csrc', 'mobi_tables.h')).read()
    body = txt[txt.index('MOBI_PART_CODE[2][4][4]'):]
    entries = re.findall(r'\{(\d+), \{([\d,]+)\}, \{([\d,]+)\}\}', body)
    assert len(entries) >= 16
    checked = 0
    for lw in range(4):
        for lh in range(4):
            peek, lens, syms = entries[lw * 4 + lh]          # version 0 = Moflex3DS comes first
            peek, lens, syms = int(peek), [int(x) for x in lens.split(',')], [int(x) for x in syms.split(',')]
            for s in range(10):
                bits, val = enc['bits'][lw][lh][s], enc['value'][lw][lh][s]
                if bits < 0:
                    assert lens[s] == 0, 'shape %d,%d: the decoder accepts symbol %d, the encoder cannot write it' % (lw, lh, s)
                    continue
                assert lens[s] == bits and bits <= peek
                for fill in range(1 << (peek - bits)):
                    assert syms[(val << (peek - bits)) | fill] == s
                checked += 1
    assert checked > 100


def test_coded_block_pattern_tables_match_the_encoders_inverse_tables():
    """byte_116160 / byte_1165C4 (inter, MD:1809, 2904) and byte_115FC4 / byte_1164F4 (intra, MD:1748, 2863) map a varint to a
    coded-block pattern; the encoder holds the inverse maps (MobiEncoder.cs:149-161, 407-415).  decoder[encoder[p]] == p for
    every pattern the encoder can write."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    enc = json.load(open(os.path.join(os.path.dirname(PATH), 'tables_partition_encoder.json')))
    txt = open(os.path.join(root, 'mobiclipdecoder_b200', 'csrc', 'mobi_tables.h')).read()

    def dec(name):
        body = re.search(r'%s\[\d+\] = \{(.*?)\};' % name, txt, flags=re.S).group(1)
        return [int(x, 0) for x in re.findall(r'0x[0-9A-Fa-f]+|\d+', body)]
    for dname, ename, n in (('MOBI_CBP6_INTER', 'REV_byte_116160', 64), ('MOBI_CBP4_INTER', 'REV_byte_1165C4', 16),
                            ('MOBI_CBP6_INTRA', 'REV_byte_115FC4', 64), ('MOBI_CBP4_INTRA', 'REV_byte_1164F4', 16)):
        d, e = dec(dname), enc[ename]
        assert len(e) == n
        for pattern in range(n):
            assert d[e[pattern]] == pattern, (dname, pattern)


def test_dequantisation_tables_match_the_encoders_setup():
    """MobiEncoder.SetupQuantizationTables (ME:930-960) builds its scale tables from its own copies of the constant tables:
    scale4[pos] = (byte_118F94[(q % 6) * 16 + pos] << (q / 6 + 8)) >> 8, scale8[pos] = (byte_118DD4[(q % 6) * 64 + pos] << (q / 6 + 6)) >> 8
    (q / 6 and q % 6 through byte_119004 / byte_11903A).  The product's host parser must hand the kernels exactly these
    scales for every legal quantiser (hdr.qtab = scale << 8 | position, MD:3897-3912)."""
    from mobiclipdecoder_b200 import MobiParser
    from mobiclipdecoder_b200.workloads import frames
    enc = json.load(open(os.path.join(os.path.dirname(PATH), 'tables_partition_encoder.json')))
    t8, t4, div6, mod6 = enc['enc_byte_118DD4'], enc['enc_byte_118F94'], enc['enc_byte_119004'], enc['enc_byte_11903A']
    assert len(t8) == 384 and len(t4) == 96 and len(div6) >= 53 and len(mod6) >= 53
    for q in range(12, 47):   # the stream generator writes quantisers 12..46
        (data, key), = frames('moflex_400x240', 3, 1, quant=q)
        par = MobiParser(400, 240, 2)
        rc, off, pf = par.parse(data, 0)
        assert rc == 0 and pf.hdr.contents.quantizer == q
        qtab = list(pf.hdr.contents.qtab)
        for pos in range(64):
            assert qtab[pos] >> 8 == (t8[mod6[q] * 64 + pos] << (div6[q] + 6)) >> 8, (q, pos)
        for pos in range(16):
            assert qtab[64 + pos] >> 8 == (t4[mod6[q] * 16 + pos] << (div6[q] + 8)) >> 8, (q, pos)
        par.close()
