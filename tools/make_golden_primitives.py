#!/usr/bin/env python3
"""Generates tests/golden/primitives_second_copies.json: SHA-256 digests of the outputs of the reference's SECOND
(encoder-side) copies of the reconstruction primitives -- FrameUtil.GetPBlock, MobiEncoder.IDCT64 / IDCT16,
MacroBlock.GetCompvals8x8 / 4x4, MacroBlock.PredictIntraPlane16x16 / 8x8 / 4x4, compiled from the reference's files by
oracle/build_ref.py -- on seeded random inputs.  tests/test_golden_primitives.py regenerates the same inputs and checks
the ORACLE's primitives against the digests, so the pin holds where neither /root/reference nor oracle/_ref exists.

    python tools/make_golden_primitives.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def main():
    from mobiclipdecoder_b200 import _build
    _build.build_all()
    from oracle_lib import have_ref
    if not have_ref():
        raise SystemExit('oracle/_ref/libmobiref.so is not built: these digests come from the compiled reference source only')
    import primitive_cases as pc
    out = {'made_by': 'tools/make_golden_primitives.py', 'source': 'reference second copies (oracle/build_ref.py: gen_SecondCopies.h)', 'digests': pc.digests(pc.run_second_copies)}
    path = os.path.join(ROOT, 'tests', 'golden', 'primitives_second_copies.json')
    with open(path, 'w') as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print('wrote', path, {k: v[:12] for k, v in out['digests'].items()})
    # the encoder's inverse partition-tree code tables (Analyzer.cs:472-526, Moflex3DS): [log2 w - 1][log2 h - 1][symbol] -> (value, bits)
    import re
    an = open('/root/reference/LibMobiclip/Codec/Mobiclip/Analyzer.cs', encoding='utf-8-sig').read()

    def table(name):
        m = re.search(r'%s = new int\[4, 4, 10\]\s*(\{.*?\n            \});' % name, an, flags=re.S)
        import ast
        txt = re.sub(r'(\d+)\s*>>\s*(\d+)', lambda k: str(int(k.group(1)) >> int(k.group(2))), m.group(1).rstrip(';'))   # the only operator in the table
        return ast.literal_eval(txt.replace('{', '[').replace('}', ']'))   # integer literals only; never eval() reference text
    tabs = {'source': 'LibMobiclip/Codec/Mobiclip/Analyzer.cs:472-526 (HuffEncodeValTable, HuffEncodeBitTable)', 'value': table('HuffEncodeValTable'), 'bits': table('HuffEncodeBitTable')}
    # the encoder's inverse coded-block-pattern tables (MobiEncoder.cs:149-161, 407-415): pattern -> varint value
    me = open('/root/reference/LibMobiclip/Codec/Mobiclip/Encoder/MobiEncoder.cs', encoding='utf-8-sig').read()
    for name in ('REV_byte_116160', 'REV_byte_1165C4', 'REV_byte_115FC4', 'REV_byte_1164F4'):
        m = re.search(r'%s\s*=\s*\{(.*?)\};' % name, me, flags=re.S)
        tabs[name] = [int(x) for x in re.findall(r'\d+', m.group(1))]
    tabs['source_cbp'] = 'LibMobiclip/Codec/Mobiclip/Encoder/MobiEncoder.cs:149-161, 407-415'
    # the encoder's own copies of the dequantisation tables (MobiEncoder.cs:862-917) behind its SetupQuantizationTables (:930-960)
    for name in ('byte_118DD4', 'byte_118F94', 'byte_119004', 'byte_11903A'):
        m = re.search(r'byte\[\] %s\s*=\s*\{(.*?)\};' % name, me, flags=re.S)
        tabs['enc_' + name] = [int(x, 16) for x in re.findall(r'0x([0-9A-Fa-f]+)', m.group(1))]
    tabs['source_quant'] = 'LibMobiclip/Codec/Mobiclip/Encoder/MobiEncoder.cs:862-960'
    path = os.path.join(ROOT, 'tests', 'golden', 'tables_partition_encoder.json')
    with open(path, 'w') as f:
        json.dump(tabs, f, sort_keys=True)
    print('wrote', path)


if __name__ == '__main__':
    main()
