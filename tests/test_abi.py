"""The C-ABI boundary: libmobicuda.so loads without a GPU and exports every function include/mobicuda.h declares
(and nothing is declared twice or bound with the wrong arity); same for libmobisynth.so.  No compute calls."""
import ctypes as C
import os
import re

import pytest

from mobiclipdecoder_b200 import _native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    src = open(os.path.join(ROOT, 'include', header)).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    src = re.sub(r'//[^\n]*', '', src)
    src = re.sub(r'typedef\s+(struct|enum)\s+\w+\s*\{.*?\}\s*\w+\s*;', '', src, flags=re.S)
    out = {}
    for m in re.finditer(r'\b(\w+)\s*\(([^;{}()]*)\)\s*;', src):
        name, args = m.group(1), m.group(2).strip()
        n = 0 if args in ('', 'void') else args.count(',') + 1
        out[name] = n
    return out


def test_mobicuda_exports_match_header():
    decl = _declared('mobicuda.h')
    assert len(decl) >= 35
    lib = C.CDLL(os.path.join(ROOT, 'mobiclipdecoder_b200', 'lib', 'libmobicuda.so'))
    for name, nargs in decl.items():
        assert hasattr(lib, name), 'libmobicuda.so does not export %s' % name
        assert name in _native.MOBICUDA_EXPORTS, 'no ctypes binding for %s' % name
        assert len(_native.MOBICUDA_EXPORTS[name][1]) == nargs, 'binding of %s has the wrong arity' % name
    assert set(_native.MOBICUDA_EXPORTS) == set(decl), 'bindings for undeclared functions: %s' % (set(_native.MOBICUDA_EXPORTS) - set(decl))
    assert _native.mobicuda().mobicuda_abi_version() == 1


def test_mobisynth_exports_match_header():
    decl = _declared('mobisynth.h')
    lib = C.CDLL(os.path.join(ROOT, 'mobiclipdecoder_b200', 'lib', 'libmobisynth.so'))
    for name, nargs in decl.items():
        assert hasattr(lib, name)
        assert len(_native.MOBISYNTH_EXPORTS[name][1]) == nargs
    assert set(_native.MOBISYNTH_EXPORTS) == set(decl)


def test_struct_layouts_match_header_comments():
    assert C.sizeof(_native.FrameHdr) == 384
    assert C.sizeof(_native.Mb) == 16
    assert C.sizeof(_native.Part) == 8
    assert C.sizeof(_native.Coef) == 4


def test_product_library_does_not_link_the_oracle():
    """The product path must not route through oracle/: no oracle symbol, no dependency on its libraries."""
    import subprocess
    so = os.path.join(ROOT, 'mobiclipdecoder_b200', 'lib', 'libmobicuda.so')
    syms = subprocess.run(['nm', '-D', so], capture_output=True, text=True).stdout
    assert 'mobi_oracle' not in syms and 'mobiref' not in syms
    needed = subprocess.run(['readelf', '-d', so], capture_output=True, text=True).stdout
    assert 'libmobioracle' not in needed and 'libmobiref' not in needed
    for f in os.listdir(os.path.join(ROOT, 'mobiclipdecoder_b200')):
        if f.endswith('.py'):
            text = open(os.path.join(ROOT, 'mobiclipdecoder_b200', f)).read()
            assert 'oracle_lib' not in text and 'libmobioracle' not in text.replace("'oracle', '_build'", '') or f == '_build.py'


def test_mobidemux_exports_match_header():
    decl = _declared('mobidemux.h')
    lib = C.CDLL(os.path.join(ROOT, 'mobiclipdecoder_b200', 'lib', 'libmobicuda.so'))
    for name, nargs in decl.items():
        assert hasattr(lib, name), 'libmobicuda.so does not export %s' % name
        assert len(_native.MOBIDEMUX_EXPORTS[name][1]) == nargs
    assert set(_native.MOBIDEMUX_EXPORTS) == set(decl)
    assert C.sizeof(_native.ModsHeader) == 0x30


def _build_c_driver(tmp):
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib = os.path.join(root, 'mobiclipdecoder_b200', 'lib')
    exe = os.path.join(str(tmp), 'c_driver')
    subprocess.check_call(['gcc', '-std=c99', '-Wall', '-Wextra', '-Werror', '-pedantic', '-I', os.path.join(root, 'include'),
                           os.path.join(root, 'tests', 'c_driver.c'), '-o', exe, '-L', lib, '-lmobicuda', '-Wl,-rpath,' + lib])
    return exe


def test_headers_are_plain_c_and_a_c_program_links(tmp_path):
    """include/*.h compile as C99 (-Wall -Wextra -Werror -pedantic) and a C caller making the C# shim's call sequence links
    against libmobicuda.so with nothing but the header: the boundary is a C ABI."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for hname in ('mobicuda.h', 'mobidemux.h', 'mobisynth.h'):
        src = os.path.join(str(tmp_path), 'inc_' + hname.replace('.h', '.c'))
        open(src, 'w').write('#include "%s"\nint main(void) { return 0; }\n' % hname)
        subprocess.check_call(['gcc', '-std=c99', '-Wall', '-Wextra', '-Werror', '-pedantic', '-fsyntax-only', '-I', os.path.join(root, 'include'), src])
    assert os.path.exists(_build_c_driver(tmp_path))


@pytest.mark.gpu
def test_c_driver_decodes_like_the_python_mirror(tmp_path):
    """The C caller and the ctypes mirror make the same calls; their planes, bitmaps, Offset and Quantizer agree frame by frame
    (and the mirror is what the parity tests compare with the oracle)."""
    import struct
    import subprocess
    import numpy as np
    from mobiclipdecoder_b200 import MobiclipDecoder
    from mobiclipdecoder_b200.workloads import CONFIGS, frames
    name = 'moflex_400x240'
    w, h, ver, _ = CONFIGS[name]
    fr = frames(name, 21, 8)
    path = os.path.join(str(tmp_path), 'frames.bin')
    with open(path, 'wb') as f:
        for data, key in fr:
            f.write(struct.pack('<I', len(data)) + bytes(data))
    out = subprocess.run([_build_c_driver(tmp_path), str(w), str(h), str(int(ver)), path], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l.split() for l in out.stdout.strip().splitlines()]
    assert len(lines) == len(fr)

    def fnv(arrs):
        hsh = 14695981039346656037
        for a in arrs:
            for b in np.ascontiguousarray(a).tobytes():
                hsh = ((hsh ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
        return '%016x' % hsh
    dec = MobiclipDecoder(w, h, ver)
    for (data, key), row in zip(fr[:3], lines):   # the pure-Python FNV is slow: three frames are enough to tie the two callers together
        dec.Data, dec.Offset = data, 0
        bmp = dec.DecodeFrame()
        assert bmp is not None and int(row[1]) == 0 and int(row[2]) == dec.Offset and int(row[3]) == dec.Quantizer
        assert row[4] == fnv([dec.Y[0], dec.UV[0]]) and row[5] == fnv([bmp])
    assert all(int(r[1]) == 0 for r in lines)
    dec.close()
