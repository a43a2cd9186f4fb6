"""The C-ABI boundary: libmobicuda.so loads without a GPU and exports every function include/mobicuda.h declares
(and nothing is declared twice or bound with the wrong arity); same for libmobisynth.so.  No compute calls."""
import ctypes as C
import os
import re

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
