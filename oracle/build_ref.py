#!/usr/bin/env python3
"""TEST INFRASTRUCTURE.  Build oracle/_ref/libmobiref.so from the reference's own decoder source.

The reference (Gericom/MobiclipDecoder) is C# and this image has no .NET/Mono, so the reference
cannot be *run* here.  What we can do is compile its decoder *source, where it lies under
/root/reference*, after a purely syntactic C# -> C++ transliteration (type spellings, `ref`
parameters, `new T[n]`, `fixed`, access modifiers).  No statement of decoder logic is rewritten:
every arithmetic expression, table literal and control-flow construct of
  LibMobiclip/Codec/Mobiclip/MobiclipDecoder.cs, MobiConst.cs and Utils/IOUtil.cs
is compiled as the reference wrote it.  Managed-array semantics (zero-init, bounds-checked ->
exception -> frame aborted by the catch-all at MobiclipDecoder.cs:325) come from oracle/ref_shim.h.

Outputs go ONLY to oracle/_ref/ (git-ignored, ships to the GPU box as a built artefact):
    oracle/_ref/gen_*.h          transliterated sources (never committed)
    oracle/_ref/libmobiref.so    C API in oracle/ref_capi.cpp

Used by tests/ and bench.py's cpu_baseline / --impl reference legs only.  What this is NOT: the
.NET runtime executing the original assembly; differences could hide in C# vs C++ integer
promotion (uint+int -> long in C#) and shift-count masking.  Both are irrelevant for in-contract
streams (all operands stay far below 2^31 and shift counts below 32); DESIGN.md says so too.
"""
import os, re, subprocess, sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get('MOBI_REFERENCE_DIR', '/root/reference')
OUT = os.path.join(HERE, '_ref')

PRIMS = r'(?:byte|sbyte|ushort|short|uint|int|ulong|long|float|bool)'
STATIC_CLASSES = r'(?:IOUtil|MobiConst|Array|Color|ImageLockMode|PixelFormat|MobiclipVersion)'


def cs_to_cpp(src: str) -> str:
    src = src.lstrip('﻿')
    src = re.sub(r'^\s*using\s+[\w.]+;\s*$', '', src, flags=re.M)
    # 3-D reverse table only used by the encoder; not on the decode path
    src = re.sub(r'public static readonly int\[, ,\].*?\n        \};', '', src, flags=re.S)
    src = re.sub(r'namespace\s+([\w.]+)', lambda m: 'namespace ' + m.group(1).replace('.', '_'), src)
    src = re.sub(r'public\s+(?:unsafe\s+)?(?:static\s+)?class\s+(\w+)', r'struct \1', src)
    src = re.sub(r'public\s+enum\s+(\w+)\s*\{([^}]*)\}', r'enum \1 {\2};', src)
    src = re.sub(r'fixed\s*\(([^()]*)\)\s*\{', r'{ \1;', src)
    # allocations
    src = re.sub(r'new\s+(%s)\[([^\]]+)\]\[\]' % PRIMS, r'Arr<Arr<\1>>::New(\2)', src)
    src = re.sub(r'new\s+(%s)\[([^\]]+)\]' % PRIMS, r'Arr<\1>::New(\2)', src)
    src = re.sub(r'new\s+(Bitmap|Rectangle|Exception|NotImplementedException)\s*\(', r'\1(', src)
    # array types
    src = re.sub(r'\b(%s)\[\]\[\]' % PRIMS, r'Arr<Arr<\1>>', src)
    src = re.sub(r'\b(%s)\[\]' % PRIMS, r'Arr<\1>', src)
    # ref parameters: declarations first (type + name), then call sites (name only)
    src = re.sub(r'\bref\s+((?:Arr<\w+>|\w+))\s+(\w+)', r'\1& \2', src)
    src = re.sub(r'\bref\s+(\w+)\s*([,)])', r'\1\2', src)
    # modifiers
    src = re.sub(r'\b(?:public|private|protected|internal|unsafe|readonly)\s+', '', src)
    src = re.sub(r'\bstatic\s+', 'static inline ', src)
    src = src.replace('this.', 'this->')
    src = re.sub(r'\b(%s)\.' % STATIC_CLASSES, r'\1::', src)
    src = re.sub(r'catch\s*\{', 'catch (...) {', src)
    src = re.sub(r'(?<![\w.])(\d+)f\b', r'\1.0f', src)
    # `struct X { ... }` needs a trailing ';' : the class brace is the second-to-last '}'
    last = src.rstrip().rfind('}')
    cls = src.rfind('}', 0, last)
    src = src[:cls] + '};' + src[cls + 1:]
    return src


def main():
    os.makedirs(OUT, exist_ok=True)
    files = {
        'gen_IOUtil.h': 'LibMobiclip/Utils/IOUtil.cs',
        'gen_MobiConst.h': 'LibMobiclip/Codec/Mobiclip/MobiConst.cs',
        'gen_MobiclipDecoder.h': 'LibMobiclip/Codec/Mobiclip/MobiclipDecoder.cs',
    }
    if not os.path.isdir(REF):
        print('build_ref: %s not present; keeping any prebuilt oracle/_ref/libmobiref.so' % REF)
        return 0
    for out, rel in files.items():
        cs = open(os.path.join(REF, rel), encoding='utf-8-sig').read()
        with open(os.path.join(OUT, out), 'w') as f:
            f.write('// transliterated at build time from %s -- NOT committed, do not edit\n' % rel)
            f.write(cs_to_cpp(cs))
    so = os.path.join(OUT, 'libmobiref.so')
    cmd = ['g++', '-std=c++17', '-O2', '-fPIC', '-shared', '-fwrapv', '-ffp-contract=off', '-fno-strict-aliasing',
           '-w', '-fmax-errors=30', '-I', HERE, '-I', OUT, os.path.join(HERE, 'ref_capi.cpp'), '-o', so]
    print(' '.join(cmd))
    subprocess.check_call(cmd)
    print('built', so)
    return 0


if __name__ == '__main__':
    sys.exit(main())
