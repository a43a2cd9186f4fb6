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
STATIC_CLASSES = r'(?:IOUtil|MobiConst|MobiConstRef|Array|Color|ImageLockMode|PixelFormat|MobiclipVersion|FrameUtil|Math|MoLive)'


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
    src = re.sub(r'new\s+(List<\w+>)\s*\(', r'\1(', src)
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



# ---- containers: classes with reference semantics, inheritance, Stream / Dictionary / delegates -------------------------
REF_CLASSES = ['MoLiveChunk', 'MoLiveStream', 'MoLiveStreamCodec', 'MoLiveStreamVideo', 'MoLiveStreamVideoWithLayout', 'MoLiveStreamAudio',
               'MoLiveStreamTimeline', 'MoLiveChunkFoo', 'MoLiveInBitStream', 'Endpoint', 'ModsHeader', 'KeyFrameInfo']
# identifiers that hold a reference to a class instance in these files: `x.` becomes `x->`
PTR_VARS = ['chunk', 'Chunk', 'p', 'bs', 'Reader', 'Stream', 'Header', 'mDestinationStream', 'Destination', 'dst', 'src']


def close_structs(src: str) -> str:
    """C++ wants `};` after every struct body (nested ones too)."""
    out, i = [], 0
    for m in re.finditer(r'\bstruct\s+\w+[^;{()]*\{', src):
        depth, j = 0, m.end() - 1
        while True:
            if src[j] == '{':
                depth += 1
            elif src[j] == '}':
                depth -= 1
                if depth == 0:
                    break
            j += 1
        out.append(j)
    for j in sorted(out, reverse=True):
        if not src[j + 1:].lstrip().startswith(';'):
            src = src[:j + 1] + ';' + src[j + 1:]
    return src


def cs_container_to_cpp(src: str) -> str:
    """The same kind of rewriting as cs_to_cpp -- spellings only, no statement of container logic is touched -- for the
    container classes: a C# class instance is a reference, so variables of those types become pointers (and `.` on them
    `->`), `out` parameters references, properties fields, virtual / override / abstract their C++ spellings, object
    initialisers a statement expression, System.IO.Stream the CsStream of ref_shim.h."""
    src = src.lstrip('﻿')
    for a, b in (('UInt64', 'ulong'), ('UInt32', 'uint'), ('UInt16', 'ushort'), ('Byte', 'byte')):
        src = re.sub(r'\b%s\b' % a, b, src)
    src = re.sub(r'public\s+(\w+(?:\[\])*)\s+(\w+)\s*\{\s*get;\s*(?:private\s+)?set;\s*\}', r'public \1 \2;', src)
    src = re.sub(r'public\s+delegate\s+void\s+(\w+)\s*\(([^)]*)\)\s*;', r'typedef std::function<void(\2)> \1;', src)
    src = re.sub(r'public\s+event\s+(\w+)\s+(\w+)\s*;', r'\1 \2;', src)
    src = re.sub(r'public\s+enum\s+(\w+)(\s*:\s*\w+)?\s*\{([^}]*)\}', r'enum \1\2 {\3};', src)
    # class headers; `base` is the one base class named there
    m = re.search(r'class\s+\w+\s*:\s*(\w+)', src)
    if m:
        src = re.sub(r'\bbase\.', m.group(1) + '::', src)
        src = re.sub(r':\s*base\s*\(', ': ' + m.group(1) + '(', src)
    src = re.sub(r'(?:public|private)\s+(?:abstract\s+)?class\s+(\w+)', r'struct \1', src)
    src = re.sub(r'\babstract\s+([\w\[\]<>]+)\s+(\w+)\s*\(([^)]*)\)\s*;', r'virtual \1 \2(\3) = 0;', src)
    src = re.sub(r'\b(?:sealed\s+)?override\s+([\w\[\]<>]+)\s+(\w+)\s*\(([^)]*)\)', r'virtual \1 \2(\3) override', src)
    # forward declarations of nested classes at the top of the enclosing one (members may name them before their definition)
    top = re.search(r'struct\s+(\w+)[^{;]*\{', src)
    if top:
        nested = [n for n in re.findall(r'struct\s+(\w+)[^{;]*\{', src[top.end():])]
        if nested:
            src = src[:top.end()] + ' ' + ' '.join('struct %s;' % n for n in nested) + src[top.end():]
    # object initialisers: new X() { A = 1, B = 2 }
    def obj_init(mm):
        sets = ' '.join('o_->%s;' % part.strip() for part in mm.group(2).split(',') if part.strip())
        return '({ %s* o_ = new %s(); %s o_; })' % (mm.group(1), mm.group(1), sets)
    src = re.sub(r'new\s+(\w+)\s*\(\)\s*\{([^{}]*)\}', obj_init, src)
    src = re.sub(r'new\s+(Dictionary<[^>]*>)\s*\(', r'\1(', src)
    src = re.sub(r'new\s+(ArgumentException)\s*\(', r'\1(', src)
    # reference types -> pointers
    cls = '|'.join(REF_CLASSES)
    src = re.sub(r'new\s+(%s)\[([^\]]+)\]' % cls, r'Arr<\1*>::New(\2)', src)
    src = re.sub(r'\b(%s)\[\]' % cls, r'Arr<\1*>', src)
    src = re.sub(r'(?<!struct )(?<!new )\b(%s)[ \t]+(\w+)[ \t]*(?=[=;,)])' % cls, r'\1* \2', src)   # (same line only: comments name classes too)
    src = re.sub(r'\bStream[ \t]+(\w+)[ \t]*(?=[=;,)])', r'CsStream* \1', src)
    src = re.sub(r'<int,\s*(%s)>' % cls, r'<int, \1*>', src)
    src = re.sub(r'\(\((%s)\)(\w+)\)\.' % cls, r'((\1*)\2)->', src)
    src = re.sub(r'foreach\s*\(\s*(%s)\s+(\w+)\s+in\s+([\w.]+)\s*\)' % cls, r'for (\1* \2 : \3())', src)
    # C++ does not let a goto jump over an initialised declaration in the same scope; split it (no logic changes)
    src = re.sub(r'\b((?:%s)\*) (\w+)\s*=\s*null;' % cls, r'\1 \2; \2 = null;', src)
    src = re.sub(r'\b(%s)\.' % '|'.join(PTR_VARS), r'\1->', src)
    src = re.sub(r'\b(KeyFrames|Streams)\[([^\]]+)\]\.', r'\1[\2]->', src)
    # out parameters: like ref
    src = re.sub(r'\bout\s+(\w+)\s+(\w+)', r'\1& \2', src)
    src = re.sub(r'\bout\s+(\w+)\s*([,)])', r'\1\2', src)
    src = re.sub(r'\bEncoding\.', 'Encoding::', src)
    src = re.sub(r'\bString\b', 'CsString', src)
    src = close_structs(cs_to_cpp(src))
    # C# zero-initialises fields (and demands definite assignment of locals): declarations without an initialiser get `{}`
    src = re.sub(r'^([ \t]+)((?!return\b|goto\b|throw\b|typedef\b|using\b|struct\b|else\b|delete\b|case\b|break\b|continue\b)[\w:]+(?:<[^;()]*>)?\*?)[ \t]+(\w+);[ \t]*(//[^\n]*)?$',
                 r'\1\2 \3{};', src, flags=re.M)
    return src


def extract_methods(cs: str, names) -> str:
    """The full text (signature to matching brace) of the named methods of a C# file, in file order.  Used for the
    reference's SECOND copies of the reconstruction primitives, which live as static helpers inside encoder classes whose
    remaining members (List / LINQ / Bitmap based mode search) the syntactic transliteration cannot carry."""
    out = []
    for m in re.finditer(r'^[ \t]*(?:public|private)\s+(?:static\s+)?(?:unsafe\s+)?(?:static\s+)?[\w\[\]]+\s+(\w+)\s*\([^)]*\)\s*\{', cs, flags=re.M):
        if m.group(1) not in names:
            continue
        depth, i = 0, m.end() - 1
        while True:
            c = cs[i]
            if c == '{':
                depth += 1
            elif c == '}':
                depth -= 1
                if depth == 0:
                    break
            i += 1
        out.append(cs[m.start():i + 1])
    return '\n\n'.join(out)


def fix_fixed(src: str) -> str:
    """`fixed (byte* a = x, b = y) {` pins two arrays and declares two pointers; in C++ it becomes a block that opens with
    the declarations, one per pointer (`byte* a = x, b = y;` would make b a byte).  Initialisers may contain one level of
    parentheses, which cs_to_cpp's own `fixed` rule does not accept."""
    def repl(m):
        ty, decls = m.group(1), m.group(2)
        parts, depth, cur = [], 0, ''
        for ch in decls:
            if ch in '([':
                depth += 1
            elif ch in ')]':
                depth -= 1
            if ch == ',' and depth == 0:
                parts.append(cur.strip()); cur = ''
            else:
                cur += ch
        parts.append(cur.strip())
        return '{ ' + ' '.join('%s* %s;' % (ty, d) for d in parts)
    return re.sub(r'fixed\s*\(\s*(\w+)\s*\*\s*((?:[^()]|\([^()]*\))*)\)\s*\{', repl, src)


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
    # The reference's second copies of the primitives: FrameUtil (whole file: block get/set and
    # GetPBlock = CopyBlock's twin), and the static transform / predictor helpers of the encoder classes.
    second = {
        'LibMobiclip/Codec/Mobiclip/Encoder/MobiEncoder.cs': ('EncTransforms', ['DCT64', 'IDCT64', 'DCT16', 'IDCT16']),
        'LibMobiclip/Codec/Mobiclip/Encoder/MacroBlock.cs': ('EncPredictors', ['GetCompvals8x8', 'GetCompvals4x4', 'PredictIntraPlane16x16',
                                                                              'PredictIntraPlane8x8', 'PredictIntraPlane4x4']),
    }
    with open(os.path.join(OUT, 'gen_SecondCopies.h'), 'w') as f:
        f.write('// transliterated at build time from the reference -- NOT committed, do not edit\n')
        cs = open(os.path.join(REF, 'LibMobiclip/Utils/FrameUtil.cs'), encoding='utf-8-sig').read()
        f.write(cs_to_cpp(fix_fixed(cs)))
        for rel, (cls, names) in second.items():
            cs = open(os.path.join(REF, rel), encoding='utf-8-sig').read()
            body = extract_methods(cs, names)
            f.write('\n' + cs_to_cpp(fix_fixed('namespace LibMobiclip.Codec.Mobiclip.Encoder\n{\n    public class %s\n    {\n%s\n    }\n}\n' % (cls, body))))
    # The reference's own bit writer and coefficient entropy CODER (BitWriter.cs whole; MobiEncoder.EncodeDCT, which uses no
    # instance state; the [32, 64, 2] reverse code table of MobiConst.cs that cs_to_cpp drops from the decoder build): an
    # independent writer for the streams the parsers are tested on (tests/test_reference_entropy_writer.py).
    with open(os.path.join(OUT, 'gen_EntropyWriter.h'), 'w') as f:
        f.write('// transliterated at build time from the reference -- NOT committed, do not edit\n')
        mc = open(os.path.join(REF, 'LibMobiclip/Codec/Mobiclip/MobiConst.cs'), encoding='utf-8-sig').read()
        m = re.search(r'public static readonly int\[, ,\] VxTable0_A_Ref\s*=\s*(\{.*?\n        \});', mc, flags=re.S)
        import ast
        table = ast.literal_eval(m.group(1).rstrip(';').replace('{', '[').replace('}', ']'))   # integer literals only; never eval() reference text
        d0, d1, d2 = len(table), len(table[0]), len(table[0][0])
        assert all(len(r) == d1 and all(len(c) == d2 for c in r) for r in table)
        f.write('struct MobiConstRef {\n    static int At(long long a, long long b, long long c) {\n')
        f.write('        static const int T[%d][%d][%d] = %s;\n' % (d0, d1, d2, m.group(1).rstrip(';')))
        f.write('        if (a < 0 || a >= %d || b < 0 || b >= %d || c < 0 || c >= %d) throw IndexOutOfRangeException();\n        return T[a][b][c];\n    }\n};\n' % (d0, d1, d2))
        f.write(cs_to_cpp(open(os.path.join(REF, 'LibMobiclip/Codec/Mobiclip/BitWriter.cs'), encoding='utf-8-sig').read()))
        enc = extract_methods(open(os.path.join(REF, 'LibMobiclip/Codec/Mobiclip/Encoder/MobiEncoder.cs'), encoding='utf-8-sig').read(), ['EncodeDCT'])
        enc = enc.replace('private void EncodeDCT', 'public static void EncodeDCT').replace('BitWriter b)', 'BitWriterRef b)')
        enc = re.sub(r'MobiConst\.VxTable0_A_Ref\[(.*?)\]', r'MobiConstRef.At(\1)', enc)
        # C++ does not let `goto end` jump over an initialised declaration in the same scope; split it (no logic changes)
        enc = enc.replace('int newval = val -', 'int newval; newval = val -')
        f.write('\ntypedef LibMobiclip_Codec_Mobiclip::BitWriter& BitWriterRef;\n')
        f.write(cs_to_cpp('namespace LibMobiclip.Codec.Mobiclip.Encoder\n{\n    public class EncEntropy\n    {\n%s\n    }\n}\n' % enc))
    # Containers: ModsDemuxer (whole file), the MoLive demuxer with its chunk classes and bit reader, and the reference's own
    # Moflex MUXER base class (an independent writer for the files the demuxers are tested on).
    cont = ['Moflex/MoLive.cs', 'Moflex/MoLiveInBitStream.cs', 'Moflex/MoLiveChunk.cs', 'Moflex/MoLiveStream.cs', 'Moflex/MoLiveStreamCodec.cs',
            'Moflex/MoLiveStreamVideo.cs', 'Moflex/MoLiveStreamVideoWithLayout.cs', 'Moflex/MoLiveStreamAudio.cs', 'Moflex/MoLiveStreamTimeline.cs',
            'Moflex/MoLiveChunkFoo.cs', 'Moflex/MoLiveDemux.cs', 'Moflex/MoflexMuxer.cs', 'Mods/ModsDemuxer.cs']
    with open(os.path.join(OUT, 'gen_Containers.h'), 'w') as f:
        f.write('// transliterated at build time from the reference -- NOT committed, do not edit\n')
        for rel in cont:
            cs = open(os.path.join(REF, 'LibMobiclip/Containers', rel), encoding='utf-8-sig').read()
            f.write('\n// ---- %s ----\n' % rel)
            f.write(cs_container_to_cpp(cs))
    so = os.path.join(OUT, 'libmobiref.so')
    cmd = ['g++', '-std=c++17', '-O2', '-fPIC', '-shared', '-fwrapv', '-ffp-contract=off', '-fno-strict-aliasing',
           '-w', '-fpermissive', '-fmax-errors=30', '-I', HERE, '-I', OUT, os.path.join(HERE, 'ref_capi.cpp'), '-o', so]
    print(' '.join(cmd))
    subprocess.check_call(cmd)
    print('built', so)
    return 0


if __name__ == '__main__':
    sys.exit(main())
