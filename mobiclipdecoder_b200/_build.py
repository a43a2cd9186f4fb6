"""Builds the native libraries in-tree (they travel to the GPU box with the snapshot).

  mobiclipdecoder_b200/lib/libmobicuda.so   the product: host parser + sm_100a kernels + C ABI (include/mobicuda.h)
  mobiclipdecoder_b200/lib/libmobisynth.so  test/bench input generator (include/mobisynth.h), host only
  oracle/_build/libmobioracle.so            TEST INFRASTRUCTURE: C restatement of the reference decoder
  oracle/_ref/libmobiref.so                 TEST INFRASTRUCTURE: the reference's own source, transliterated and
                                            compiled (only where /root/reference exists; see oracle/build_ref.py)
"""
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, 'csrc')
LIB = os.path.join(PKG, 'lib')

NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17', '-fmad=false',
              '-Xcompiler', '-fPIC,-O3,-fno-strict-aliasing,-pthread', '-Xptxas', '-v']


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _run(cmd, verbose):
    if verbose:
        print(' '.join(cmd), flush=True)
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError('build failed: ' + ' '.join(cmd))
    if verbose:
        print(r.stdout)
    return r.stdout


def nvcc_path():
    p = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(p):
        raise RuntimeError('nvcc not found')
    return p


def build_mobicuda(force=False, verbose=False):
    os.makedirs(LIB, exist_ok=True)
    out = os.path.join(LIB, 'libmobicuda.so')
    srcs = [os.path.join(CSRC, f) for f in ('mobi_kernels.cu', 'mobi_runtime.cu', 'mobi_parse.cpp', 'mobi_demux.cpp')]
    deps = srcs + [os.path.join(CSRC, f) for f in ('mobi_kernels.h', 'mobi_inter_v3.cuh', 'mobi_inter_split.cuh', 'mobi_parse.h', 'mobi_tables.h')] + [os.path.join(ROOT, 'include', 'mobicuda.h'), os.path.join(ROOT, 'include', 'mobidemux.h')]
    if force or _newer(out, deps):
        log = _run([nvcc_path()] + NVCC_FLAGS + ['-shared', '-o', out] + srcs, verbose)
        with open(os.path.join(LIB, 'ptxas.log'), 'w') as f:
            f.write(log)
    return out


def build_mobisynth(force=False, verbose=False):
    os.makedirs(LIB, exist_ok=True)
    out = os.path.join(LIB, 'libmobisynth.so')
    src = os.path.join(CSRC, 'mobi_synth.cpp')
    deps = [src, os.path.join(CSRC, 'mobi_tables.h'), os.path.join(ROOT, 'include', 'mobisynth.h')]
    if force or _newer(out, deps):
        _run(['g++', '-std=c++17', '-O2', '-fPIC', '-shared', '-o', out, src], verbose)
    return out


def build_oracle(force=False, verbose=False):
    d = os.path.join(ROOT, 'oracle', '_build')
    os.makedirs(d, exist_ok=True)
    out = os.path.join(d, 'libmobioracle.so')
    src = os.path.join(ROOT, 'oracle', 'mobi_oracle.c')
    deps = [src, os.path.join(ROOT, 'oracle', 'mobi_oracle.h'), os.path.join(CSRC, 'mobi_tables.h')]
    if force or _newer(out, deps):
        _run(['gcc', '-std=c11', '-O2', '-fPIC', '-shared', '-fwrapv', '-ffp-contract=off', '-fno-strict-aliasing', '-o', out, src], verbose)
    return out


def build_ref(force=False, verbose=False):
    """oracle/_ref/libmobiref.so from the reference sources where they lie; a no-op where /root/reference is absent."""
    out = os.path.join(ROOT, 'oracle', '_ref', 'libmobiref.so')
    ref = os.environ.get('MOBI_REFERENCE_DIR', '/root/reference')
    if not os.path.isdir(ref):
        return out if os.path.exists(out) else None
    deps = [os.path.join(ROOT, 'oracle', f) for f in ('build_ref.py', 'ref_capi.cpp', 'ref_shim.h')]
    if force or _newer(out, deps):
        _run([sys.executable, os.path.join(ROOT, 'oracle', 'build_ref.py')], verbose)
    return out


def build_all(force=False, verbose=False):
    return {'mobicuda': build_mobicuda(force, verbose), 'mobisynth': build_mobisynth(force, verbose),
            'oracle': build_oracle(force, verbose), 'ref': build_ref(force, verbose)}


if __name__ == '__main__':
    print(build_all(force='--force' in sys.argv, verbose=True))
