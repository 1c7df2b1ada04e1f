"""Build libdwg_sm100.so (in-tree) with nvcc for sm_100a.  No torch dependency: the library is a
plain C-ABI shared object (include/dwg.h); Python talks to it through ctypes (dwg/_lib.py).

    python -m dwg.build            # or dwg.build.build()
"""
import concurrent.futures as cf
import hashlib
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))      # dreamwaltz-g_b200/
CSRC = os.path.join(PKG, 'csrc')
OBJ = os.path.join(PKG, 'build')
SO = os.path.join(PKG, 'libdwg_sm100.so')

ARCH = ['-gencode', 'arch=compute_100a,code=sm_100a']
COMMON = ['-O3', '-std=c++17', '-lineinfo', '-Xcompiler', '-fPIC', '--expt-relaxed-constexpr', '-Xptxas', '-v']
# per-file extra flags (see the header comment of each file)
EXTRA = {
    'raster_pre.cu': ['-fmad=false'],       # bit-exact binning: no FMA contraction
}


def _nvcc():
    for c in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return 'nvcc'


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith('.cu'))


def _digest(path, flags):
    h = hashlib.sha1()
    h.update(' '.join(flags).encode())
    for dep in [path] + [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(('.cuh', '.h', '.inc'))] + \
            [os.path.join(os.path.dirname(PKG), 'include', 'dwg.h')]:
        with open(dep, 'rb') as fh:
            h.update(fh.read())
    return h.hexdigest()


def _compile(src, verbose):
    flags = ARCH + COMMON + EXTRA.get(src, [])
    obj = os.path.join(OBJ, src[:-3] + '.o')
    stamp = obj + '.sha1'
    dig = _digest(os.path.join(CSRC, src), flags)
    if os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == dig:
        return obj, ''
    cmd = [_nvcc()] + flags + ['-c', os.path.join(CSRC, src), '-o', obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f'nvcc failed for {src}:\n{r.stdout}\n{r.stderr}')
    with open(stamp, 'w') as fh:
        fh.write(dig)
    return obj, r.stderr if verbose else ''


def build(verbose=False, force=False):
    """Compile every csrc/*.cu for sm_100a and link libdwg_sm100.so.  Returns the .so path."""
    os.makedirs(OBJ, exist_ok=True)
    if force:
        for f in os.listdir(OBJ):
            os.remove(os.path.join(OBJ, f))
    srcs = sources()
    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        res = list(ex.map(lambda s: _compile(s, verbose), srcs))
    objs = [o for o, _ in res]
    if verbose:
        for (_, log), s in zip(res, srcs):
            if log:
                print(f'--- {s}\n{log}')
    newest = max(os.path.getmtime(o) for o in objs)
    if force or not os.path.exists(SO) or os.path.getmtime(SO) < newest:
        cmd = [_nvcc()] + ARCH + ['-shared', '-o', SO] + objs + ['-lcudart']
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f'link failed:\n{r.stdout}\n{r.stderr}')
    return SO


if __name__ == '__main__':
    print(build(verbose='-v' in sys.argv, force='-f' in sys.argv))
