"""Build recipe for oracle/_ref/ (TEST INFRASTRUCTURE ONLY -- never imported by the product).

Compiles the reference's OWN native sources, from where they lie under /root/reference, into
oracle/_ref/ (git-ignored, shipped to the GPU box with the snapshot).  Nothing is copied into the
repository; /root/reference is never written.

  _gridencoder_ref*.so   <- core/nerf/gridencoder/src/{gridencoder.cu,bindings.cpp}  (the reference's
                            only native code on the hot path), UNMODIFIED, built with the flags of
                            core/nerf/gridencoder/backend.py:8-12 plus -gencode sm_100a.
  libref_raster_simt.so  <- oracle/ref_gpu_raster.cu (OUR plain-SIMT restatement of the published 3DGS
                            rasteriser, the labelled stand-in for the un-installable third-party
                            diff_gaussian_rasterization in the reference-equivalent GPU arm of bench.py).

It needs torch headers (pybind11 module taking at::Tensor) and a GPU to RUN, so it is used only
by `tests/golden/make_grid_golden.py` (golden vectors for R6, generated on a B200 through gpurun)
and by `bench.py --impl reference-gpu` (the reference-equivalent GPU arm).

    python -m oracle.build_ref            # no-op (returns None) when /root/reference is absent
"""
import glob
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get('DWG_REFERENCE_ROOT', '/root/reference')
OUT = os.path.join(_HERE, '_ref')
NAME = '_gridencoder_ref'


def built_so():
    hits = sorted(glob.glob(os.path.join(OUT, NAME + '*.so')))
    return hits[0] if hits else None


def build(force=False, verbose=False):
    src_dir = os.path.join(REF_ROOT, 'core', 'nerf', 'gridencoder', 'src')
    srcs = [os.path.join(src_dir, f) for f in ('gridencoder.cu', 'bindings.cpp')]
    if not all(os.path.exists(s) for s in srcs):
        return built_so()                       # GPU box: only the prebuilt file exists
    so = built_so()
    if so and not force and os.path.getmtime(so) >= max(os.path.getmtime(s) for s in srcs):      # the reference sources are read-only
        return so
    os.makedirs(os.path.join(OUT, 'build'), exist_ok=True)
    os.environ.setdefault('TORCH_CUDA_ARCH_LIST', '10.0a')
    os.environ.setdefault('MAX_JOBS', '4')
    from torch.utils.cpp_extension import load
    nvcc_flags = ['-O3', '-std=c++17', '-U__CUDA_NO_HALF_OPERATORS__', '-U__CUDA_NO_HALF_CONVERSIONS__',
                  '-U__CUDA_NO_HALF2_OPERATORS__', '-gencode', 'arch=compute_100a,code=sm_100a']
    load(name=NAME, sources=srcs, extra_cflags=['-O3', '-std=c++17'], extra_cuda_cflags=nvcc_flags,
         extra_include_paths=[src_dir], build_directory=os.path.join(OUT, 'build'), is_python_module=False, verbose=verbose)
    import shutil
    built = glob.glob(os.path.join(OUT, 'build', NAME + '*.so'))
    assert built, 'reference gridencoder did not produce a .so'
    dst = os.path.join(OUT, os.path.basename(built[0]))
    shutil.copy2(built[0], dst)
    return dst


RASTER_SO = os.path.join(OUT, 'libref_raster_simt.so')


def build_raster(force=False):
    """nvcc oracle/ref_gpu_raster.cu -> oracle/_ref/libref_raster_simt.so (works without /root/reference)."""
    import subprocess
    src = os.path.join(_HERE, 'ref_gpu_raster.cu')
    if not force and os.path.exists(RASTER_SO) and os.path.getmtime(RASTER_SO) >= os.path.getmtime(src):
        return RASTER_SO
    os.makedirs(OUT, exist_ok=True)
    nvcc = '/usr/local/cuda/bin/nvcc' if os.path.exists('/usr/local/cuda/bin/nvcc') else 'nvcc'
    subprocess.check_call([nvcc, '-O3', '-std=c++17', '-gencode', 'arch=compute_100a,code=sm_100a', '-shared', '-Xcompiler', '-fPIC',
                           '-o', RASTER_SO, src, '-lcudart'])
    return RASTER_SO


def load_module():
    """Import the prebuilt reference extension (needs torch; raises if it was never built)."""
    so = built_so()
    if so is None:
        raise RuntimeError('oracle/_ref/_gridencoder_ref*.so missing: run `python -m oracle.build_ref` where /root/reference exists')
    import importlib.util
    import torch  # noqa: F401  (libtorch symbols must be loaded first)
    spec = importlib.util.spec_from_file_location(NAME, so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == '__main__':
    print(build(force='-f' in sys.argv, verbose='-v' in sys.argv))
    print(build_raster(force='-f' in sys.argv))
