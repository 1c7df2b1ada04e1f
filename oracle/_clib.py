"""ctypes loader for the C oracle (oracle/oracle_c.c).  TEST INFRASTRUCTURE ONLY."""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, '_build', 'liboracle.so')
_lib = None


def build(force=False):
    """Compile oracle_c.c with the committed Makefile (gcc, -ffp-contract=off)."""
    src = os.path.join(_HERE, 'oracle_c.c')
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(['make', '-C', _HERE, '-s'])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.orc_spec_expf.restype = ctypes.c_float
        _lib.orc_spec_expf.argtypes = [ctypes.c_float]
        _lib.orc_raster_preprocess.restype = ctypes.c_int64
    return _lib


class OrcCamera(ctypes.Structure):
    _fields_ = [('H', ctypes.c_int), ('W', ctypes.c_int),
                ('tanfovx', ctypes.c_float), ('tanfovy', ctypes.c_float),
                ('view', ctypes.c_float * 16), ('proj', ctypes.c_float * 16),
                ('bg', ctypes.c_float * 3), ('scale_modifier', ctypes.c_float)]


def ptr(a):
    """numpy array (C-contiguous) -> void*; None -> NULL."""
    if a is None:
        return ctypes.c_void_p(0)
    assert a.flags['C_CONTIGUOUS']
    return ctypes.c_void_p(a.ctypes.data)
