"""CPU: the C-ABI library loads and exports every symbol include/dwg.h declares; the product
never imports the oracle; the product refuses to run without its CUDA library."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, 'dreamwaltz-g_b200')


def _declared_symbols():
    src = open(os.path.join(ROOT, 'include', 'dwg.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(dwg_[a-z0-9_]+)\s*\(', src)))


def test_library_builds_and_exports_every_declared_symbol():
    import ctypes
    from dwg import build
    so = build.build()
    L = ctypes.CDLL(so)
    names = _declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(L, n), f'{n} declared in include/dwg.h but not exported'
    assert L.dwg_version() >= 100


def test_ctypes_table_covers_the_header():
    from dwg import _lib
    assert set(_declared_symbols()) == set(_lib.SIGNATURES), set(_declared_symbols()) ^ set(_lib.SIGNATURES)


def test_product_never_imports_the_oracle():
    bad = []
    for d, _, files in os.walk(PKG):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h', '.cpp')):
                s = open(os.path.join(d, f)).read()
                if re.search(r'^\s*(from|import)\s+oracle\b', s, flags=re.M) or 'oracle_c.c' in s and f.endswith('.py'):
                    bad.append(os.path.join(d, f))
    assert not bad, bad


def test_missing_library_fails_loudly(monkeypatch):
    from dwg import _lib
    monkeypatch.setattr(_lib, '_lib', None)
    monkeypatch.setattr(_lib, 'SO_PATH', os.path.join(PKG, 'does_not_exist.so'))
    with pytest.raises(RuntimeError):
        _lib.lib()


def test_bad_arguments_return_error_codes_without_a_gpu():
    from dwg import _lib
    L = _lib.lib()
    rc = L.dwg_lbs_skin_fwd(None, None, None, None, None, None, 10, 55, None)
    assert rc == -1 and b'null' in L.dwg_last_error()
    with pytest.raises(RuntimeError):
        _lib.check(rc, 'dwg_lbs_skin_fwd')
    rc = L.dwg_grid_encode_fwd(1, 2.0, 1, 1, 1, 1, 8, 32, 2, None, 4, 33, 1, 0, 1, None)
    assert rc == -1 and b'1 <= L <= 32' in L.dwg_last_error()
    assert L.dwg_raster_geom_bytes(1000) > 1000 * 56
