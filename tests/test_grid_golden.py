"""Row R6 pinned: the grid encoder against vectors produced by the REFERENCE'S OWN CUDA kernel
(core/nerf/gridencoder/src/gridencoder.cu, compiled unmodified for sm_100a by oracle/build_ref.py and run on a B200 by
tests/golden/make_grid_golden.py -> tests/golden/grid.npz).  CPU part: the C oracle (oracle/oracle_c.c orc_grid_*) vs
the golden vectors.  GPU part: dwg_grid_encode_fwd/bwd through the `_gridencoder` drop-in module (reference call
signature, grid.py:28-94) vs the same vectors.  Corner indices are compared bit-exactly through the support of
grad_embeddings (every touched table row receives a non-zero gradient by construction of the case)."""
import ctypes
import importlib.util
import os

import numpy as np
import pytest

from oracle._clib import lib as olib, ptr as optr

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, 'golden', 'grid.npz'))
_spec = importlib.util.spec_from_file_location('make_grid_golden', os.path.join(HERE, 'golden', 'make_grid_golden.py'))
mk = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(mk)
CASES = list(mk.CASES)


def _case(name):
    c = mk.CASES[name]
    L = c['L']
    pls = float(np.exp2(np.log2(c['desired'] / c['base']) / (L - 1)))
    offsets = mk.level_offsets(L, c['base'], pls, c['log2'], c['align'])
    assert np.array_equal(offsets, GOLD[f'{name}.offsets'])
    S = np.float32(np.log2(pls))
    assert abs(float(S) - float(GOLD[f'{name}.S'])) < 1e-6
    # exp2f(level * S) * H - 1 as the GPU evaluates it (stored in the fixture; a host exp2 differs by 1 ulp at some levels)
    scale = GOLD[f'{name}.level_scale'].astype(np.float32)
    lv = np.arange(L, dtype=np.float32)
    host = (np.exp2(lv * S).astype(np.float32) * np.float32(c['base']) - np.float32(1.0)).astype(np.float32)
    assert np.all(np.abs(scale.view(np.int32).astype(np.int64) - host.view(np.int32).astype(np.int64)) <= 2)
    res = (np.ceil(scale).astype(np.uint32) + np.uint32(1)).astype(np.uint32)
    x, grad = mk.case_inputs(c)
    assert np.array_equal(x, GOLD[f'{name}.x01']) and np.array_equal(grad, GOLD[f'{name}.grad'])       # the committed script regenerates the inputs
    return c, offsets, float(S), scale, res, x, grad, mk.table_pattern(int(offsets[-1]))


def _golden_dense_grad(name, shape):
    g = np.zeros(int(np.prod(shape)), np.float32)
    g[GOLD[f'{name}.grad_emb_idx']] = GOLD[f'{name}.grad_emb_val']
    return g.reshape(shape)


@pytest.mark.parametrize('name', CASES)
def test_oracle_matches_reference_kernel_golden(name):
    c, offsets, S, scale, res, x, grad, table = _case(name)
    B, L, C = x.shape[0], c['L'], 2
    out = np.zeros((B, L * C), np.float32)
    dy = np.zeros((B, L * 3 * C), np.float32)
    idx = np.zeros((B, L, 8), np.uint32)
    olib().orc_grid_forward(optr(x), optr(table), optr(offsets), optr(scale), optr(res), ctypes.c_int(B), ctypes.c_int(L), ctypes.c_int(C),
                            ctypes.c_int(c['gridtype']), ctypes.c_int(int(c['align'])), ctypes.c_int(c['interp']), optr(out), optr(dy), optr(idx))
    ref_out = GOLD[f'{name}.outputs_LBC'].transpose(1, 0, 2).reshape(B, L * C)
    np.testing.assert_allclose(out, ref_out, rtol=0, atol=2e-6)
    np.testing.assert_allclose(dy, GOLD[f'{name}.dy_dx'], rtol=2e-4, atol=2e-3)     # derivative = scale (<= 4096) x table differences
    gt = np.zeros(table.shape, np.float64)
    gx = np.zeros((B, 3), np.float32)
    olib().orc_grid_backward(optr(grad), optr(x), optr(offsets), optr(scale), optr(res), ctypes.c_int(B), ctypes.c_int(L), ctypes.c_int(C),
                             ctypes.c_int(c['gridtype']), ctypes.c_int(int(c['align'])), ctypes.c_int(c['interp']), optr(gt), optr(dy), optr(gx))
    # uint32 corner indices, bit-exact: the support of grad_embeddings IS the set of (row, channel) entries the kernel touched
    touched = np.flatnonzero(gt.reshape(-1))
    assert np.array_equal(touched, GOLD[f'{name}.grad_emb_idx'])
    np.testing.assert_allclose(gt.reshape(-1)[touched].astype(np.float32), GOLD[f'{name}.grad_emb_val'], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(gx, GOLD[f'{name}.grad_inputs'], rtol=2e-4, atol=2e-3 * max(1.0, float(np.abs(GOLD[f'{name}.grad_inputs']).max())))
    # out-of-range rows (3 and 4 of every case): zeros everywhere (gridencoder.cu:110-135)
    assert np.all(ref_out[3:5] == 0) and np.all(GOLD[f'{name}.grad_inputs'][3:5] == 0)


@pytest.mark.gpu
@pytest.mark.parametrize('name', CASES)
def test_cuda_kernel_matches_reference_kernel_golden(name):
    import torch
    import _gridencoder as ge                       # the dwg drop-in (dreamwaltz-g_b200/_gridencoder.py)
    c, offsets, S, scale, res, x, grad, table = _case(name)
    dev = 'cuda'
    B, L, C, D = x.shape[0], c['L'], 2, 3
    t = lambda a: torch.from_numpy(a).to(dev)
    xt, tt, ot = t(x), t(table), t(offsets)
    # the library's device-evaluated level table == what the reference kernel's exp2f produced (bitwise)
    from dwg.ops import device_level_table
    d_scale, d_res = device_level_table(np.float32(S), c['base'], L, dev)
    assert np.array_equal(d_scale.cpu().numpy().view(np.uint32), scale.view(np.uint32))
    assert np.array_equal(d_res.cpu().numpy().view(np.uint32), res)
    outputs = torch.empty(L, B, C, device=dev)
    dy_dx = torch.empty(B, L * D * C, device=dev)
    ge.grid_encode_forward(xt, tt, ot, outputs, B, D, C, L, S, c['base'], dy_dx, c['gridtype'], c['align'], c['interp'])
    np.testing.assert_allclose(outputs.cpu().numpy(), GOLD[f'{name}.outputs_LBC'], rtol=0, atol=2e-6)
    np.testing.assert_allclose(dy_dx.cpu().numpy(), GOLD[f'{name}.dy_dx'], rtol=2e-4, atol=2e-3)
    gl = t(grad).view(B, L, C).permute(1, 0, 2).contiguous()
    g_emb, g_in = torch.zeros_like(tt), torch.zeros_like(xt)
    ge.grid_encode_backward(gl, xt, tt, ot, g_emb, B, D, C, L, S, c['base'], dy_dx, g_in, c['gridtype'], c['align'], c['interp'])
    ge_flat = g_emb.reshape(-1).cpu().numpy()
    touched = np.flatnonzero(ge_flat)
    assert np.array_equal(touched, GOLD[f'{name}.grad_emb_idx'])                    # bit-exact corner indices
    np.testing.assert_allclose(ge_flat[touched], GOLD[f'{name}.grad_emb_val'], rtol=1e-4, atol=1e-6)
    gi = GOLD[f'{name}.grad_inputs']
    np.testing.assert_allclose(g_in.cpu().numpy(), gi, rtol=2e-4, atol=2e-3 * max(1.0, float(np.abs(gi).max())))
