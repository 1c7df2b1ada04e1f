"""Oracle: multi-resolution tiled/hash grid encoder (row R6).

TEST INFRASTRUCTURE (see oracle/__init__.py).  PINNED (round 2): the reference kernel is CUDA-only
(core/nerf/gridencoder/src/gridencoder.cu:66-366; module core/nerf/gridencoder/grid.py:99-166), so its own
sources were compiled unmodified for sm_100a (oracle/build_ref.py) and run on a B200
(tests/golden/make_grid_golden.py); tests/test_grid_golden.py holds this oracle to those outputs (forward,
dy_dx, grad_inputs, grad_embeddings incl. the bit-exact set of touched table rows, 4 configurations).
Arithmetic in oracle/oracle_c.c; table construction (offsets, per-level scale/resolution)
restated here from grid.py:120-133 and gridencoder.cu:137-139.
"""
import ctypes
import os

import numpy as np

from ._clib import lib, ptr


def level_table(num_levels=16, level_dim=2, base_resolution=16, log2_hashmap_size=19,
                desired_resolution=4096, per_level_scale=None, input_dim=3, align_corners=False):
    """grid.py:104-133 -> (offsets int32 [L+1], per_level_scale float, S float32,
    level_scale float32 [L], level_res uint32 [L]).

    level_scale[l] = exp2f(l*S)*H - 1 and level_res[l] = ceil(scale)+1 are the constants the
    reference kernel recomputes per thread (gridencoder.cu:137-139); they are computed once
    on the host in float32 and handed to both oracle and CUDA kernel."""
    if desired_resolution is not None:
        per_level_scale = np.exp2(np.log2(desired_resolution / base_resolution) / (num_levels - 1))
    offsets, offset = [], 0
    max_params = 2 ** log2_hashmap_size
    for i in range(num_levels):
        resolution = int(np.ceil(base_resolution * per_level_scale ** i))
        params = min(max_params, (resolution if align_corners else resolution + 1) ** input_dim)
        params = int(np.ceil(params / 8) * 8)
        offsets.append(offset)
        offset += params
    offsets.append(offset)
    S = np.float32(np.log2(per_level_scale))
    lv = np.arange(num_levels, dtype=np.float32)
    scale = (np.exp2(lv * S).astype(np.float32) * np.float32(base_resolution) - np.float32(1.0)).astype(np.float32)
    # The reference evaluates exp2f(level * S) on the GPU (MUFU.EX2 based, <= 2 ulp), a host libm exp2 is correctly
    # rounded: they differ by one ulp at some levels, which moves outputs by up to 1e-4 at the finest levels.  For the
    # avatar's configuration the GPU-evaluated constants are part of the committed fixture and are used here.
    if (num_levels, base_resolution, desired_resolution, log2_hashmap_size, input_dim, bool(align_corners)) == (16, 16, 4096, 19, 3, False):
        gp = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden', 'grid.npz')
        if os.path.exists(gp):
            with np.load(gp) as z:
                if 'avatar.level_scale' in z.files:
                    scale = z['avatar.level_scale'].astype(np.float32)
    res = (np.ceil(scale).astype(np.uint32) + np.uint32(1)).astype(np.uint32)
    return np.array(offsets, np.int32), float(per_level_scale), S, scale, res


def forward(x, table, offsets, level_scale, level_res, bound=2.0, gridtype=1, align_corners=False,
            interp=1, want_dy_dx=True, want_index=False):
    """x [B,3] world positions in [-bound, bound]; table [rows,C] -> (enc [B,L*C], dy_dx, idx)."""
    x = np.asarray(x, np.float32)
    x01 = np.ascontiguousarray(((x + np.float32(bound)) / np.float32(2 * bound)).astype(np.float32))
    table = np.ascontiguousarray(np.asarray(table, np.float32))
    B, L, C = x01.shape[0], len(level_scale), table.shape[1]
    out = np.zeros((B, L * C), np.float32)
    dy = np.zeros((B, L * 3 * C), np.float32) if want_dy_dx else None
    idx = np.zeros((B, L, 8), np.uint32) if want_index else None
    lib().orc_grid_forward(ptr(x01), ptr(table), ptr(np.ascontiguousarray(offsets, np.int32)),
                           ptr(np.ascontiguousarray(level_scale, np.float32)),
                           ptr(np.ascontiguousarray(level_res, np.uint32)),
                           ctypes.c_int(B), ctypes.c_int(L), ctypes.c_int(C), ctypes.c_int(gridtype),
                           ctypes.c_int(int(align_corners)), ctypes.c_int(interp), ptr(out), ptr(dy), ptr(idx))
    return out, dy, idx


def backward(grad, x, table_shape, offsets, level_scale, level_res, dy_dx=None, bound=2.0, gridtype=1,
             align_corners=False, interp=1):
    """grad [B,L*C] -> (grad_table [rows,C] float64-accumulated, grad_x [B,3] wrt world x)."""
    x = np.asarray(x, np.float32)
    x01 = np.ascontiguousarray(((x + np.float32(bound)) / np.float32(2 * bound)).astype(np.float32))
    grad = np.ascontiguousarray(np.asarray(grad, np.float32))
    B, L, C = x01.shape[0], len(level_scale), table_shape[1]
    gt = np.zeros(table_shape, np.float64)
    gx = np.zeros((B, 3), np.float32) if dy_dx is not None else None
    lib().orc_grid_backward(ptr(grad), ptr(x01), ptr(np.ascontiguousarray(offsets, np.int32)),
                            ptr(np.ascontiguousarray(level_scale, np.float32)),
                            ptr(np.ascontiguousarray(level_res, np.uint32)),
                            ctypes.c_int(B), ctypes.c_int(L), ctypes.c_int(C), ctypes.c_int(gridtype),
                            ctypes.c_int(int(align_corners)), ctypes.c_int(interp), ptr(gt),
                            ptr(None if dy_dx is None else np.ascontiguousarray(dy_dx, np.float32)), ptr(gx))
    if gx is not None:
        gx = gx / np.float32(2 * bound)            # chain through (x + bound) / (2 bound), grid.py:153
    return gt, gx
