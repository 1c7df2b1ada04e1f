"""Oracle: tile-based EWA Gaussian rasteriser forward + backward (rows R11-R13).

TEST INFRASTRUCTURE (see oracle/__init__.py).  PARITY UNPINNED: restates the un-vendored
third-party ashawkey/diff-gaussian-rasterization (git HEAD, reference scripts/install.sh:30;
call site core/gaussian/gaussian_renderer.py:186-195) from its published algorithm
(SURVEY.md appendix B).  The arithmetic lives in oracle/oracle_c.c; this file marshals numpy
arrays and mirrors the Python surface ``rasterizer(means3D, means2D, opacities, shs,
colors_precomp, scales, rotations, cov3D_precomp) -> (color, radii, depth, alpha)``.
"""
import ctypes

import numpy as np

from ._clib import OrcCamera, lib, ptr

TILE = 16


def make_camera(H, W, tanfovx, tanfovy, viewmatrix, projmatrix, bg, scale_modifier=1.0):
    """viewmatrix / projmatrix: [4,4] row-vector matrices (= extrinsic^T, (proj @ extrinsic)^T),
    i.e. exactly GaussianRasterizationSettings.viewmatrix / .projmatrix
    (gaussian_renderer.py:36-37)."""
    cam = OrcCamera()
    cam.H, cam.W = int(H), int(W)
    cam.tanfovx, cam.tanfovy = float(tanfovx), float(tanfovy)
    v = np.ascontiguousarray(np.asarray(viewmatrix, np.float32)).reshape(16)
    p = np.ascontiguousarray(np.asarray(projmatrix, np.float32)).reshape(16)
    for i in range(16):
        cam.view[i] = float(v[i]); cam.proj[i] = float(p[i])
    for i in range(3):
        cam.bg[i] = float(bg[i])
    cam.scale_modifier = float(scale_modifier)
    return cam


def f32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def forward(cam, means3D, scales, rots, opacities, colors):
    """Full forward.  Returns a dict with every intermediate the parity tests compare:
    radii, xy, depth, cov3D, conic_opacity, rect, tiles_touched, P, keys, vals, ranges,
    color [3,H,W], out_depth [H,W], out_alpha [H,W], final_T, n_contrib."""
    L = lib()
    means3D, scales, rots, colors = f32(means3D), f32(scales), f32(rots), f32(colors)
    opacities = f32(opacities).reshape(-1)
    N = means3D.shape[0]
    H, W = cam.H, cam.W
    gx, gy = (W + TILE - 1) // TILE, (H + TILE - 1) // TILE
    o = {
        'radii': np.zeros(N, np.int32), 'xy': np.zeros((N, 2), np.float32), 'depth': np.zeros(N, np.float32),
        'cov3D': np.zeros((N, 6), np.float32), 'conic_opacity': np.zeros((N, 4), np.float32),
        'rect': np.zeros((N, 4), np.int32), 'tiles_touched': np.zeros(N, np.uint32),
    }
    P = L.orc_raster_preprocess(ctypes.c_int(N), ptr(means3D), ptr(scales), ptr(rots), ptr(opacities),
                                ctypes.byref(cam), ptr(o['radii']), ptr(o['xy']), ptr(o['depth']),
                                ptr(o['cov3D']), ptr(o['conic_opacity']), ptr(o['rect']), ptr(o['tiles_touched']))
    o['P'] = int(P)
    o['keys'] = np.zeros(max(P, 1), np.uint64)
    o['vals'] = np.zeros(max(P, 1), np.uint32)
    o['ranges'] = np.zeros((gx * gy, 2), np.uint32)
    L.orc_raster_bin(ctypes.c_int(N), ctypes.byref(cam), ptr(o['depth']), ptr(o['rect']), ptr(o['tiles_touched']),
                     ctypes.c_int64(P), ptr(o['keys']), ptr(o['vals']), ptr(o['ranges']))
    o['keys'], o['vals'] = o['keys'][:P], o['vals'][:P]
    o['color'] = np.zeros((3, H, W), np.float32)
    o['out_depth'] = np.zeros((H, W), np.float32)
    o['out_alpha'] = np.zeros((H, W), np.float32)
    o['final_T'] = np.zeros((H, W), np.float32)
    o['n_contrib'] = np.zeros((H, W), np.uint32)
    vals = o['vals'] if P > 0 else np.zeros(1, np.uint32)
    L.orc_raster_render(ctypes.byref(cam), ptr(o['ranges']), ptr(vals), ptr(o['xy']), ptr(o['conic_opacity']),
                        ptr(colors), ptr(o['depth']), ptr(o['color']), ptr(o['out_depth']), ptr(o['out_alpha']),
                        ptr(o['final_T']), ptr(o['n_contrib']))
    o['_inputs'] = (means3D, scales, rots, opacities, colors)
    return o


def backward(cam, fwd, dL_dcolor, dL_ddepth=None, dL_dalpha=None):
    """Backward from pixel gradients to per-Gaussian gradients.  Returns dict of float32 arrays:
    means3D [N,3], means2D [N,3] (z = 0), colors [N,3], opacities [N,1], scales [N,3], rots [N,4],
    plus the intermediates conic [N,3] and depth [N]."""
    L = lib()
    means3D, scales, rots, opacities, colors = fwd['_inputs']
    N = means3D.shape[0]
    H, W = cam.H, cam.W
    dL_dcolor = f32(dL_dcolor).reshape(3, H, W)
    dL_ddepth = None if dL_ddepth is None else f32(dL_ddepth).reshape(H, W)
    dL_dalpha = None if dL_dalpha is None else f32(dL_dalpha).reshape(H, W)
    g2 = np.zeros((N, 2), np.float64); gc = np.zeros((N, 3), np.float64)
    go = np.zeros(N, np.float64); gcol = np.zeros((N, 3), np.float64); gd = np.zeros(N, np.float64)
    vals = fwd['vals'] if fwd['P'] > 0 else np.zeros(1, np.uint32)
    L.orc_raster_render_backward(ctypes.byref(cam), ptr(fwd['ranges']), ptr(vals), ptr(fwd['xy']),
                                 ptr(fwd['conic_opacity']), ptr(colors), ptr(fwd['depth']),
                                 ptr(fwd['final_T']), ptr(fwd['n_contrib']),
                                 ptr(dL_dcolor), ptr(dL_ddepth), ptr(dL_dalpha), ctypes.c_int(N),
                                 ptr(g2), ptr(gc), ptr(go), ptr(gcol), ptr(gd))
    g2f, gcf, gdf = f32(g2), f32(gc), f32(gd)
    gm = np.zeros((N, 3), np.float32); gs = np.zeros((N, 3), np.float32); gr = np.zeros((N, 4), np.float32)
    L.orc_raster_preprocess_backward(ctypes.c_int(N), ptr(means3D), ptr(scales), ptr(rots), ctypes.byref(cam),
                                     ptr(fwd['radii']), ptr(fwd['cov3D']), ptr(g2f), ptr(gcf), ptr(gdf),
                                     ptr(gm), ptr(gs), ptr(gr))
    m2 = np.zeros((N, 3), np.float32); m2[:, :2] = g2f
    return {'means3D': gm, 'means2D': m2, 'colors': f32(gcol), 'opacities': f32(go).reshape(N, 1),
            'scales': gs, 'rots': gr, 'conic': gcf, 'depth': gdf}


def spec_expf(x):
    L = lib()
    return np.array([L.orc_spec_expf(ctypes.c_float(float(v))) for v in np.asarray(x, np.float32).reshape(-1)],
                    np.float32).reshape(np.shape(x))
