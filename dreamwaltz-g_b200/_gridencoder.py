"""Drop-in replacement for the reference's ``_gridencoder`` pybind module
(core/nerf/gridencoder/src/bindings.cpp:5-9, gridencoder.h:12-15): same function names and
argument order, writing into caller-owned tensors, backed by libdwg_sm100.so.

Only the configuration the reference uses is supported: D = 3, C = 2, float32 (the avatar grid,
nerf_model.py:223-231).  Anything else raises RuntimeError (the reference raises
std::runtime_error for unsupported C/D, gridencoder.cu:378,395).
"""
import numpy as np
import torch

from dwg._lib import check, lib, ptr, stream


_TABLE_CACHE = {}


def _tables(offsets, L, S, H, device):
    """Per-level (scale, resolution) device tables, evaluated once per configuration by dwg_grid_level_table."""
    from dwg.ops import device_level_table
    key = (int(L), float(np.float32(S)), int(H), str(device))
    if key not in _TABLE_CACHE:
        _TABLE_CACHE[key] = device_level_table(np.float32(S), H, L, device)
    return _TABLE_CACHE[key]


def _check(inputs, embeddings, D, C):
    if D != 3 or C != 2:
        raise RuntimeError('dwg _gridencoder: only D == 3, C == 2 are supported')
    if not (inputs.is_cuda and embeddings.is_cuda and inputs.is_contiguous() and embeddings.is_contiguous()):
        raise RuntimeError('dwg _gridencoder: tensors must be contiguous CUDA tensors')
    if inputs.dtype != torch.float32 or embeddings.dtype != torch.float32:
        raise RuntimeError('dwg _gridencoder: float32 only')


def grid_encode_forward(inputs, embeddings, offsets, outputs, B, D, C, L, S, H, dy_dx, gridtype, align_corners, interp):
    """inputs [B,D] in [0,1]; outputs [L,B,C] (written); dy_dx [B,L*D*C] or None (written)."""
    _check(inputs, embeddings, D, C)
    scale, res = _tables(offsets, L, S, H, inputs.device)
    # bound <= 0 tells the C ABI that inputs are already in [0,1]
    check(lib().dwg_grid_encode_fwd(ptr(inputs), 0.0, ptr(embeddings), ptr(offsets), ptr(scale), ptr(res),
                                    ptr(outputs), C, B * C, ptr(dy_dx), B, L, int(gridtype), int(bool(align_corners)),
                                    int(interp), stream()), 'grid_encode_forward')


def grid_encode_backward(grad, inputs, embeddings, offsets, grad_embeddings, B, D, C, L, S, H, dy_dx, grad_inputs,
                         gridtype, align_corners, interp):
    """grad [L,B,C]; grad_embeddings accumulated in place; grad_inputs [B,D] written when dy_dx is given."""
    _check(inputs, embeddings, D, C)
    scale, res = _tables(offsets, L, S, H, inputs.device)
    gx = grad_inputs if (dy_dx is not None and grad_inputs is not None) else None
    check(lib().dwg_grid_encode_bwd(ptr(grad), C, B * C, ptr(inputs), 0.0, ptr(embeddings), ptr(offsets),
                                    ptr(scale), ptr(res), ptr(grad_embeddings), ptr(gx), B, L, int(gridtype),
                                    int(bool(align_corners)), int(interp), stream()), 'grid_encode_backward')


def grad_total_variation(*args, **kwargs):
    raise RuntimeError('dwg _gridencoder: grad_total_variation is a stage-I NeRF regulariser, out of scope '
                       '(SURVEY.md section 8: not on the 3DGS SDS step)')
