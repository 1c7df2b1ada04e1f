"""Drop-in replacement for the third-party ``diff_gaussian_rasterization`` package (ashawkey
fork) that the reference imports at core/gaussian/gaussian_renderer.py:5 -- same names, same
call signature and return tuple, backed by libdwg_sm100.so (no torch extension, no fallback).

    from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer
    color, radii, depth, alpha = GaussianRasterizer(settings)(means3D, means2D, opacities,
        shs=None, colors_precomp=colors, scales=scales, rotations=quats, cov3D_precomp=None)
"""
from typing import NamedTuple

import torch
import torch.nn as nn

from dwg import ops as _ops


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool


class GaussianRasterizer(nn.Module):
    def __init__(self, raster_settings: GaussianRasterizationSettings):
        super().__init__()
        self.raster_settings = raster_settings
        self.last_state = None          # RasterState of the most recent forward (parity / debugging)

    def markVisible(self, positions):
        """Frustum test of the upstream mark_visible: view-space z > 0.2."""
        with torch.no_grad():
            v = self.raster_settings.viewmatrix.to(positions.device)
            z = positions @ v[:3, 2] + v[3, 2]
            return z > 0.2

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None):
        rs = self.raster_settings
        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception('Please provide excatly one of either SHs or precomputed colors!')
        if ((scales is None or rotations is None) and cov3D_precomp is None) or \
                ((scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')
        if cov3D_precomp is not None:
            raise RuntimeError('dwg rasteriser: cov3D_precomp is not supported (the DreamWaltz-G path always '
                               'passes scales+rotations, gaussian_renderer.py:167-173 defaults)')
        if shs is not None:
            # SH colours are evaluated by the stand-alone SH kernel (same clamp semantics and the
            # same dL/dmeans3D contribution through the view direction as the upstream in-kernel path)
            colors_precomp = _ops.sh_colors(shs, means3D, rs.campos.to(means3D.device), rs.sh_degree + 1)
        states = []
        color, radii, depth, alpha = _ops.rasterize(
            means3D, means2D, colors_precomp, opacities, scales, rotations,
            image_height=rs.image_height, image_width=rs.image_width, tanfovx=rs.tanfovx, tanfovy=rs.tanfovy,
            viewmatrix=rs.viewmatrix, projmatrix=rs.projmatrix, bg=rs.bg, scale_modifier=rs.scale_modifier,
            state_out=states)
        self.last_state = states[0]
        return color, radii, depth, alpha
