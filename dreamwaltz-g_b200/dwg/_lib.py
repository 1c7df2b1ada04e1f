"""ctypes binding of libdwg_sm100.so (the C ABI declared in include/dwg.h).

There is NO fallback: if the library is missing or a call fails, a RuntimeError is raised (the
reference trainer's ``except RuntimeError`` checkpoint-and-exit path, core/trainer.py:919-923,
keeps working).  torch is used only for device memory and the current stream.
"""
import ctypes
import os

import torch

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO_PATH = os.environ.get('DWG_SO') or os.path.join(_PKG, 'libdwg_sm100.so')      # DWG_SO: A/B against another build of the same ABI (tools only)
_lib = None

c_void_p, c_int, c_int64, c_float = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float


class DwgRasterCamera(ctypes.Structure):
    _fields_ = [('image_height', ctypes.c_int32), ('image_width', ctypes.c_int32),
                ('tanfovx', c_float), ('tanfovy', c_float),
                ('viewmatrix', c_float * 16), ('projmatrix', c_float * 16),
                ('bg', c_float * 3), ('scale_modifier', c_float)]


# name -> (restype, argtypes); this table is also what tests use to check the exported symbols
SIGNATURES = {
    'dwg_last_error': (ctypes.c_char_p, []),
    'dwg_version': (c_int, []),
    'dwg_device_cc': (c_int, []),
    'dwg_lbs_skin_fwd': (c_int, [c_void_p] * 6 + [c_int64, c_int, c_void_p]),
    'dwg_lbs_skin_bwd': (c_int, [c_void_p] * 10 + [c_int64, c_int, c_void_p]),
    'dwg_sh_eval_fwd': (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    'dwg_sh_eval_bwd': (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                c_int64, c_void_p]),
    'dwg_grid_level_table': (c_int, [c_float, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    'dwg_grid_encode_fwd': (c_int, [c_void_p, c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64,
                                    c_int64, c_void_p, c_int64, c_int, c_int, c_int, c_int, c_void_p]),
    'dwg_grid_encode_bwd': (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_float, c_void_p, c_void_p, c_void_p,
                                    c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_int, c_void_p]),
    'dwg_raster_geom_bytes': (c_int64, [c_int64]),
    'dwg_raster_bin_bytes': (c_int64, [c_int64, c_int, c_int]),
    'dwg_raster_img_bytes': (c_int64, [c_int, c_int]),
    'dwg_raster_bwd_scratch_bytes': (c_int64, [c_int64]),
    'dwg_raster_forward': (c_int, [ctypes.POINTER(DwgRasterCamera), c_int64] + [c_void_p] * 9 +
                           [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    'dwg_raster_backward': (c_int, [ctypes.POINTER(DwgRasterCamera), c_int64] + [c_void_p] * 5 +
                            [c_void_p, c_void_p, c_int64, c_void_p] + [c_void_p] * 3 + [c_void_p] * 6 +
                            [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    'dwg_gemm_f16': (c_int, [c_void_p, c_int64, c_int64, c_int64, c_void_p, c_int64, c_int64, c_int64,
                              c_void_p, c_int64, c_int64, c_int64, c_int, c_int, c_int, c_int, c_int, c_int,
                              c_void_p, c_void_p, c_int, c_void_p, c_int64, c_int64, c_int64, c_float, c_int, c_void_p]),
    'dwg_gemm_workspace_bytes': (c_int64, []),
    'dwg_gemm_f16_ws': (c_int, [c_void_p, c_int64, c_int64, c_int64, c_void_p, c_int64, c_int64, c_int64,
                                c_void_p, c_int64, c_int64, c_int64, c_int, c_int, c_int, c_int, c_int, c_int,
                                c_void_p, c_void_p, c_int, c_void_p, c_int64, c_int64, c_int64, c_float, c_int, c_void_p, c_int64, c_void_p, c_int, c_void_p]),
    'dwg_gemm_f16_ln': (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_int64, c_int64, c_void_p, c_int64, c_int64, c_int, c_int, c_int, c_int,
                                c_void_p, c_void_p, c_int64, c_int64, c_int,
                                c_int, c_void_p, c_void_p, c_void_p, c_int, c_float, c_void_p, c_void_p, c_int64, c_void_p, c_int, c_void_p]),
    'dwg_conv2d_nhwc_f16_ws': (c_int, [c_void_p, c_void_p, c_void_p, c_int] + [c_int] * 11 +
                                [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int64, c_void_p, c_void_p]),
    'dwg_avatar_mlp_param_count': (c_int64, []),
    'dwg_avatar_mlp_set_tc': (c_int, [c_int]),
    'dwg_avatar_mlp_bwd_scratch_bytes': (c_int64, [c_int64]),
    'dwg_avatar_mlp_scratch_bytes': (c_int64, []),
    'dwg_avatar_mlp_fwd': (c_int, [c_void_p] * 11 + [c_int64, c_int64, c_float, c_float, c_float, c_void_p]),
    'dwg_avatar_mlp_bwd': (c_int, [c_void_p] * 16 + [c_int64, c_int64, c_float, c_float, c_void_p]),
    'dwg_gemm_tune': (c_int, [c_int, c_int]),
    'dwg_gemm_last_plan': (c_int, [c_void_p]),
    'dwg_gemm_tune_pair': (c_int, [c_int]),
    'dwg_gemm_tune_halo': (c_int, [c_int, c_int]),
    'dwg_gemm_last_halo': (c_int, []),
    'dwg_gemm_last_pair': (c_int, []),
    'dwg_gemm_last_key': (c_int, [c_void_p]),
    'dwg_gemm_trace': (c_int, [c_void_p]),
    'dwg_raster_probe': (c_int, [c_void_p]),
    'dwg_gemm_set_lane': (c_int, [c_int]),
    'dwg_conv2d_nhwc_f16': (c_int, [c_void_p, c_void_p, c_void_p, c_int] + [c_int] * 11 +
                             [c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    'dwg_nn_set_carveout': (c_int, [c_int]),
    'dwg_groupnorm_last_launches': (c_int, []),
    'dwg_groupnorm_set_fused': (c_int, [c_int]),
    'dwg_groupnorm_fwd': (c_int, [c_void_p] * 5 + [c_int, c_int, c_int, c_int, c_float, c_int, c_void_p]),
    'dwg_groupnorm_apply_cs': (c_int, [c_void_p] * 6 + [c_int, c_int, c_int, c_int, c_float, c_int, c_void_p]),
    'dwg_nchw_f32_to_nhwc_f16': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int64, c_int, c_float, c_float, c_void_p]),
    'dwg_nhwc_f16_to_nchw_f32': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int64, c_int, c_float, c_void_p]),
    'dwg_groupnorm_apply_cs2': (c_int, [c_void_p, c_void_p, c_int] + [c_void_p] * 6 + [c_int, c_int, c_int, c_int, c_float, c_int, c_void_p]),
    'dwg_groupnorm_bwd': (c_int, [c_void_p] * 8 + [c_int, c_int, c_int, c_int, c_float, c_int, c_void_p]),
    'dwg_layernorm_fwd': (c_int, [c_void_p] * 4 + [c_int64, c_int, c_float, c_void_p]),
    'dwg_softmax_rows': (c_int, [c_void_p, c_int64, c_int, c_int, c_void_p]),
    'dwg_softmax_rows_bwd': (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p]),
    'dwg_geglu': (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p]),
    'dwg_eltwise_f16': (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p]),
    'dwg_sds_grad': (c_int, [c_void_p] * 5 + [c_float, c_float, c_int64, c_void_p]),
    'dwg_attention_fwd': (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int64, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                  c_float, c_void_p]),
    'dwg_glbs_joints': (c_int, [c_void_p] * 9 + [c_int, c_void_p, c_int] + [c_void_p] * 11),
    'dwg_glbs_vertices': (c_int, [c_int, c_int] + [c_void_p] * 10),
    'dwg_mesh_gaussians_fwd': (c_int, [c_int, c_int, c_int] + [c_void_p] * 11),
    'dwg_mesh_gaussians_bwd': (c_int, [c_int, c_int] + [c_void_p] * 11),
    'dwg_pose_keypoints_2d': (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_float, c_float, c_float, c_float, c_void_p, c_void_p, c_int, c_int, c_float, c_float,
                                      c_float, c_float, c_float, c_void_p, c_void_p]),
    'dwg_pose_image': (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    'dwg_adam_step': (c_int, [c_void_p] * 4 + [c_int64, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    'dwg_frame_pack': (c_int, [c_void_p] * 8 + [c_int, c_int, c_float, c_void_p]),
    'dwg_raster_view': (c_void_p, [c_int, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int, c_int]),
}


def lib():
    """Load libdwg_sm100.so (raises RuntimeError if it has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise RuntimeError(f'{SO_PATH} not found: build it with `python -m dwg.build` '
                               '(__graft_entry__.build()); there is no CPU or PyTorch fallback')
        L = ctypes.CDLL(SO_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = _Counting(L)
    return _lib


# kernels launched by one call of each entry point (memsets not counted)
KERNELS_PER_CALL = {
    'dwg_lbs_skin_fwd': 1, 'dwg_lbs_skin_bwd': 1, 'dwg_sh_eval_fwd': 1, 'dwg_sh_eval_bwd': 1,
    'dwg_grid_encode_fwd': 1, 'dwg_grid_encode_bwd': 1, 'dwg_avatar_mlp_fwd': 1, 'dwg_avatar_mlp_bwd': 2, 'dwg_raster_forward': 12, 'dwg_raster_backward': 2,
    'dwg_gemm_f16': 1, 'dwg_conv2d_nhwc_f16': 1, 'dwg_gemm_f16_ws': 1, 'dwg_gemm_f16_ln': 1, 'dwg_conv2d_nhwc_f16_ws': 1, 'dwg_groupnorm_fwd': 2, 'dwg_groupnorm_bwd': 2, 'dwg_groupnorm_apply_cs': 1, 'dwg_groupnorm_apply_cs2': 1, 'dwg_nchw_f32_to_nhwc_f16': 1, 'dwg_nhwc_f16_to_nchw_f32': 1,
    'dwg_layernorm_fwd': 1, 'dwg_softmax_rows': 1, 'dwg_softmax_rows_bwd': 1, 'dwg_geglu': 1,
    'dwg_eltwise_f16': 1, 'dwg_sds_grad': 1, 'dwg_attention_fwd': 1, 'dwg_adam_step': 2, 'dwg_grid_level_table': 1, 'dwg_frame_pack': 1, 'dwg_glbs_joints': 1, 'dwg_glbs_vertices': 1,
    'dwg_mesh_gaussians_fwd': 2, 'dwg_mesh_gaussians_bwd': 1, 'dwg_pose_keypoints_2d': 1, 'dwg_pose_image': 1,
}


class _Counting:
    """Thin proxy over the CDLL that counts kernel launches issued through the C ABI (bench.py's
    ``gpu_launches``)."""

    def __init__(self, cdll):
        object.__setattr__(self, '_cdll', cdll)
        object.__setattr__(self, 'launches', 0)
        object.__setattr__(self, 'flops', 0.0)

    def __getattr__(self, name):
        fn = getattr(self._cdll, name)
        k = KERNELS_PER_CALL.get(name, 0)
        if k == 0:
            return fn

        def wrapped(*a):
            object.__setattr__(self, 'launches', self.launches + k)
            return fn(*a)
        return wrapped


def check(rc, what=''):
    if rc != 0:
        msg = lib().dwg_last_error().decode(errors='replace')
        raise RuntimeError(f'libdwg {what} failed ({rc}): {msg}')


def ptr(t):
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), 'libdwg needs contiguous CUDA tensors'
    return t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


def f32c(t):
    """float32 + contiguous view/copy of a tensor (no-op when already so)."""
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()
