"""torch.autograd wrappers over the C ABI (include/dwg.h).  Thin: allocate outputs/workspaces,
pass raw pointers + the current stream, convert error codes to RuntimeError."""
import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import check, f32c, lib, ptr, stream


# ------------------------------------------------------------------------------ LBS skinning
class _LbsSkin(torch.autograd.Function):
    """x' , q' = skin(W, A, x, q)  (dwg_lbs_skin_fwd/bwd)."""

    @staticmethod
    def forward(ctx, W, A, x, q):
        W, A, x = f32c(W), f32c(A), f32c(x)
        q = None if q is None else f32c(q)
        N, J = W.shape
        assert A.shape == (J, 4, 4) and x.shape == (N, 3)
        x_out = torch.empty_like(x)
        q_out = None if q is None else torch.empty_like(q)
        check(lib().dwg_lbs_skin_fwd(ptr(W), ptr(A), ptr(x), ptr(q), ptr(x_out), ptr(q_out), N, J, stream()),
              'dwg_lbs_skin_fwd')
        ctx.save_for_backward(W, A, x, q)
        ctx.has_q = q is not None
        if q is None:
            return x_out
        return x_out, q_out

    @staticmethod
    def backward(ctx, g_x_out, g_q_out=None):
        W, A, x, q = ctx.saved_tensors
        N, J = W.shape
        need_W, need_A = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        g_x_out = f32c(g_x_out) if g_x_out is not None else torch.zeros_like(x)
        if ctx.has_q:
            g_q_out = f32c(g_q_out) if g_q_out is not None else torch.zeros_like(q)
        g_x = torch.empty_like(x)
        g_q = torch.empty_like(q) if ctx.has_q else None
        g_W = torch.empty_like(W) if need_W else None
        g_A = torch.zeros_like(A) if need_A else None
        check(lib().dwg_lbs_skin_bwd(ptr(W), ptr(A), ptr(x), ptr(q), ptr(g_x_out), ptr(g_q_out if ctx.has_q else None),
                                     ptr(g_x), ptr(g_q), ptr(g_W), ptr(g_A), N, J, stream()), 'dwg_lbs_skin_bwd')
        return g_W, g_A, g_x, g_q


def lbs_skin(W, A, x, q=None):
    """Fused linear-blend skinning: W [N,J], A [J,4,4], x [N,3], q [N,4] or None."""
    return _LbsSkin.apply(W, A, x, q)


# ------------------------------------------------------------------------------ R1 / R5 kernels
def glbs_joints(pose_parts, pose_mean, betas, expression, J_template, JS, parents_i32, transl=None):
    """dwg_glbs_joints: pose_parts = (global_orient, body_pose, jaw, leye, reye, left_hand, right_hand) fp32 device tensors.
    -> dict(A [55,4,4], A_t [55,4,4], pose_feature [486], shape [ns], joints [55,3]).  No autograd (pose inputs carry none)."""
    dev = betas.device
    pp = [f32c(t).reshape(-1) for t in pose_parts]
    betas, expression = f32c(betas).reshape(-1), f32c(expression).reshape(-1)
    nb, ne = betas.numel(), expression.numel()
    e = lambda *s: torch.empty(*s, device=dev, dtype=torch.float32)
    out = {'A': e(55, 4, 4), 'A_t': e(55, 4, 4), 'pose_feature': e(486), 'shape': e(nb + ne), 'joints': e(55, 3), 'posed_joints': e(55, 3)}
    tr = None if transl is None else f32c(transl).reshape(-1)
    check(lib().dwg_glbs_joints(*[ptr(t) for t in pp], ptr(pose_mean), ptr(betas), nb, ptr(expression), ne, ptr(J_template), ptr(JS),
                                ptr(parents_i32), ptr(tr), ptr(out['A']), ptr(out['A_t']), ptr(out['pose_feature']), ptr(out['shape']),
                                ptr(out['joints']), ptr(out['posed_joints']), stream()), 'dwg_glbs_joints')
    out['transl'] = tr
    return out


def glbs_vertices(jt, shapedirs_sel, posedirs_sel, weights_sel, points):
    """dwg_glbs_vertices: the vertex composite transform of GLBS applied to the predefined vertices `points` [Vp,3]."""
    Vp = points.shape[0]
    out = torch.empty(Vp, 3, device=points.device, dtype=torch.float32)
    check(lib().dwg_glbs_vertices(Vp, jt['shape'].numel(), ptr(jt['shape']), ptr(jt['pose_feature']), ptr(jt['A']), ptr(jt['transl']),
                                  ptr(shapedirs_sel), ptr(posedirs_sel), ptr(weights_sel), ptr(f32c(points)), ptr(out), stream()), 'dwg_glbs_vertices')
    return out


class _MeshGaussians(torch.autograd.Function):
    """(positions, scales, quaternions) of the mesh-bound Gaussians of one part (dwg_mesh_gaussians_fwd/bwd)."""

    @staticmethod
    def forward(ctx, bary, scales_param, vertex_coords, triangles, adj_ptr, adj_tri, n_per_tri):
        b, sp, vc = f32c(bary).reshape(-1, 3), f32c(scales_param), f32c(vertex_coords)
        Vp, F, P = vc.shape[0], triangles.shape[0], b.shape[0]
        assert P == F * n_per_tri and sp.shape == (P, 3)
        e = lambda *s: torch.empty(*s, device=vc.device, dtype=torch.float32)
        vn, pos, sc, q = e(Vp, 3), e(P, 3), e(P, 3), e(P, 4)
        check(lib().dwg_mesh_gaussians_fwd(Vp, F, int(n_per_tri), ptr(vc), ptr(triangles), ptr(adj_ptr), ptr(adj_tri), ptr(b), ptr(sp), ptr(vn),
                                           ptr(pos), ptr(sc), ptr(q), stream()), 'dwg_mesh_gaussians_fwd')
        ctx.save_for_backward(b, sp, vc, vn, triangles)
        ctx.n, ctx.bshape = int(n_per_tri), bary.shape
        return pos, sc, q

    @staticmethod
    def backward(ctx, g_pos, g_sc, g_q):
        b, sp, vc, vn, triangles = ctx.saved_tensors
        gc = lambda g: None if g is None else f32c(g)
        g_b, g_sp = torch.empty_like(b), torch.empty_like(sp)
        check(lib().dwg_mesh_gaussians_bwd(triangles.shape[0], ctx.n, ptr(vc), ptr(vn), ptr(triangles), ptr(b), ptr(sp), ptr(gc(g_pos)), ptr(gc(g_sc)),
                                           ptr(gc(g_q)), ptr(g_b), ptr(g_sp), stream()), 'dwg_mesh_gaussians_bwd')
        return g_b.reshape(ctx.bshape), g_sp, None, None, None, None, None


def mesh_gaussians(bary, scales_param, vertex_coords, triangles_i32, adj_ptr, adj_tri, n_per_tri):
    return _MeshGaussians.apply(bary, scales_param, vertex_coords, triangles_i32, adj_ptr, adj_tri, n_per_tri)


# ------------------------------------------------------------------------------ SH colour
class _ShEval(torch.autograd.Function):
    @staticmethod
    def forward(ctx, sh, pos, campos, deg):
        sh, pos, campos = f32c(sh), f32c(pos), f32c(campos).reshape(3)
        N, stride = sh.shape[0], sh.shape[1]
        assert sh.shape[2] == 3 and pos.shape == (N, 3)
        rgb = torch.empty(N, 3, device=sh.device, dtype=torch.float32)
        clamped = torch.empty(N, device=sh.device, dtype=torch.uint8)
        check(lib().dwg_sh_eval_fwd(ptr(sh), stride, deg, ptr(pos), ptr(campos), ptr(rgb), ptr(clamped), N, stream()),
              'dwg_sh_eval_fwd')
        ctx.save_for_backward(sh, pos, campos, clamped)
        ctx.deg = deg
        return rgb

    @staticmethod
    def backward(ctx, g_rgb):
        sh, pos, campos, clamped = ctx.saved_tensors
        N, stride = sh.shape[0], sh.shape[1]
        g_sh = torch.empty_like(sh)
        g_pos = torch.empty_like(pos) if ctx.needs_input_grad[1] else None
        check(lib().dwg_sh_eval_bwd(ptr(sh), stride, ctx.deg, ptr(pos), ptr(campos), ptr(clamped), ptr(f32c(g_rgb)),
                                    ptr(g_sh), ptr(g_pos), N, stream()), 'dwg_sh_eval_bwd')
        return g_sh, g_pos, None, None


def sh_colors(sh_features, positions, campos, sh_levels):
    """rgb = clamp_min(eval_sh(sh_levels-1, sh, normalize(pos - campos)) + 0.5, 0); sh [N,K,3]."""
    return _ShEval.apply(sh_features, positions, campos, int(sh_levels) - 1)


# ------------------------------------------------------------------------------ grid encoder
def grid_level_table(num_levels=16, level_dim=2, base_resolution=16, log2_hashmap_size=19,
                     desired_resolution=4096, per_level_scale=None, input_dim=3, align_corners=False):
    """GridEncoder.__init__ (core/nerf/gridencoder/grid.py:104-133): level offsets, per_level_scale and
    S = log2(per_level_scale) as float32 (grid.py:40).  The per-level kernel constants (gridencoder.cu:138-139) are
    evaluated on the device by device_level_table()."""
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
    return np.array(offsets, np.int32), float(per_level_scale), np.float32(np.log2(per_level_scale))


def device_level_table(S, H, L, device):
    """(level_scale f32 [L], level_res i32-bit-pattern-of-u32 [L]) as DEVICE tensors: exp2f(l * S) * H - 1 and ceil + 1
    evaluated by the GPU exactly as reference gridencoder.cu:138-139 does (host exp2 differs by 1 ulp at some levels)."""
    scale = torch.empty(int(L), device=device, dtype=torch.float32)
    res = torch.empty(int(L), device=device, dtype=torch.int32)
    with torch.cuda.device(scale.device):
        check(lib().dwg_grid_level_table(float(np.float32(S)), int(H), int(L), ptr(scale), ptr(res), stream()), 'dwg_grid_level_table')
    return scale, res


class GridSpec:
    """Device-resident level table of one grid encoder."""

    def __init__(self, device, bound=2.0, gridtype='tiled', align_corners=False, interpolation='smoothstep', **kw):
        offsets, pls, S = grid_level_table(align_corners=align_corners, **kw)
        self.num_levels = len(offsets) - 1
        self.n_rows = int(offsets[-1])
        self.per_level_scale = pls
        self.bound = float(bound)
        self.gridtype = {'hash': 0, 'tiled': 1}[gridtype]
        self.interp = {'linear': 0, 'smoothstep': 1}[interpolation]
        self.align_corners = bool(align_corners)
        self.offsets_np = offsets
        self.offsets = torch.from_numpy(offsets).to(device)
        # per-level constants evaluated on the device with the reference kernel's own expression (dwg_grid_level_table)
        self.level_scale, self.level_res = device_level_table(S, kw.get('base_resolution', 16), self.num_levels, device)
        self.scale_np, self.res_np = self.level_scale.cpu().numpy(), self.level_res.cpu().numpy().view(np.uint32)


class _GridEncode(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, table, spec):
        x, table = f32c(x), f32c(table)
        B, L = x.shape[0], spec.num_levels
        assert x.shape[1] == 3 and table.shape == (spec.n_rows, 2)
        out = torch.empty(B, L * 2, device=x.device, dtype=torch.float32)
        check(lib().dwg_grid_encode_fwd(ptr(x), spec.bound, ptr(table), ptr(spec.offsets), ptr(spec.level_scale),
                                        ptr(spec.level_res), ptr(out), L * 2, 2, None, B, L, spec.gridtype,
                                        int(spec.align_corners), spec.interp, stream()), 'dwg_grid_encode_fwd')
        ctx.save_for_backward(x, table)
        ctx.spec = spec
        return out

    @staticmethod
    def backward(ctx, grad):
        x, table = ctx.saved_tensors
        spec = ctx.spec
        B, L = x.shape[0], spec.num_levels
        grad = f32c(grad)
        g_table = torch.zeros_like(table)
        g_x = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        check(lib().dwg_grid_encode_bwd(ptr(grad), L * 2, 2, ptr(x), spec.bound, ptr(table), ptr(spec.offsets),
                                        ptr(spec.level_scale), ptr(spec.level_res), ptr(g_table), ptr(g_x), B, L,
                                        spec.gridtype, int(spec.align_corners), spec.interp, stream()),
              'dwg_grid_encode_bwd')
        return g_x, g_table, None


def grid_encode(x, table, spec):
    """Multi-resolution grid encoding of world positions x [B,3] -> [B, 2L] (GridEncoder.forward)."""
    return _GridEncode.apply(x, table, spec)


# ------------------------------------------------------------------------------ avatar MLPs
class _AvatarMLP(torch.autograd.Function):
    """(colors [N,3], opacities [N,1], positions' [Nu,3], scales [Nu,3]) = fused sigma-net + deform-net +
    non-rigid transform (dwg_avatar_mlp_fwd/bwd).  Inputs after the scalars are the 18 parameter tensors in
    the order of AVATAR_MLP_PARAMS; layers.0.weight is split into its grid ([:, :32]) and pose ([:, 32:]) parts."""

    @staticmethod
    def forward(ctx, enc, positions, body_pose, n_unc, init_offset, init_scale, max_scale, *params):
        enc, positions = f32c(enc), f32c(positions)
        N, Nu = enc.shape[0], int(n_unc)
        assert enc.shape[1] == 32 and positions.shape == (Nu, 3) and len(params) == 18
        w0 = params[6]                                           # layers.0.weight [64, 95]
        flat = torch.cat([p.detach()[:, :32].reshape(-1) if i == 6 else p.detach().reshape(-1) for i, p in enumerate(params)])
        L = lib()
        assert flat.numel() == L.dwg_avatar_mlp_param_count()
        w_pose = w0.detach()[:, 32:].contiguous()
        pose = f32c(body_pose).reshape(-1)
        dev = enc.device
        e = lambda *s: torch.empty(*s, device=dev, dtype=torch.float32)
        colors, opac, pos, scales = e(N, 3), e(N, 1), e(Nu, 3), e(Nu, 3)
        need_bwd = any(ctx.needs_input_grad)
        Np, Nup = (N + 127) // 128 * 128, (Nu + 127) // 128 * 128          # tile-blocked [layer][tile][64][128] (include/dwg.h)
        acts_s = e(2, 64, Np) if need_bwd else None
        acts_d = e(4, 64, max(Nup, 128)) if need_bwd else None
        check(L.dwg_avatar_mlp_fwd(ptr(enc), ptr(positions), ptr(flat), ptr(w_pose), ptr(pose), ptr(colors), ptr(opac), ptr(pos), ptr(scales),
                                   ptr(acts_s), ptr(acts_d), N, Nu, float(init_offset), float(init_scale), float(max_scale), stream()),
              'dwg_avatar_mlp_fwd')
        ctx.save_for_backward(enc, flat, pose, colors, opac, scales, acts_s, acts_d)
        ctx.dims = (N, Nu, float(init_offset), float(max_scale))
        ctx.shapes = [p.shape for p in params]
        return colors, opac, pos, scales

    @staticmethod
    def backward(ctx, g_colors, g_opac, g_pos, g_scales):
        enc, flat, pose, colors, opac, scales, acts_s, acts_d = ctx.saved_tensors
        N, Nu, init_offset, max_scale = ctx.dims
        L = lib()
        gc = lambda g: None if g is None else f32c(g)
        g_colors, g_opac, g_pos, g_scales = gc(g_colors), gc(g_opac), gc(g_pos), gc(g_scales)
        g_enc = torch.empty_like(enc)
        g_flat = torch.empty_like(flat)
        g_w_pose = torch.empty(64, 63, device=enc.device, dtype=torch.float32)
        scratch = torch.empty(int(L.dwg_avatar_mlp_bwd_scratch_bytes(N)), device=enc.device, dtype=torch.uint8)
        check(L.dwg_avatar_mlp_bwd(ptr(enc), ptr(flat), ptr(pose), ptr(colors), ptr(opac), ptr(scales), ptr(acts_s), ptr(acts_d),
                                   ptr(g_colors), ptr(g_opac), ptr(g_pos), ptr(g_scales), ptr(g_enc), ptr(g_flat), ptr(g_w_pose), ptr(scratch),
                                   N, Nu, init_offset, max_scale, stream()), 'dwg_avatar_mlp_bwd')
        grads, off = [], 0
        for i, shp in enumerate(ctx.shapes):
            if i == 6:
                ge = g_flat[off:off + 64 * 32].view(64, 32)
                grads.append(torch.cat([ge, g_w_pose], dim=1))
                off += 64 * 32
            else:
                n = int(np.prod(shp))
                grads.append(g_flat[off:off + n].view(shp))
                off += n
        return (g_enc, g_pos, None, None, None, None, None, *grads)


# parameter order of the flat vector (module-relative names)
AVATAR_MLP_PARAMS = ('nerf_opacity_and_color_net.net.0.weight', 'nerf_opacity_and_color_net.net.0.bias',
                     'nerf_opacity_and_color_net.net.1.weight', 'nerf_opacity_and_color_net.net.1.bias',
                     'nerf_opacity_and_color_net.net.2.weight', 'nerf_opacity_and_color_net.net.2.bias',
                     'nerf_scale_and_quaternion_net.layers.0.weight', 'nerf_scale_and_quaternion_net.layers.0.bias',
                     'nerf_scale_and_quaternion_net.layers.1.weight', 'nerf_scale_and_quaternion_net.layers.1.bias',
                     'nerf_scale_and_quaternion_net.layers.2.weight', 'nerf_scale_and_quaternion_net.layers.2.bias',
                     'nerf_scale_and_quaternion_net.layers.3.weight', 'nerf_scale_and_quaternion_net.layers.3.bias',
                     'nerf_scale_and_quaternion_net.gaussian_warp.weight', 'nerf_scale_and_quaternion_net.gaussian_warp.bias',
                     'nerf_scale_and_quaternion_net.gaussian_scaling.weight', 'nerf_scale_and_quaternion_net.gaussian_scaling.bias')


def avatar_mlp(enc, positions, body_pose, params, n_unconstrained, init_offset=0.01, init_scale=1e-3, max_scale=0.01):
    """enc [N,32] (first n_unconstrained rows: unconstrained Gaussians), positions [Nu,3], body_pose [1,63],
    params: the 18 tensors of AVATAR_MLP_PARAMS order -> colors [N,3], opacities [N,1], positions' [Nu,3], scales [Nu,3]."""
    p = list(params)
    assert len(p) == 18
    return _AvatarMLP.apply(enc, positions, body_pose, n_unconstrained, init_offset, init_scale, max_scale, *p)


# ------------------------------------------------------------------------------ rasteriser
def _camera_struct(H, W, tanfovx, tanfovy, viewmatrix, projmatrix, bg, scale_modifier):
    cam = _lib.DwgRasterCamera()
    cam.image_height, cam.image_width = int(H), int(W)
    cam.tanfovx, cam.tanfovy = float(tanfovx), float(tanfovy)
    v = viewmatrix.detach().float().reshape(16).cpu().tolist()
    p = projmatrix.detach().float().reshape(16).cpu().tolist()
    b = bg.detach().float().reshape(3).cpu().tolist()
    for i in range(16):
        cam.viewmatrix[i], cam.projmatrix[i] = v[i], p[i]
    for i in range(3):
        cam.bg[i] = b[i]
    cam.scale_modifier = float(scale_modifier)
    return cam


CAMERA_WORDS = 40          # sizeof(DwgRasterCamera) / 4


def pack_camera(H, W, tanfovx, tanfovy, viewmatrix, projmatrix, bg, scale_modifier=1.0, out=None):
    """Host-side image of DwgRasterCamera as a float32[40] tensor (int fields bit-cast), for the
    device-resident camera used by CUDA-graph replays: ``cam_dev.copy_(pack_camera(...))``."""
    t = torch.empty(CAMERA_WORDS, dtype=torch.float32) if out is None else out
    ti = t.view(torch.int32)
    ti[0], ti[1] = int(H), int(W)
    t[2], t[3] = float(tanfovx), float(tanfovy)
    t[4:20] = viewmatrix.detach().float().reshape(16).cpu()
    t[20:36] = projmatrix.detach().float().reshape(16).cpu()
    t[36:39] = bg.detach().float().reshape(3).cpu()
    t[39] = float(scale_modifier)
    return t


class RasterState:
    """Workspaces kept between forward and backward (and inspected by the parity tests)."""
    __slots__ = ('cam', 'N', 'H', 'W', 'P_cap', 'geom', 'bin', 'img', 'status', 'radii', 'cam_dev')

    def view(self, which, dtype, shape):
        """Typed torch view of an internal buffer (dwg_raster_view)."""
        L = lib()
        p = L.dwg_raster_view(which, ptr(self.geom), ptr(self.bin), ptr(self.img), self.N, self.P_cap, self.H, self.W)
        base = {0: self.geom, 1: self.geom, 2: self.geom, 3: self.geom, 4: self.geom, 5: self.geom}.get(which)
        if base is None:
            base = self.img if which in (9, 10) else self.bin
        off = p - base.data_ptr()
        n = int(np.prod(shape))
        isz = torch.empty(0, dtype=dtype).element_size()
        return base[off:off + n * isz].view(dtype).reshape(shape)


_CAP_FLOOR = {}            # device index -> capacity learnt from an overflow report (next allocation uses it)
MAX_TILE_LOAD = 131072     # SORT_CHUNK * 2^MAX_MERGE_PASSES of csrc/raster_sort.cu: a tile holding more instances cannot be sorted


def default_instance_capacity(N, device=None):
    floor = _CAP_FLOOR.get(torch.device(device).index if device is not None else None, 0)
    return int(max(4 * N, 1 << 20, floor))


class _StatusMonitor:
    """Deferred check of the rasteriser's device status words {overflow flag, P, max tile load, -} without a host
    synchronisation: every forward copies them asynchronously into pinned host memory (a memcpy node under CUDA-graph
    capture), the NEXT forward looks at what has landed.  An overflow (P > P_cap: the kernel clamps tile ranges and drops
    instances) or an unsortable tile raises RuntimeError one call late -- upstream sizes its buffers from num_rendered
    with a host sync instead (diff_gaussian_rasterization resizeFunctional); the reference trainer's
    `except RuntimeError` path (core/trainer.py:919-923) then checkpoints and exits."""

    def __init__(self):
        self.host = {}

    def check(self, device, P_cap):
        h = self.host.get(device.index)
        if h is None:
            return
        flag, P, load = int(h[0]), int(h[1]), int(h[2])
        if flag or load > MAX_TILE_LOAD:
            h.zero_()
            if flag:
                _CAP_FLOOR[device.index] = max(_CAP_FLOOR.get(device.index, 0), int(P * 1.25) + 1024)
                raise RuntimeError(f'dwg rasteriser: instance capacity exceeded in the previous call (P = {P} (tile, Gaussian) instances > '
                                   f'capacity): instances were dropped.  Pass instance_capacity >= {P} (the next default allocation will use '
                                   f'{_CAP_FLOOR[device.index]}).')
            raise RuntimeError(f'dwg rasteriser: a tile held {load} instances in the previous call; at most {MAX_TILE_LOAD} can be depth-sorted')

    def reset(self):
        """Forget pending reports and learnt capacities (tests)."""
        for h in self.host.values():
            h.zero_()
        _CAP_FLOOR.clear()

    def post(self, status):
        dev = status.device
        h = self.host.get(dev.index)
        if h is None:
            h = self.host[dev.index] = torch.zeros(4, dtype=torch.int32).pin_memory()
        h.copy_(status, non_blocking=True)


STATUS_MONITOR = _StatusMonitor()


class _Rasterize(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, colors, opacities, scales, rotations, bg_image, cam_args, state_out):
        H, W, tanfovx, tanfovy, viewmatrix, projmatrix, bg, scale_modifier, P_cap, cam_dev = cam_args
        dev = means3D.device
        means3D, colors, scales, rotations = f32c(means3D), f32c(colors), f32c(scales), f32c(rotations)
        opac = f32c(opacities).reshape(-1)
        N = means3D.shape[0]
        L = lib()
        cam = _camera_struct(H, W, tanfovx, tanfovy, viewmatrix, projmatrix, bg, scale_modifier)
        STATUS_MONITOR.check(dev, P_cap)                       # overflow / unsortable tile reported by the previous call
        P_cap = int(P_cap or default_instance_capacity(N, dev))
        st = RasterState()
        st.cam, st.N, st.H, st.W, st.P_cap = cam, N, H, W, P_cap
        st.cam_dev = cam_dev
        u8 = lambda n: torch.empty(int(n), device=dev, dtype=torch.uint8)
        st.geom = u8(L.dwg_raster_geom_bytes(N))
        st.bin = u8(L.dwg_raster_bin_bytes(P_cap, H, W))
        st.img = u8(L.dwg_raster_img_bytes(H, W))
        st.status = torch.zeros(4, device=dev, dtype=torch.int32)
        st.radii = torch.zeros(N, device=dev, dtype=torch.int32)
        color = torch.empty(3, H, W, device=dev, dtype=torch.float32)
        depth = torch.empty(1, H, W, device=dev, dtype=torch.float32)
        alpha = torch.empty(1, H, W, device=dev, dtype=torch.float32)
        bgi = None
        color_fg = color
        if bg_image is not None:
            bgi = f32c(bg_image).reshape(3, H, W)
            color_fg = torch.empty(3, H, W, device=dev, dtype=torch.float32)
        check(L.dwg_raster_forward(ctypes.byref(cam), N, ptr(means3D), ptr(colors), ptr(opac), ptr(scales),
                                   ptr(rotations), ptr(color), ptr(depth), ptr(alpha), ptr(st.radii), ptr(st.geom),
                                   ptr(st.bin), P_cap, ptr(st.img), ptr(st.status), ptr(cam_dev), ptr(bgi),
                                   ptr(color_fg) if bgi is not None else None, stream()), 'dwg_raster_forward')
        STATUS_MONITOR.post(st.status)
        ctx.save_for_backward(means3D, colors, opac, scales, rotations, bgi, alpha if bgi is not None else None)
        ctx.st = st
        ctx.opac_shape = opacities.shape
        ctx.bg_shape = None if bg_image is None else bg_image.shape
        if state_out is not None:
            state_out.append(st)
        ctx.mark_non_differentiable(st.radii)
        if bgi is not None:
            ctx.mark_non_differentiable(color_fg)
            return color, st.radii, depth, alpha, color_fg
        return color, st.radii, depth, alpha, color.detach()

    @staticmethod
    def backward(ctx, g_color, _g_radii, g_depth, g_alpha, _g_fg):
        means3D, colors, opac, scales, rotations, bgi, alpha = ctx.saved_tensors
        st = ctx.st
        N, dev = st.N, means3D.device
        L = lib()
        g_color = f32c(g_color) if g_color is not None else torch.zeros(3, st.H, st.W, device=dev)
        g_depth = f32c(g_depth) if g_depth is not None else None
        g_alpha = f32c(g_alpha) if g_alpha is not None else None
        e = lambda *s: torch.empty(*s, device=dev, dtype=torch.float32)
        g_m3, g_m2, g_c, g_o, g_s, g_r = e(N, 3), e(N, 3), e(N, 3), e(N), e(N, 3), e(N, 4)
        g_bg = e(3, st.H, st.W) if (bgi is not None and ctx.needs_input_grad[6]) else None
        scratch = torch.empty(int(L.dwg_raster_bwd_scratch_bytes(N)), device=dev, dtype=torch.uint8)
        check(L.dwg_raster_backward(ctypes.byref(st.cam), N, ptr(means3D), ptr(colors), ptr(opac), ptr(scales),
                                    ptr(rotations), ptr(st.geom), ptr(st.bin), st.P_cap, ptr(st.img), ptr(g_color),
                                    ptr(g_depth), ptr(g_alpha), ptr(g_m3), ptr(g_m2), ptr(g_c), ptr(g_o), ptr(g_s),
                                    ptr(g_r), ptr(scratch), ptr(st.cam_dev), ptr(bgi), ptr(alpha), ptr(g_bg), stream()), 'dwg_raster_backward')
        if g_bg is not None:
            g_bg = g_bg.reshape(ctx.bg_shape)
        return g_m3, g_m2, g_c, g_o.reshape(ctx.opac_shape), g_s, g_r, g_bg, None, None


def rasterize(means3D, means2D, colors, opacities, scales, rotations, *, image_height, image_width, tanfovx,
              tanfovy, viewmatrix, projmatrix, bg, scale_modifier=1.0, instance_capacity=None, state_out=None, cam_dev=None,
              bg_image=None):
    """Differentiable tile rasteriser -> (color [3,H,W], radii i32 [N], depth [1,H,W], alpha [1,H,W]).
    With ``bg_image`` [3,H,W] the per-pixel background is composited in the blend epilogue
    (color = color_fg + bg_image * (1 - alpha), scene.py:153-166) and a 5th output color_fg is returned."""
    cam_args = (int(image_height), int(image_width), float(tanfovx), float(tanfovy), viewmatrix, projmatrix, bg,
                float(scale_modifier), instance_capacity, cam_dev)
    out = _Rasterize.apply(means3D, means2D, colors, opacities, scales, rotations, bg_image, cam_args, state_out)
    return out if bg_image is not None else out[:4]


# ------------------------------------------------------------------------------ tensor-core GEMM / conv
ACT = {None: 0, 'none': 0, 'silu': 1, 'gelu': 2, 'geglu': 3}
PROFILE = None        # set to a list to record (start_event, end_event, flops, kind) per tensor-core launch
PROFILE_BYTES = 0.0   # running sum of the compulsory operand bytes (A + B + C [+ residual], each once) of the recorded launches
TUNE_RECORD = None    # set to a list to record the arguments of every GEMM / conv call (tools/gemm_autotune.py)


class _prof:
    def __init__(self, flops, kind, nbytes=0.0):
        self.flops, self.kind, self.nbytes = flops, kind, nbytes

    def __enter__(self):
        if PROFILE is not None:
            self.e0, self.e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            self.e0.record()

    def __exit__(self, *a):
        if PROFILE is not None:
            self.e1.record()
            PROFILE.append((self.e0, self.e1, self.flops, self.kind))
            global PROFILE_BYTES
            PROFILE_BYTES += self.nbytes


_GEMM_LANE = 0
_GEMM_WS = {}


def gemm_lane(lane):
    """Select the split-K workspace (lane 0 / 1) of the GEMM / conv launches that follow.  Launches enqueued on a second,
    possibly concurrent stream must use lane 1.  Host-side state of THIS wrapper: the C ABI receives the workspace as an
    argument (dwg_gemm_f16_ws / dwg_conv2d_nhwc_f16_ws), the library itself keeps none."""
    global _GEMM_LANE
    assert lane in (0, 1, 2, 3)
    _GEMM_LANE = int(lane)


class forked:
    """`with ops.forked(helper) as f:` runs the enclosed launches on a helper stream (with its own split-K lane) beside the
    current stream -- for a launch that is independent of the next few on the main chain (a ResNet's 1x1 shortcut, the V
    projection of an attention layer); `f.join(*tensors)` makes the current stream wait before the results are used.
    helper = (stream, lane) or None (then everything simply runs in line).  Fork / join are graph edges under capture."""

    def __init__(self, helper):
        self.helper = helper

    def __enter__(self):
        if self.helper is None:
            return self
        self.stream, lane = self.helper
        self.cur = torch.cuda.current_stream()
        self.stream.wait_stream(self.cur)
        self.prev_lane = _GEMM_LANE
        self.ctx = torch.cuda.stream(self.stream)
        self.ctx.__enter__()
        gemm_lane(lane)
        return self

    def __exit__(self, *a):
        if self.helper is not None:
            gemm_lane(self.prev_lane)
            self.ctx.__exit__(*a)

    def join(self, *tensors):
        if self.helper is not None:
            self.cur.wait_stream(self.stream)
            for t in tensors:
                t.record_stream(self.cur)


def _gemm_workspace(device):
    """Caller-owned split-K scratch of (device, lane): allocated and zero-filled once (include/dwg.h)."""
    dev = torch.device(device)
    key = (dev.index if dev.index is not None else torch.cuda.current_device(), _GEMM_LANE)
    ws = _GEMM_WS.get(key)
    if ws is None:
        ws = _GEMM_WS[key] = torch.zeros(int(lib().dwg_gemm_workspace_bytes()), device=dev, dtype=torch.uint8)
    return ws


def _chk_f16(t):
    assert t.is_cuda and t.dtype == torch.float16, 'tcgen05 layers take fp16 CUDA tensors'
    return t


def _take_colstats(device, groups, N):
    """Zero-initialised i64 [4, groups, N, 2] for the epilogue's column statistics: an arena slice (zeroed by the step's one
    memset) or, outside a step, a private zero tensor."""
    cs = STATS_ARENA.take(device, 4 * groups * N * 2)
    if cs is None:
        cs = torch.zeros(4 * groups * N * 2, device=device, dtype=torch.int64)
    return cs.view(4, groups, N, 2)                             # 4 slots (include/dwg.h)


def gemm(a, b, *, bias=None, bias2=None, bias2_rows_per=0, residual=None, alpha=1.0, act=None, out_dtype=torch.float16,
         out=None, colstats_rows=None):
    """C[..., M, N] = act(alpha * A[..., M, K] @ B[..., N, K]^T + bias + bias2) + residual  (dwg_gemm_f16).
    a: [M,K] / [b1,M,K] / [b2,b1,M,K] fp16, last dim contiguous (any strides that are multiples of 8);
    b: same rank with N rows."""
    _chk_f16(a), _chk_f16(b)
    assert a.stride(-1) == 1 and b.stride(-1) == 1 and a.dim() == b.dim() and 2 <= a.dim() <= 4
    lead = (1,) * (4 - a.dim())
    a4 = a.as_strided(lead + tuple(a.shape), tuple(a.stride(0) * a.shape[0] for _ in lead) + tuple(a.stride()))
    b4 = b.as_strided(lead + tuple(b.shape), tuple(b.stride(0) * b.shape[0] for _ in lead) + tuple(b.stride()))
    nb2, nb1, M, K = a4.shape
    N = b4.shape[2]
    assert b4.shape[3] == K and b4.shape[0] == nb2 and b4.shape[1] == nb1
    No = N // 2 if act == 'geglu' else N                  # fused GEGLU halves the output width
    if out is None:
        out = torch.empty(nb2, nb1, M, No, device=a.device, dtype=out_dtype)
        c4 = out
        out = out.reshape(tuple(a.shape[:-1]) + (No,))
    else:
        c4 = out.as_strided((1,) * (4 - out.dim()) + tuple(out.shape), (0,) * (4 - out.dim()) + tuple(out.stride())) if out.dim() < 4 else out
        assert c4.stride(-1) == 1
    r4 = None
    if residual is not None:
        _chk_f16(residual)
        r4 = residual.as_strided((1,) * (4 - residual.dim()) + tuple(residual.shape), (0,) * (4 - residual.dim()) + tuple(residual.stride())) if residual.dim() < 4 else residual
        assert r4.stride(-1) == 1
    bias = None if bias is None else f32c(bias)
    bias2 = None if bias2 is None else f32c(bias2)
    if TUNE_RECORD is not None:
        TUNE_RECORD.append(('gemm', M, N, K, nb1, nb2, act, residual is not None, c4.dtype, bias is not None, bias2 is not None))
    cs = None
    if colstats_rows is not None and out_dtype == torch.float16 and act != 'geglu' and colstats_rows % 32 == 0 and (nb1 * nb2 * M) % colstats_rows == 0 \
            and N % 8 == 0:
        cs = _take_colstats(a.device, (nb1 * nb2 * M) // colstats_rows, N)
    nbytes = nb1 * nb2 * (2.0 * (M * K + N * K) + M * No * (c4.element_size() + (2 if r4 is not None else 0)))
    with _prof(2.0 * M * N * K * nb1 * nb2, f'gemm M{M} N{N} K{K} b{nb1 * nb2}', nbytes):
        ws = _gemm_workspace(a.device)
        check(lib().dwg_gemm_f16_ws(a4.data_ptr(), a4.stride(2), a4.stride(1), a4.stride(0),
                                  b4.data_ptr(), b4.stride(2), b4.stride(1), b4.stride(0),
                                  c4.data_ptr(), c4.stride(2), c4.stride(1), c4.stride(0), int(c4.dtype == torch.float16),
                                  M, N, K, nb1, nb2, ptr(bias), ptr(bias2), int(bias2_rows_per),
                                  None if r4 is None else r4.data_ptr(), 0 if r4 is None else r4.stride(2),
                                  0 if r4 is None else r4.stride(1), 0 if r4 is None else r4.stride(0),
                                  float(alpha), ACT[act], ws.data_ptr(), ws.numel(), ptr(cs), int(colstats_rows or 0), stream()), 'dwg_gemm_f16_ws')
    if cs is not None:
        out._cs = cs
    return out


def take_rowstats(device, rows):
    """Zero-initialised i64 [rows, 2] for an epilogue's LayerNorm row statistics (arena slice inside a step)."""
    rs = STATS_ARENA.take(device, rows * 2)
    if rs is None:
        rs = torch.zeros(rows * 2, device=device, dtype=torch.int64)
    return rs.view(rows, 2)


def gemm_ln(a, b, *, bias=None, residual=None, act=None, ln=None, ln_cols=None, rowstats=False, colstats_rows=None, out=None):
    """fp16 GEMM C = act(A B^T + bias) + residual ([M,K] x [N,K], or one batch dimension [nb,M,K] x [nb,N,K]; a batch stride of 0
    broadcasts A) with a LayerNorm folded into the epilogue (dwg_gemm_f16_ln) and / or LayerNorm statistics of the output rows.
      ln = (stats i64 [nb*M,2], c1 [N], dim, eps):                   the A rows are un-normalised activations, b = W * gamma, bias = W beta (+ bias)
      ln_cols = (stats i64 [nb*N,2], c1 [M], rowbias [M], dim, eps):  the B rows are the activations (V^T = Wv' x^T), a = Wv * gamma
      rowstats=True: out._rs = i64 [nb*M,2] (sum, sum of squares of the stored fp16 values, 2^-20 fixed point)."""
    _chk_f16(a), _chk_f16(b)
    assert a.dim() == b.dim() and a.dim() in (2, 3) and a.stride(-1) == 1 and b.stride(-1) == 1 and not (ln and ln_cols)
    if a.dim() == 2:
        a3, b3 = a.unsqueeze(0), b.unsqueeze(0)
    else:
        a3, b3 = a, b
    nb, M, K = a3.shape
    N = b3.shape[1]
    assert b3.shape[0] == nb and b3.shape[2] == K
    No = N // 2 if act == 'geglu' else N
    if out is None:
        out = torch.empty((nb, M, No) if a.dim() == 3 else (M, No), device=a.device, dtype=torch.float16)
    o3 = out if out.dim() == 3 else out.unsqueeze(0)
    assert o3.stride(-1) == 1
    r3 = None
    if residual is not None:
        _chk_f16(residual)
        r3 = residual if residual.dim() == 3 else residual.unsqueeze(0)
        assert r3.stride(-1) == 1
    mode, st, c1, rb, dim, eps = 0, None, None, None, 0, 0.0
    if ln is not None:
        mode, (st, c1, dim, eps) = 1, ln
        assert st.shape == (nb * M, 2) and c1.numel() == N
    if ln_cols is not None:
        mode, (st, c1, rb, dim, eps) = 2, ln_cols
        assert st.shape == (nb * N, 2) and c1.numel() == M
    rs = take_rowstats(a.device, nb * M) if rowstats else None
    cs = None
    if colstats_rows is not None and act != 'geglu' and colstats_rows % 32 == 0 and (nb * M) % colstats_rows == 0 and N % 8 == 0:
        cs = _take_colstats(a.device, (nb * M) // colstats_rows, N)
    bias = None if bias is None else f32c(bias)
    bs = lambda t: t.stride(0) if nb > 1 else 0
    with _prof(2.0 * M * N * K * nb, f'gemm M{M} N{N} K{K} b{nb}', 2.0 * nb * (M * K + N * K + M * No)):
        ws = _gemm_workspace(a.device)
        check(lib().dwg_gemm_f16_ln(a3.data_ptr(), a3.stride(1), bs(a3), b3.data_ptr(), b3.stride(1), bs(b3), o3.data_ptr(), o3.stride(1), bs(o3),
                                    M, N, K, nb, ptr(bias), None if r3 is None else r3.data_ptr(), 0 if r3 is None else r3.stride(1),
                                    0 if r3 is None else bs(r3), ACT[act], mode, ptr(st), ptr(c1), ptr(rb), int(dim), float(eps), ptr(rs),
                                    ws.data_ptr(), ws.numel(), ptr(cs), int(colstats_rows or 0) if cs is not None else 0, stream()), 'dwg_gemm_f16_ln')
    if rs is not None:
        out._rs = rs
    if cs is not None:
        out._cs = cs
    return out


def _conv_stats_ok(Nimg, Ho, Wo, Cout):
    """Mirror of the tile geometry of dwg_conv2d_nhwc_f16 (csrc/gemm_tcgen05.cu): column statistics need exactly tiled images
    whose share of a 128-row tile is a multiple of 32 pixels (true for every layer of the SD UNets / VAE)."""
    BW = 1
    while BW * 2 <= Wo and BW * 2 <= 128:
        BW *= 2
    BH = 1
    while BH * 2 <= Ho and BW * BH * 2 <= 128:
        BH *= 2
    BNI = 128 // (BW * BH)
    if BNI > Nimg:
        BNI = 1 << (Nimg.bit_length() - 1)
    return Wo % BW == 0 and Ho % BH == 0 and Nimg % BNI == 0 and (BW * BH) % 32 == 0 and Cout % 8 == 0


def conv2d_nhwc(x, w, *, bias=None, bias2=None, residual=None, stride=1, padding=1, out_hw=None, act=None,
                out_dtype=torch.float16, stats=False):
    """NHWC implicit-GEMM convolution (dwg_conv2d_nhwc_f16).  x [N,H,W,Cin], w [Cout,k,k,Cin] fp16.
    padding: int (symmetric) or (top, left) with out_hw=(Ho, Wo) for asymmetric cases."""
    _chk_f16(x), _chk_f16(w)
    assert x.is_contiguous() and w.is_contiguous()
    Nimg, H, W, Cin = x.shape
    Cout, k, k2, Cin2 = w.shape
    assert k == k2 and Cin2 == Cin
    ph, pw = (padding, padding) if isinstance(padding, int) else padding
    if out_hw is None:
        Ho, Wo = (H + 2 * ph - k) // stride + 1, (W + 2 * pw - k) // stride + 1
    else:
        Ho, Wo = out_hw
    y = torch.empty(Nimg, Ho, Wo, Cout, device=x.device, dtype=out_dtype)
    if residual is not None:
        _chk_f16(residual)
        assert residual.shape == y.shape and residual.is_contiguous()
    bias = None if bias is None else f32c(bias)
    bias2 = None if bias2 is None else f32c(bias2)
    if TUNE_RECORD is not None:
        TUNE_RECORD.append(('conv', Nimg, H, W, Cin, Cout, k, stride, ph, pw, Ho, Wo, residual is not None, out_dtype, bias2 is not None))
    cs = _take_colstats(x.device, Nimg, Cout) if (stats and out_dtype == torch.float16 and _conv_stats_ok(Nimg, Ho, Wo, Cout)) else None
    nbytes = 2.0 * (x.numel() + w.numel()) + y.numel() * (y.element_size() + (2 if residual is not None else 0))
    with _prof(2.0 * Nimg * Ho * Wo * Cout * Cin * k * k, f'conv{k}x{k}s{stride} {Nimg}x{Ho}x{Wo} {Cin}->{Cout}', nbytes):
        ws = _gemm_workspace(x.device)
        check(lib().dwg_conv2d_nhwc_f16_ws(ptr(x), ptr(w), ptr(y), int(out_dtype == torch.float16), Nimg, H, W, Cin, Cout, k,
                                            stride, ph, pw, Ho, Wo, ptr(bias), ptr(bias2), ptr(residual), ACT[act], ws.data_ptr(), ws.numel(),
                                            ptr(cs), stream()), 'dwg_conv2d_nhwc_f16_ws')
    if cs is not None:
        y._cs = cs
    return y


# ------------------------------------------------------------------------------ norm / activation kernels (fp16 NHWC)
class StatsArena:
    """One zero-initialised int64 buffer per device that serves the statistics workspace of EVERY GroupNorm (forward and
    backward) of a step: ONE memset per step (reset(), called by the guidance at the top of __call__) instead of one memset
    node in front of each of the ~130 GroupNorm launches.  A slice is handed out at most once between two resets; when the
    arena is exhausted or disabled the caller falls back to a private buffer that the C entry point zeroes itself."""

    def __init__(self, entries=1 << 22):
        self.entries, self.buf, self.cur, self.enabled = entries, {}, {}, True

    def reset(self, device):
        dev = torch.device(device)
        key = dev.index if dev.index is not None else torch.cuda.current_device()
        if key not in self.buf:
            self.buf[key] = torch.zeros(self.entries, device=dev, dtype=torch.int64)
        else:
            self.buf[key].zero_()
        self.cur[key] = 0

    def take(self, device, n):
        if not self.enabled:
            return None
        key = device.index if device.index is not None else torch.cuda.current_device()
        cur = self.cur.get(key)
        if cur is None or cur + n > self.entries:
            return None
        self.cur[key] = cur + n
        return self.buf[key][cur:cur + n]


STATS_ARENA = StatsArena()


def group_norm(x, gamma, beta, groups=32, eps=1e-5, silu=False, return_stats=False, colstats=None):
    """x [N, ..., C] fp16 channels-last -> [SiLU](GroupNorm(x)).  ``colstats`` i64 [N,C,2]: the per-column statistics the
    producing GEMM / conv epilogue accumulated (ops.gemm(colstats_rows=...) / ops.conv2d_nhwc(stats=True)) -- then ONE kernel
    (dwg_groupnorm_apply_cs); otherwise statistics pass + apply pass (dwg_groupnorm_fwd)."""
    _chk_f16(x)
    assert x.is_contiguous()
    N, C = x.shape[0], x.shape[-1]
    HW = x.numel() // (N * C)
    y = torch.empty_like(x)
    L = lib()
    if colstats is not None and groups <= 32:
        assert colstats.shape == (4, N, C, 2) and colstats.dtype == torch.int64 and colstats.is_contiguous()
        stats = None
        if return_stats:
            stats = STATS_ARENA.take(x.device, N * groups * 2)
            stats = (stats if stats is not None else torch.empty(N * groups * 2, device=x.device, dtype=torch.int64)).view(N, groups, 2)
        check(L.dwg_groupnorm_apply_cs(ptr(x), ptr(gamma), ptr(beta), ptr(y), ptr(colstats), ptr(stats), N, HW, C, groups, float(eps), int(silu),
                                       stream()), 'dwg_groupnorm_apply_cs')
        return (y, stats) if return_stats else y
    stats = STATS_ARENA.take(x.device, N * groups * 2)        # fixed-point (sum, sumsq), include/dwg.h
    flags = int(silu) | (2 if stats is not None else 0)
    if stats is None:
        stats = torch.empty(N * groups * 2, device=x.device, dtype=torch.int64)
    stats = stats.view(N, groups, 2)
    check(L.dwg_groupnorm_fwd(ptr(x), ptr(gamma), ptr(beta), ptr(y), ptr(stats), N, HW, C, groups, float(eps), flags,
                              stream()), 'dwg_groupnorm_fwd')
    object.__setattr__(L, 'launches', L.launches + L.dwg_groupnorm_last_launches() - 2)      # the proxy counted 2
    return (y, stats) if return_stats else y


def group_norm_cat(x1, x2, gamma, beta, groups=32, eps=1e-5, silu=False):
    """[SiLU](GroupNorm(cat([x1, x2], channel))) without the concatenation (dwg_groupnorm_apply_cs2): both inputs carry the column
    statistics of their producing epilogues (x._cs).  Returns the normalised [N, ..., C1 + C2] tensor."""
    _chk_f16(x1), _chk_f16(x2)
    assert x1.is_contiguous() and x2.is_contiguous() and x1.shape[:-1] == x2.shape[:-1] and groups <= 32
    N, C1, C2 = x1.shape[0], x1.shape[-1], x2.shape[-1]
    C = C1 + C2
    HW = x1.numel() // (N * C1)
    cs1, cs2 = x1._cs, x2._cs
    assert cs1.shape == (4, N, C1, 2) and cs2.shape == (4, N, C2, 2) and cs1.is_contiguous() and cs2.is_contiguous()
    y = torch.empty(x1.shape[:-1] + (C,), device=x1.device, dtype=torch.float16)
    check(lib().dwg_groupnorm_apply_cs2(ptr(x1), ptr(x2), C1, ptr(gamma), ptr(beta), ptr(y), ptr(cs1), ptr(cs2), None, N, HW, C, groups,
                                        float(eps), int(silu), stream()), 'dwg_groupnorm_apply_cs2')
    return y


def nchw_to_nhwc_f16(x, scale=1.0, shift=0.0):
    """[N,C,H,W] fp32 -> [N,H,W,C8] fp16 (C padded to a multiple of 8 with zeros) = scale * x + shift, one kernel."""
    x = f32c(x)
    N, C, H, W = x.shape
    Cp = (C + 7) // 8 * 8
    y = torch.empty(N, H, W, Cp, device=x.device, dtype=torch.float16)
    check(lib().dwg_nchw_f32_to_nhwc_f16(ptr(x), ptr(y), N, C, H * W, Cp, float(scale), float(shift), stream()), 'dwg_nchw_f32_to_nhwc_f16')
    return y


def nhwc_f16_to_nchw(x, C, scale=1.0):
    """[N,H,W,Cp] fp16 (Cp >= C, padded or not) -> [N,C,H,W] fp32 = scale * x[..., :C] (C <= 8), one kernel."""
    _chk_f16(x)
    assert x.is_contiguous() and x.dim() == 4 and C <= 8
    N, H, W, Cp = x.shape
    y = torch.empty(N, C, H, W, device=x.device, dtype=torch.float32)
    check(lib().dwg_nhwc_f16_to_nchw_f32(ptr(x), ptr(y), N, C, H * W, Cp, float(scale), stream()), 'dwg_nhwc_f16_to_nchw_f32')
    return y


def cat_channels(a, b):
    """torch.cat([a, b], dim=-1) of two channels-last activations, carrying their column statistics along (the statistics of
    a channel concatenation are the concatenation of the statistics)."""
    y = torch.cat([a, b], dim=-1)
    ca, cb = getattr(a, '_cs', None), getattr(b, '_cs', None)
    if ca is not None and cb is not None:
        y._cs = torch.cat([ca, cb], dim=2)
    return y


def group_norm_bwd(x, dy, stats, gamma, beta, groups=32, eps=1e-5, silu=False, dx_add=None):
    _chk_f16(x), _chk_f16(dy)
    assert x.is_contiguous() and dy.is_contiguous()
    N, C = x.shape[0], x.shape[-1]
    HW = x.numel() // (N * C)
    dx = torch.empty_like(x)
    bstats = STATS_ARENA.take(x.device, N * groups * 2)
    flags = int(silu) | (2 if bstats is not None else 0)
    if bstats is None:
        bstats = torch.empty(N * groups * 2, device=x.device, dtype=torch.int64)
    check(lib().dwg_groupnorm_bwd(ptr(x), ptr(dy), ptr(stats), ptr(gamma), ptr(beta), ptr(dx_add), ptr(dx), ptr(bstats), N, HW, C,
                                  groups, float(eps), flags, stream()), 'dwg_groupnorm_bwd')
    return dx


def layer_norm(x, gamma, beta, eps=1e-5):
    _chk_f16(x)
    assert x.is_contiguous()
    C = x.shape[-1]
    y = torch.empty_like(x)
    check(lib().dwg_layernorm_fwd(ptr(x), ptr(gamma), ptr(beta), ptr(y), x.numel() // C, C, float(eps), stream()), 'dwg_layernorm_fwd')
    return y


def softmax_rows_(s, cols):
    """In-place softmax over the last dim of fp16 scores [..., cols_pad]; columns >= cols become 0."""
    _chk_f16(s)
    assert s.is_contiguous()
    cp = s.shape[-1]
    check(lib().dwg_softmax_rows(ptr(s), s.numel() // cp, int(cols), cp, stream()), 'dwg_softmax_rows')
    return s


def softmax_rows_bwd_(p, dp):
    _chk_f16(p), _chk_f16(dp)
    assert p.is_contiguous() and dp.is_contiguous()
    cp = p.shape[-1]
    check(lib().dwg_softmax_rows_bwd(ptr(p), ptr(dp), p.numel() // cp, cp, stream()), 'dwg_softmax_rows_bwd')
    return dp


def geglu(x):
    _chk_f16(x)
    assert x.is_contiguous()
    inner = x.shape[-1] // 2
    y = torch.empty(x.shape[:-1] + (inner,), device=x.device, dtype=torch.float16)
    check(lib().dwg_geglu(ptr(x), ptr(y), x.numel() // x.shape[-1], inner, stream()), 'dwg_geglu')
    return y


def silu(x):
    _chk_f16(x)
    y = torch.empty_like(x)
    check(lib().dwg_eltwise_f16(ptr(x.contiguous()), None, ptr(y), x.numel(), 0, stream()), 'dwg_eltwise_f16')
    return y


def add(x, a):
    _chk_f16(x), _chk_f16(a)
    y = torch.empty_like(x)
    check(lib().dwg_eltwise_f16(ptr(x.contiguous()), ptr(a.contiguous()), ptr(y), x.numel(), 1, stream()), 'dwg_eltwise_f16')
    return y


def sds_grad(eps_uncond, eps_cond, noise, guidance_scale, weight=1.0):
    """(grad, noise_pred) of basic.py:595-603,642 in fp32."""
    eu, ec, nz = f32c(eps_uncond), f32c(eps_cond), f32c(noise)
    grad, npred = torch.empty_like(nz), torch.empty_like(nz)
    check(lib().dwg_sds_grad(ptr(eu), ptr(ec), ptr(nz), ptr(grad), ptr(npred), float(guidance_scale), float(weight), nz.numel(),
                             stream()), 'dwg_sds_grad')
    return grad, npred


def attention(q, k, vt, heads, Tk, scale=None):
    """Fused attention forward (dwg_attention_fwd).  q [B,T,C], k [B,Tk,C] fp16 (last dim contiguous),
    vt [B,C,Tkp] = V transposed.  Returns [B,T,C] fp16."""
    _chk_f16(q), _chk_f16(k), _chk_f16(vt)
    B, T, C = q.shape
    hd = C // heads
    assert q.stride(2) == 1 and k.stride(2) == 1 and vt.stride(2) == 1 and vt.stride(1) == vt.shape[2]
    assert q.stride(0) == T * q.stride(1) and k.stride(0) == Tk * k.stride(1)
    out = torch.empty(B, T, C, device=q.device, dtype=torch.float16)
    scale = hd ** -0.5 if scale is None else scale
    fl = 4.0 * B * heads * T * Tk * hd
    with _prof(fl, f'attention B{B} h{heads} T{T} Tk{Tk} d{hd}'):
        check(lib().dwg_attention_fwd(q.data_ptr(), q.stride(1), k.data_ptr(), k.stride(1), vt.data_ptr(), vt.shape[2], vt.stride(0), ptr(out), B, heads, T, Tk,
                                      hd, float(scale), stream()), 'dwg_attention_fwd')
    return out
