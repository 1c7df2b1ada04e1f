/* dwg.h -- C ABI of libdwg_sm100.so: the B200-native DreamWaltz-G SDS hot path.
 *
 * Conventions (all entry points):
 *   - plain C types only: device pointers, int/int64_t sizes, float scalars; `stream` is a
 *     cudaStream_t passed as void* (NULL = legacy default stream);
 *   - the caller owns every buffer (inputs, outputs, workspaces): every entry point on the product path takes its
 *     scratch from the caller (rasteriser workspaces, dwg_gemm_f16_ws / dwg_conv2d_nhwc_f16_ws + dwg_gemm_workspace_bytes,
 *     GroupNorm statistics, optimiser state).  The forms WITHOUT the _ws suffix are conveniences that lazily allocate one
 *     per-process split-K scratch (two lanes, dwg_gemm_set_lane) and are therefore not re-entrant across host threads;
 *   - all tensors are dense row-major fp32 unless stated; "i32"/"u32"/"u64" = integer types;
 *   - return value 0 = success, negative = error (dwg_last_error() gives the message);
 *     kernels are launched asynchronously on `stream`, errors of the launch itself are
 *     reported, execution errors surface at the caller's next synchronisation;
 *   - re-entrant per stream; besides the last-error string the only process-global state is (a) the convenience scratch above
 *     and (b) the tuning / introspection knobs (dwg_gemm_tune*, dwg_gemm_last_*, dwg_groupnorm_set_fused, dwg_nn_set_carveout),
 *     which are debugging aids of tools/ and never change results.
 *
 * Each function cites the reference interface it replaces (paths into the reference repo).
 */
#ifndef DWG_H_
#define DWG_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DWG_OK 0
#define DWG_ERR_INVALID -1      /* bad argument (shape / null / unsupported option)  */
#define DWG_ERR_CUDA -2         /* CUDA runtime error at launch                       */
#define DWG_ERR_CAPACITY -3     /* caller-provided workspace too small                */

const char* dwg_last_error(void);
int dwg_version(void);
/* device sanity: returns compute capability major*10+minor of the current device, <0 on error */
int dwg_device_cc(void);

/* ------------------------------------------------------------------------------------------
 * R2/R3  Linear-blend skinning of Gaussians.
 * Replaces RigidTransform.transform_points(weights=W) + transform_quaternions(weights=W,
 * flip_rotation_axis=True)  (core/human/inverse_lbs.py:190-242) as called from
 * DreamWaltzG.lbs_transform (core/system/avatar.py:1446-1460).
 *   W  [N,J]   blend weights (already row-normalised, avatar.py:914-917)
 *   A  [J,4,4] joint SE3 (J_pose_rigid composed with G_transl_offset); rows 0..2 are used
 *   x  [N,3]   positions           q  [N,4] real-first quaternions or NULL (positions only)
 *   x_out [N,3], q_out [N,4] (NULL iff q NULL)
 * M_n = sum_j W[n,j] A[j,:3,:];  x' = M[:,:3] x + M[:,3];
 * q' = mat2quat(F (M[:,:3] (F quat2mat(q))))  with F = diag(1,-1,-1)   (pytorch3d 0.7.5 maths).
 */
int dwg_lbs_skin_fwd(const float* W, const float* A, const float* x, const float* q,
                     float* x_out, float* q_out, int64_t N, int J, void* stream);
/* Backward.  g_x_out [N,3], g_q_out [N,4] or NULL  ->  g_x [N,3], g_q [N,4] (NULL iff q NULL),
 * optional g_W [N,J] (written) and g_A [J,4,4] (ACCUMULATED with atomics; caller zeroes). */
int dwg_lbs_skin_bwd(const float* W, const float* A, const float* x, const float* q,
                     const float* g_x_out, const float* g_q_out,
                     float* g_x, float* g_q, float* g_W, float* g_A,
                     int64_t N, int J, void* stream);

/* ------------------------------------------------------------------------------------------
 * R10  Spherical-harmonic colour.  Replaces eval_sh + get_colors + compute_colors
 * (core/gaussian/spherical_harmonics.py:117-172, gaussian_utils.py:12-17,
 * gaussian_renderer.py:72-105):  dirs = normalize(pos - campos), rgb = max(sum_k Y_k sh_k + 0.5, 0).
 *   sh [N, sh_stride, 3] (first (deg+1)^2 coefficients used), pos [N,3], campos [3] (device)
 *   rgb [N,3]; clamped [N] u8 bit c set when channel c was clamped (needed by the backward).
 */
int dwg_sh_eval_fwd(const float* sh, int sh_stride, int deg, const float* pos, const float* campos,
                    float* rgb, uint8_t* clamped, int64_t N, void* stream);
/* g_rgb [N,3] -> g_sh [N, sh_stride, 3] (first (deg+1)^2 rows written, rest zeroed), g_pos [N,3] or NULL */
int dwg_sh_eval_bwd(const float* sh, int sh_stride, int deg, const float* pos, const float* campos,
                    const uint8_t* clamped, const float* g_rgb, float* g_sh, float* g_pos,
                    int64_t N, void* stream);

/* ------------------------------------------------------------------------------------------
 * R6  Multi-resolution grid encoder.  Replaces _gridencoder.grid_encode_forward/backward
 * (core/nerf/gridencoder/src/gridencoder.h:12-15, gridencoder.cu:87-500) incl. the
 * (x+bound)/(2*bound) mapping of GridEncoder.forward (grid.py:153).  D = 3, C = 2.
 *   x [B,3] world positions (bound > 0) or inputs already mapped to [0,1] (bound <= 0); table [rows,2]; offsets i32 [L+1]; level_scale f32 [L] and
 *   level_res u32 [L] = exp2f(l*S)*H-1 and ceil(scale)+1 (gridencoder.cu:138-139) as written by dwg_grid_level_table;
 *   out: element (b, l, c) at out[b*out_stride_b + l*out_stride_l + c]  (so both the final
 *   [B, L*C] layout and the reference's [L,B,C] staging layout are expressible);
 *   dy_dx [B, L*3*C] or NULL (derivative w.r.t. the [0,1]-mapped input, as the reference).
 *   gridtype 0 hash / 1 tiled; interp 0 linear / 1 smoothstep.
 */
/* Per-level constants, evaluated on the device exactly as the reference kernel does per thread
 * (gridencoder.cu:138-139): level_scale[l] = exp2f(l * S) * H - 1.0f, level_res[l] = (uint32)ceil(scale) + 1.
 * S = log2(per_level_scale) as float32 (grid.py:40), H = base resolution.  Outputs are DEVICE arrays of L entries. */
int dwg_grid_level_table(float S, int H, int L, float* level_scale, uint32_t* level_res, void* stream);
int dwg_grid_encode_fwd(const float* x, float bound, const float* table, const int32_t* offsets,
                        const float* level_scale, const uint32_t* level_res,
                        float* out, int64_t out_stride_b, int64_t out_stride_l, float* dy_dx,
                        int64_t B, int L, int gridtype, int align_corners, int interp, void* stream);
/* grad: element (b,l,c) at grad[b*g_stride_b + l*g_stride_l + c].  g_table [rows,2] is
 * ACCUMULATED (atomics; caller zeroes, grid.py:81).  g_x [B,3] or NULL: gradient w.r.t. the
 * WORLD position (includes the 1/(2*bound) factor); recomputed from the table, no dy_dx buffer. */
int dwg_grid_encode_bwd(const float* grad, int64_t g_stride_b, int64_t g_stride_l,
                        const float* x, float bound, const float* table, const int32_t* offsets,
                        const float* level_scale, const uint32_t* level_res,
                        float* g_table, float* g_x,
                        int64_t B, int L, int gridtype, int align_corners, int interp, void* stream);

/* ------------------------------------------------------------------------------------------
 * R11-R13  Differentiable tile rasteriser.  Replaces diff_gaussian_rasterization._C
 * rasterize_gaussians / rasterize_gaussians_backward (third party, ashawkey fork; call site
 * core/gaussian/gaussian_renderer.py:186-195).  16x16 tiles.
 */
typedef struct {
    int32_t image_height, image_width;
    float tanfovx, tanfovy;
    float viewmatrix[16];     /* GaussianRasterizationSettings.viewmatrix, row-major flat */
    float projmatrix[16];     /* GaussianRasterizationSettings.projmatrix, row-major flat */
    float bg[3];
    float scale_modifier;
} DwgRasterCamera;

/* Workspace sizing.  P_cap = capacity (in (tile,Gaussian) instances) the caller grants the
 * binning buffers; the true count P is data dependent and stays on the device (no host sync).
 * If P > P_cap the forward sets status word bit 0 and renders nothing beyond capacity. */
int64_t dwg_raster_geom_bytes(int64_t N);                       /* per-Gaussian state     */
int64_t dwg_raster_bin_bytes(int64_t P_cap, int H, int W);      /* instances + tile table */
int64_t dwg_raster_img_bytes(int H, int W);                     /* final_T, n_contrib     */

/* Forward.  colors_precomp [N,3]; opacities [N]; scales [N,3]; rotations [N,4] (w,x,y,z, used
 * unnormalised); outputs out_color [3,H,W], out_depth [H,W], out_alpha [H,W], radii i32 [N].
 * geom/bin/img are caller workspaces of the sizes above (kept for the backward).
 * status: i32[4] device words {overflow flag, P (num_rendered), max tile load, reserved}.
 * cam_dev: NULL, or a DEVICE copy of the camera struct that overrides the matrices / tanfov / bg of
 * `cam` (same image size): lets a captured CUDA graph be replayed with a new camera.
 * bg_image: NULL, or a per-pixel background [3,H,W] composited in the blend epilogue exactly as
 * Scene.forward does (core/system/scene.py:153-166): out_color = image_fg + bg_image * (1 - out_alpha), with
 * image_fg = blended colour + final_T * cam.bg written to out_color_fg [3,H,W] when that is not NULL. */
int dwg_raster_forward(const DwgRasterCamera* cam, int64_t N,
                       const float* means3D, const float* colors_precomp, const float* opacities,
                       const float* scales, const float* rotations,
                       float* out_color, float* out_depth, float* out_alpha, int32_t* radii,
                       void* geom, void* bin, int64_t P_cap, void* img, int32_t* status,
                       const void* cam_dev, const float* bg_image, float* out_color_fg, void* stream);
/* Backward.  dL_dcolor [3,H,W], dL_ddepth [H,W] or NULL, dL_dalpha [H,W] or NULL ->
 * g_means3D [N,3], g_means2D [N,3] (z = 0), g_colors [N,3], g_opacities [N], g_scales [N,3],
 * g_rotations [N,4].  All outputs are fully written (no pre-zeroing needed); `scratch` must
 * hold dwg_raster_bwd_scratch_bytes(N) bytes.  dL_dcolor is the gradient of out_color (the composite when bg_image was
 * given: pass the same bg_image and the forward's out_alpha; g_bg_image [3,H,W] or NULL receives dL/d bg_image). */
int64_t dwg_raster_bwd_scratch_bytes(int64_t N);
int dwg_raster_backward(const DwgRasterCamera* cam, int64_t N,
                        const float* means3D, const float* colors_precomp, const float* opacities,
                        const float* scales, const float* rotations,
                        const void* geom, const void* bin, int64_t P_cap, const void* img,
                        const float* dL_dcolor, const float* dL_ddepth, const float* dL_dalpha,
                        float* g_means3D, float* g_means2D, float* g_colors, float* g_opacities,
                        float* g_scales, float* g_rotations, void* scratch, const void* cam_dev,
                        const float* bg_image, const float* out_alpha, float* g_bg_image, void* stream);
/* Debug / parity views into the workspaces (device pointers into geom / bin / img):
 * which = 0 xy f32[N,2] | 1 depth f32[N] | 2 cov3D f32[N,6] | 3 conic_opacity f32[N,4]
 *       | 4 rect i32[N,4] | 5 tiles_touched u32[N] | 6 tile ranges u32[tiles,2]
 *       | 7 sorted keys u64[P_cap] | 8 sorted values u32[P_cap] | 9 final_T f32[H,W]
 *       | 10 n_contrib u32[H,W] */
void* dwg_raster_view(int which, void* geom, void* bin, void* img, int64_t N, int64_t P_cap, int H, int W);

/* ------------------------------------------------------------------------------------------
 * R14/R15  Dense layers of the UNet / ControlNet / VAE encoder on the 5th-gen tensor cores
 * (tcgen05.mma, accumulators in TMEM, operands staged by TMA).  Replaces the cuBLAS / cuDNN
 * calls diffusers makes for nn.Linear / nn.Conv2d / attention matmuls inside
 * UNet2DConditionModel, ControlNetModel and AutoencoderKL (third party; call sites
 * core/guidance/controlnet.py:98-114, core/guidance/vae.py:34-40).
 *
 * dwg_gemm_f16:  C[b2][b1][M,N] = act(alpha * A[b2][b1][M,K] . B[b2][b1][N,K]^T + bias[N]
 *                                      + bias2[row / bias2_rows_per][N]) + residual
 *   A, B bf16 with K contiguous; strides in ELEMENTS and multiples of 8; C bf16 (out_f16=1) or
 *   fp32; residual bf16 with its own strides; act 0 none / 1 SiLU / 2 GELU(erf).
 */
/* bytes of split-K scratch one workspace must hold (counters + fp32 partial-tile slices).  A workspace must be zero-filled
 * ONCE by the caller before its first use (the kernels leave the counters zero); launches that may run concurrently
 * (two streams) need different workspaces. */
int64_t dwg_gemm_workspace_bytes(void);
/* Batches: nb1 x nb2 problems with element strides a_b1 / a_b2 (b_*, c_*, r_* likewise); a_b1 == 0 with nb1 > 1 broadcasts ONE A
 * matrix over the first batch dimension (a weight matrix as the A operand of every batch entry).
 * colstats (NULL = off): i64 [4, groups, N, 2] (4 slots that spread the atomics; the consumer adds them), PRE-ZEROED by the caller; the epilogue adds, per output column, the sum and
 * the sum of squares of the fp16 values it stores (2^-20 fixed point, integer atomics => deterministic), grouped by image
 * (conv: the image index; GEMM: global row / colstats_rows, a multiple of 32).  These are the GroupNorm statistics of the
 * layer that consumes the output: dwg_groupnorm_apply_cs folds them per group, so no statistics pass reads the tensor. */
int dwg_gemm_f16_ws(const void* A, int64_t lda, int64_t a_b1, int64_t a_b2,
                    const void* B, int64_t ldb, int64_t b_b1, int64_t b_b2,
                    void* C, int64_t ldc, int64_t c_b1, int64_t c_b2, int out_f16,
                    int M, int N, int K, int nb1, int nb2,
                    const float* bias, const float* bias2, int bias2_rows_per,
                    const void* residual, int64_t ldr, int64_t r_b1, int64_t r_b2,
                    float alpha, int act, void* workspace, int64_t workspace_bytes, void* colstats, int colstats_rows, void* stream);
int dwg_gemm_f16(const void* A, int64_t lda, int64_t a_b1, int64_t a_b2,
                  const void* B, int64_t ldb, int64_t b_b1, int64_t b_b2,
                  void* C, int64_t ldc, int64_t c_b1, int64_t c_b2, int out_f16,
                  int M, int N, int K, int nb1, int nb2,
                  const float* bias, const float* bias2, int bias2_rows_per,
                  const void* residual, int64_t ldr, int64_t r_b1, int64_t r_b2,
                  float alpha, int act, void* stream);
/* dwg_avatar_mlp_fwd / _bwd: the two per-Gaussian MLPs of DreamWaltzG.animate with their
 * activations, fused.  fp32 results on the tensor cores: every operand is split into an fp16 (hi, lo) pair and a layer is
 * the three products hi*hi + hi*lo + lo*hi accumulated in fp32 (tcgen05, csrc/avatar_mlp.cu; dwg_avatar_mlp_set_tc(0)
 * selects the plain fp32 SIMT kernels).  Replaces reference core/nerf/nerf_model.py:12-33 (MLP 32-64-64-4,
 * ReLU) as used by static_mlp_forward core/system/avatar.py:1283-1290, DeformNetwork.forward
 * core/deformation/deform_model.py:102-143 (D=4, W=64, leaky_relu; heads warp / scaling) as used by
 * dynamic_mlp_forward avatar.py:1292-1294, and non_rigid_transform avatar.py:1464-1498 with the
 * shipped flags (pos = positions + init_offset*warp; scales = min(exp(scaling)*init_scale, max_scale)).
 *   enc [N,32] grid features; the first Nu Gaussians are unconstrained (both nets), the remaining
 *   N-Nu are mesh-bound (colour net only, opacity = 1).
 *   params: flat fp32 vector of dwg_avatar_mlp_param_count() floats, concatenation of
 *     net.0.{weight[64,32],bias} net.1.{weight[64,64],bias} net.2.{weight[4,64],bias}
 *     layers.0.weight[:, :32] layers.0.bias layers.{1,2,3}.{weight,bias}
 *     gaussian_warp.{weight[3,64],bias} gaussian_scaling.{weight[3,64],bias};
 *   w_pose = layers.0.weight[:, 32:95] [64,63], body_pose [63].
 *   acts_s [2][64][Np], acts_d [4][64][Nup] (Np/Nup = N/Nu rounded up to 128): hidden activations kept
 *   for the backward (null = inference); the tensor-core kernels lay them out tile-blocked, [layer][tile][64][128 points].
 *   Backward overwrites g_enc [N,32], g_params (same layout as params) and g_w_pose [64,63]; null output-gradient
 *   pointers mean zero; scratch = dwg_avatar_mlp_bwd_scratch_bytes(N) bytes. */
int64_t dwg_avatar_mlp_param_count(void);
int64_t dwg_avatar_mlp_bwd_scratch_bytes(int64_t N);   /* scratch of dwg_avatar_mlp_bwd for N Gaussians (supersedes dwg_avatar_mlp_scratch_bytes) */
int dwg_avatar_mlp_set_tc(int on);   /* debug / A-B switch: 1 (default) = tcgen05 kernels (fp16 hi+lo split, fp32 accuracy), 0 = fp32 SIMT kernels */
int64_t dwg_avatar_mlp_scratch_bytes(void);
int dwg_avatar_mlp_fwd(const float* enc, const float* positions, const float* params, const float* w_pose, const float* body_pose,
                       float* colors, float* opac, float* pos_out, float* scales, float* acts_s, float* acts_d,
                       int64_t N, int64_t Nu, float init_offset, float init_scale, float max_scale, void* stream);
int dwg_avatar_mlp_bwd(const float* enc, const float* params, const float* body_pose,
                       const float* colors, const float* opac, const float* scales, const float* acts_s, const float* acts_d,
                       const float* g_colors, const float* g_opac, const float* g_pos, const float* g_scales,
                       float* g_enc, float* g_params, float* g_w_pose, void* scratch,
                       int64_t N, int64_t Nu, float init_offset, float max_scale, void* stream);

/* Tuning / introspection of the tcgen05 GEMM tile planner (no reference counterpart; used by
 * tools/gemm_sweep.py): force BN (tile width) and the split-K factor of subsequent launches
 * (0, 0 = automatic); read the (BN, ksplit, stages) the last launch ran with. */
int dwg_gemm_tune(int force_bn, int force_ks);
int dwg_gemm_last_plan(int* out3);
/* CTA-pair (tcgen05.mma.cta_group::2, cluster of 2) mode of the following launches: -1 automatic (planner / tuned
 * table), 0 never, 1 whenever legal (even number of 128-row tiles); dwg_gemm_last_pair() -> what the last launch used. */
int dwg_gemm_tune_pair(int mode);
/* Halo mode of 3x3 / stride-1 / "same" convolutions (one (16+2)x(8+2) activation halo per 64-channel slice feeds all nine
 * taps): -1 automatic, 0 never, 1 whenever legal.  base_offset_mode: 0 (default, correct on sm_100a) leaves the
 * matrix-descriptor base-offset field of the shifted windows zero; 1 sets it to (start >> 7) & 7 (probe only). */
int dwg_gemm_tune_halo(int mode, int base_offset_mode);
int dwg_gemm_last_halo(void);
int dwg_gemm_last_pair(void);
/* dwg_gemm_f16_ln: GEMM (optionally batched) C[M,N] = act(A[M,K] B[N,K]^T ...) + residual (fp16 output) with
 *  (a) a LayerNorm of the ACTIVATION operand folded into the epilogue -- the operand holds the un-normalised activations, gamma
 *      is pre-multiplied into the weight (W' = W * gamma), c1 = row sums of W', beta enters through the bias (W beta + bias):
 *        ln_mode 1 (activations = A rows): C[m][n] = rstd_m (acc - mean_m c1[n]) + bias[n]
 *        ln_mode 2 (activations = B rows): C[m][n] = rstd_n (acc - mean_n c1[m]) + ln_rowbias[m]     (V^T = Wv' X^T)
 *      ln_stats i64 [rows,2] = fixed-point (2^-20) sum and sum of squares of each activation row over ln_dim features;
 *  (b) rowstats_out (NULL = off): i64 [M,2], PRE-ZEROED, receives the same statistics of the OUTPUT rows (of the fp16 values stored)
 *      -- the LayerNorm statistics of the layer that consumes C.  No LayerNorm kernel and no normalised copy exist.
 *  Replaces the three nn.LayerNorm of diffusers' BasicTransformerBlock behind core/guidance/controlnet.py:98-114. */
int dwg_gemm_f16_ln(const void* A, int64_t lda, int64_t a_b, const void* B, int64_t ldb, int64_t b_b, void* C, int64_t ldc, int64_t c_b,
                    int M, int N, int K, int nb /* batch count; a_b == 0 broadcasts A; statistics index = batch * rows + row */,
                    const float* bias, const void* residual, int64_t ldr, int64_t r_b, int act,
                    int ln_mode, const void* ln_stats, const float* ln_c1, const float* ln_rowbias, int ln_dim, float ln_eps,
                    void* rowstats_out, void* workspace, int64_t workspace_bytes, void* colstats, int colstats_rows, void* stream);
/* Split-K scratch lane (0 or 1) used by the launches that follow: GEMMs enqueued on two streams that may run
 * concurrently (ControlNet beside the UNet encoder, dwg/diffusion/guidance.py) must use different lanes. */
int dwg_gemm_set_lane(int lane);
int dwg_raster_probe(void* dev_u64_tiles_x6);   /* debug per-tile timeline of the forward render, NULL = off; csrc/raster_fwd.cu */
int dwg_gemm_trace(void* dev_u64x8);   /* debug timeline of CTA 0 (globaltimer ns), NULL = off; see csrc/gemm_tcgen05.cu */
int dwg_gemm_last_key(int* out6);   /* planner key of the last launch: m_tiles, nz, N, k-iterations, epilogue kind, has_residual */

/* dwg_conv2d_nhwc_f16: implicit-GEMM convolution, no im2col buffer.
 *   x [Nimg,H,W,Cin] bf16 (Cin % 8 == 0), w [Cout,k,k,Cin] bf16, y [Nimg,Ho,Wo,Cout] bf16/fp32,
 *   ksize 1|3, stride 1|2, zero padding pad_h/pad_w on the top/left (bottom/right implied by
 *   Ho/Wo: covers the VAE's asymmetric (0,1,0,1) padding); bias [Cout], bias2_per_image
 *   [Nimg,Cout] (time embedding), residual [Nimg,Ho,Wo,Cout] bf16. */
int dwg_conv2d_nhwc_f16_ws(const void* x, const void* w, void* y, int out_f16,
                           int Nimg, int H, int W, int Cin, int Cout, int ksize, int stride,
                           int pad_h, int pad_w, int Ho, int Wo,
                           const float* bias, const float* bias2_per_image,
                           const void* residual, int act, void* workspace, int64_t workspace_bytes, void* colstats, void* stream);
int dwg_conv2d_nhwc_f16(const void* x, const void* w, void* y, int out_f16,
                         int Nimg, int H, int W, int Cin, int Cout, int ksize, int stride,
                         int pad_h, int pad_w, int Ho, int Wo,
                         const float* bias, const float* bias2_per_image,
                         const void* residual, int act, void* stream);

/* ------------------------------------------------------------------------------------------
 * R14/R15  Normalisation / activation / softmax kernels of the diffusion blocks (NHWC bf16).
 * Replace torch.nn.GroupNorm / LayerNorm / softmax / GEGLU / SiLU calls inside the diffusers
 * modules, and the CFG + SDS-gradient arithmetic of core/guidance/basic.py:595-603,642.
 */
/* Shared-memory carve-out preference (percent, -1 = driver default) of the streaming kernels below: 100 keeps the SMs in
 * the partition the tcgen05 kernels need, so GEMM <-> norm alternations do not re-partition L1/shared memory. */
int dwg_nn_set_carveout(int percent);
/* y = [SiLU](GroupNorm_G(x));  x,y [N,HW,C] fp16; stats [N,G,2] i64 workspace: (sum, sumsq) in 2^-20 fixed point,
 * accumulated with integer atomics so that the statistics are run-to-run deterministic; kept for bwd.
 * do_silu: bit 0 = apply SiLU after the normalisation; bit 1 = `stats` (`bstats` in the backward) was pre-zeroed by the
 * caller (e.g. one memset of an arena that serves every GroupNorm of a step), skip the internal memset. */
int dwg_groupnorm_fwd(const void* x, const float* gamma, const float* beta, void* y, void* stats,
                      int N, int HW, int C, int G, float eps, int do_silu, void* stream);
/* kernels launched by the last dwg_groupnorm_fwd: 1 = one-launch cluster kernel (tensor fits the shared memory of 8 CTAs
 * per (image, 4-group slab): every UNet / ControlNet norm at batch 2), 2 = statistics + apply passes. */
int dwg_groupnorm_last_launches(void);
/* enable (1) / disable (0, default: measured slower inside the step) the one-launch cluster kernel */
int dwg_groupnorm_set_fused(int on);
/* y = [SiLU](GroupNorm_G(x)) from the column statistics the producing GEMM / conv epilogue accumulated (colstats i64 [4,N,C,2]);
 * stats_out i64 [N,G,2] (optional) receives the group statistics in the format dwg_groupnorm_bwd consumes. */
int dwg_groupnorm_apply_cs(const void* x, const float* gamma, const float* beta, void* y, const void* colstats, void* stats_out,
                           int N, int HW, int C, int G, float eps, int do_silu, void* stream);
/* The same for the channel concatenation [x1 (C1 channels) | x2 (C - C1)] of two tensors that are never stored side by side (the UNet
 * decoder's skip connections, diffusers UNet2DConditionModel up blocks behind core/guidance/controlnet.py:98-114): every source
 * brings its own column statistics; y [N,HW,C] is the normalised concatenation. */
int dwg_groupnorm_apply_cs2(const void* x1, const void* x2, int C1, const float* gamma, const float* beta, void* y, const void* colstats1,
                            const void* colstats2, void* stats_out, int N, int HW, int C, int G, float eps, int do_silu, void* stream);
/* dx = d/dx [SiLU](GroupNorm(x)) . dy  (+ dx_add if given);  bstats [N,G,2] i64 workspace (2^-36 fixed point) */
int dwg_groupnorm_bwd(const void* x, const void* dy, const void* stats, const float* gamma, const float* beta,
                      const void* dx_add, void* dx, void* bstats, int N, int HW, int C, int G, float eps,
                      int do_silu, void* stream);
int dwg_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y, int64_t rows, int C, float eps, void* stream);
/* in-place row softmax of bf16 scores [rows, cols_pad] (columns >= cols are written as 0) */
int dwg_softmax_rows(void* s, int64_t rows, int cols, int cols_pad, void* stream);
/* in place: dP -> dS = P * (dP - sum(dP * P)) */
int dwg_softmax_rows_bwd(const void* p, void* dp, int64_t rows, int cols_pad, void* stream);
/* y[rows, inner] = x[:, :inner] * gelu(x[:, inner:]) */
int dwg_geglu(const void* x, void* y, int64_t rows, int inner, void* stream);
/* mode 0: y = silu(x); 1: y = x + a; 2: y = a * silu'(x)   (n bf16 elements, n % 8 == 0) */
int dwg_eltwise_f16(const void* x, const void* a, void* y, int64_t n, int mode, void* stream);
/* Layout / type boundary of the diffusion path: the reference's planar fp32 tensors (images core/guidance/vae.py:34-40, latents
 * basic.py:595-603) <-> channels-last fp16 padded to 8 channels.  dst = scale * src + shift (padding channels zero); and back:
  * dst [N,C,H,W] fp32 = scale * src[..., :C], C <= 8 (rows of Cp elements; Cp need not be padded). */
int dwg_nchw_f32_to_nhwc_f16(const float* src, void* dst, int N, int C, int64_t HW, int Cp, float scale, float shift, void* stream);
int dwg_nhwc_f16_to_nchw_f32(const void* src, float* dst, int N, int C, int64_t HW, int Cp, float scale, void* stream);
/* noise_pred = eps_u + s (eps_c - eps_u);  grad = weight * (noise_pred - noise)   (fp32) */
int dwg_sds_grad(const float* eps_uncond, const float* eps_cond, const float* noise, float* grad, float* noise_pred,
                 float guidance_scale, float weight, int64_t n, void* stream);

/* Fused multi-head attention forward (no mask): out = softmax(q k^T * scale) v, scores never
 * leave the SM (tcgen05 QK^T -> TMEM -> online softmax -> tcgen05 PV).  Replaces diffusers'
 * attention processor (F.scaled_dot_product_attention) inside every BasicTransformerBlock.
 *   q [B,T,heads*hd] bf16 (row stride q_ld), k [B,Tk,heads*hd] (row stride k_ld),
 *   vt [B, heads*hd, Tkp] = V transposed (key index contiguous; batch stride vt_batch_stride
 *   elements; columns >= Tk finite),
 *   out [B,T,heads*hd] bf16.  hd multiple of 8, <= 128. */
int dwg_attention_fwd(const void* q, int64_t q_ld, const void* k, int64_t k_ld, const void* vt, int64_t Tkp,
                      int64_t vt_batch_stride, void* out, int B, int heads, int T, int Tk, int hd, float scale, void* stream);

/* ------------------------------------------------------------------------------------------
 * R1  GeneralLinearBlendSkinning.forward as DreamWaltzG.animate consumes it (core/human/inverse_lbs.py:570-784 with
 * smplx.lbs batch_rodrigues / blend_shapes / vertices2joints / batch_rigid_transform), without the full-mesh blend shapes.
 * dwg_glbs_joints (ONE CTA): pose parts (axis-angle, [1,3] / [1,63] / [1,45] as the reference passes them) + pose_mean [165],
 *   betas [n_betas] (extra_betas already added), expression [n_expr]; J_template [55,3]; JS [165, n_betas+n_expr] =
 *   J_regressor . shapedirs (constant, pre-multiplied by the caller); parents i32 [55]; transl [3] or NULL ->
 *   A [55,4,4] relative rigid joint transforms (tr['J_pose_rigid']), A_transl = transl o A (the joint transform the Gaussians are
 *   skinned with, avatar.py:1446-1460), pose_feature [486], shape_out [n_betas+n_expr], joints [55,3] (rest), posed_joints [55,3]
 *   (smplx posed joints + transl: keypoint source of the condition producer).
 * dwg_glbs_vertices (one warp per vertex): the composite V_shape_offset o V_pose_offset o V_pose_rigid [o transl] applied to
 *   points [Vp,3] of Vp PREDEFINED vertices only; shapedirs_sel [Vp,3,n_shape], posedirs_sel [Vp,3,486], weights_sel [Vp,55]
 *   are the per-vertex slices of the model tensors gathered once by the caller. */
int dwg_glbs_joints(const float* global_orient, const float* body_pose, const float* jaw_pose, const float* leye_pose,
                    const float* reye_pose, const float* left_hand_pose, const float* right_hand_pose, const float* pose_mean,
                    const float* betas, int n_betas, const float* expression, int n_expr,
                    const float* J_template, const float* JS, const int32_t* parents, const float* transl,
                    float* A, float* A_transl, float* pose_feature, float* shape_out, float* joints, float* posed_joints, void* stream);
int dwg_glbs_vertices(int Vp, int n_shape, const float* shape, const float* pose_feature, const float* A, const float* transl,
                      const float* shapedirs_sel, const float* posedirs_sel, const float* weights_sel, const float* points,
                      float* out, void* stream);

/* R5  Mesh-bound Gaussians (MeshBindingGaussianModel.get_positions / get_scales_and_quaternions, core/system/avatar.py:
 * 1016-1079, with compute_normal of utils/mesh.py:34-97).  vertex_coords [Vp,3]; triangles i32 [F,3]; adj_ptr i32 [Vp+1] /
 * adj_tri i32 [3F]: static vertex -> incident-triangle lists; bary [F*n_per_tri,3] RAW barycentric weights; scales_param
 * [F*n_per_tri,3] -> vertex_normals [Vp,3] (kept for the backward), positions [P,3], scales [P,3] (column 0 = 0),
 * quaternions [P,4] (real first, standardised).  Backward: gradients w.r.t. bary and scales_param (vertex coordinates are
 * constants of the step); NULL upstream gradients mean zero. */
int dwg_mesh_gaussians_fwd(int Vp, int F, int n_per_tri, const float* vertex_coords, const int32_t* triangles,
                           const int32_t* adj_ptr, const int32_t* adj_tri, const float* bary, const float* scales_param,
                           float* vertex_normals, float* positions, float* scales, float* quaternions, void* stream);
int dwg_mesh_gaussians_bwd(int F, int n_per_tri, const float* vertex_coords, const float* vertex_normals, const int32_t* triangles,
                           const float* bary, const float* scales_param, const float* g_positions, const float* g_scales,
                           const float* g_quaternions, float* g_bary, float* g_scales_param, void* stream);

/* ------------------------------------------------------------------------------------------
 * (f1) ControlNet condition producer.  Replaces the per-view CPU stage of the reference (core/human/smpl_condition.py:
 * 191-235 export_pose: numpy projection + Embree occlusion culling (utils/open3d.py:8-45) + cv2 drawing
 * (core/human/open_pose.py:282-333) + PIL -> tensor (core/guidance/controlnet.py:33-55)).
 * dwg_pose_keypoints_2d: kp_world [K,3] (K = 128: body 18, hands 21 + 21, face 51 + 17, smpl_condition.py:22);
 *   extrinsic_dev = DEVICE world->camera 4x4 (row-major); fx, fy, cx, cy = intrinsics at the condition size
 *   (data/camera/utils.py:116-147,233-242; fy < 0), overridden by intrinsics_dev = DEVICE (fx, fy, cx, cy) when not NULL (a captured
 *   CUDA graph is replayed with a new camera); depth / alpha [Hd,Wd] = the rasteriser's outputs of the same view or
 *   NULL (no occlusion culling) -> kp2d [K,2] pixels, NaN = behind the camera or occluded.
 * dwg_pose_image: kp2d [128,2] -> out [3,H,W] in [0,1] (RGB planes); flags: 1 body, 2 hands, 4 face, 8 flip_LR;
 *   hand_edge_colors_dev u8 [20,3] = rint(hsv_to_rgb(e / 20, 1, 1) * 255) (open_pose.py:211). */
int dwg_pose_keypoints_2d(const float* kp_world, int K, const float* extrinsic_dev, const float* intrinsics_dev,
                          float fx, float fy, float cx, float cy,
                          const float* depth, const float* alpha, int Hd, int Wd, float cond_w, float cond_h,
                          float thres_body, float thres_face, float thres_hand, float* kp2d, void* stream);
int dwg_pose_image(const float* kp2d, int H, int W, int flags, const uint8_t* hand_edge_colors_dev, float* out, void* stream);

/* ------------------------------------------------------------------------------------------
 * (f2) Fused multi-tensor Adam.  Replaces the torch.optim.Adam instances the reference steps every iteration
 * (core/trainer.py:880-882; groups of core/gaussian/gaussian_optimizer.py:49-141, core/system/avatar.py:1590-1635,
 * :1081-1094): ONE launch updates every parameter of the flat buffers (fp32, 16-byte aligned segments).
 *   params / grads / exp_avg / exp_avg_sq [n]; segment i covers elements [seg_end[i-1], seg_end[i]) and belongs to
 *   hyper-parameter group seg_group[i]; beta1 / beta2 / eps [n_group] (host arrays); lr_dev [n_group] DEVICE array
 *   (refreshed by the caller when a schedule changes it); step_dev DEVICE int64 step counter (incremented here).
 * Update rule = torch.optim.Adam(amsgrad=False, weight_decay=0): bias-corrected, eps added outside the square root. */
int dwg_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                  int n_seg, const int64_t* seg_end, const int32_t* seg_group,
                  int n_group, const float* beta1, const float* beta2, const float* eps,
                  const float* lr_dev, int64_t* step_dev, void* stream);

/* ------------------------------------------------------------------------------------------
 * (f4) Inference output stage.  Replaces the per-frame post-processing of Trainer.evaluate (core/trainer.py:1068-1084:
 * depth / 3, concat_alpha) + tensor2image (utils/image.py:52-61: (x * 255).clip(0, 255).astype(uint8), channel-last)
 * with ONE kernel from the rasteriser's planar fp32 outputs to interleaved uint8 frames ready for the encoder:
 *   image [3,H,W] -> rgb u8 [H,W,3];  image_fg [3,H,W] + alpha [H,W] -> rgba_fg u8 [H,W,4];
 *   depth [H,W] / depth_div -> depth_u8 [H,W];  alpha -> alpha_u8 [H,W].   NULL outputs are skipped. */
int dwg_frame_pack(const float* image, const float* image_fg, const float* depth, const float* alpha,
                   uint8_t* rgb, uint8_t* rgba_fg, uint8_t* depth_u8, uint8_t* alpha_u8,
                   int H, int W, float depth_div, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DWG_H_ */
