// R10: spherical-harmonic colour evaluation (degree <= 4), forward + backward.
// Reference: core/gaussian/spherical_harmonics.py:117-172, gaussian_utils.py:12-17,
// gaussian_renderer.py:72-105.  HBM-bound (216 B/Gaussian fwd at degree 3): each CTA stages
// its contiguous slab of coefficients through shared memory with 128-bit streaming loads,
// then one thread evaluates one Gaussian (row stride 3*sh_stride floats, odd multiples of 3
// -> at most 2-way bank conflicts, irrelevant next to the HBM time).
#include "common.cuh"

namespace dwg {
namespace {

constexpr int kPts = 128;

__constant__ float kC2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                             -1.0925484305920792f, 0.5462742152960396f};
__constant__ float kC3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f, 0.3731763325901154f,
                             -0.4570457994644658f, 1.445305721320277f, -0.5900435899266435f};
__constant__ float kC4[9] = {2.5033429417967046f, -1.7701307697799304f, 0.9461746957575601f, -0.6690465435572892f,
                             0.10578554691520431f, -0.6690465435572892f, 0.47308734787878004f,
                             -1.7701307697799304f, 0.6258357354491761f};
#define SH_C0 0.28209479177387814f
#define SH_C1 0.4886025119029199f

// Basis values Y[k] and (optionally) their gradients w.r.t. the unit direction.
template <bool GRAD>
__device__ __forceinline__ void sh_basis(int deg, float x, float y, float z, float* Y, float* dYx, float* dYy, float* dYz) {
    Y[0] = SH_C0;
    if (GRAD) { dYx[0] = dYy[0] = dYz[0] = 0.f; }
    if (deg < 1) return;
    Y[1] = -SH_C1 * y; Y[2] = SH_C1 * z; Y[3] = -SH_C1 * x;
    if (GRAD) {
        dYx[1] = 0; dYy[1] = -SH_C1; dYz[1] = 0;
        dYx[2] = 0; dYy[2] = 0; dYz[2] = SH_C1;
        dYx[3] = -SH_C1; dYy[3] = 0; dYz[3] = 0;
    }
    if (deg < 2) return;
    const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
    Y[4] = kC2[0] * xy; Y[5] = kC2[1] * yz; Y[6] = kC2[2] * (2.f * zz - xx - yy); Y[7] = kC2[3] * xz; Y[8] = kC2[4] * (xx - yy);
    if (GRAD) {
        dYx[4] = kC2[0] * y; dYy[4] = kC2[0] * x; dYz[4] = 0;
        dYx[5] = 0; dYy[5] = kC2[1] * z; dYz[5] = kC2[1] * y;
        dYx[6] = kC2[2] * -2.f * x; dYy[6] = kC2[2] * -2.f * y; dYz[6] = kC2[2] * 4.f * z;
        dYx[7] = kC2[3] * z; dYy[7] = 0; dYz[7] = kC2[3] * x;
        dYx[8] = kC2[4] * 2.f * x; dYy[8] = kC2[4] * -2.f * y; dYz[8] = 0;
    }
    if (deg < 3) return;
    Y[9] = kC3[0] * y * (3.f * xx - yy);
    Y[10] = kC3[1] * xy * z;
    Y[11] = kC3[2] * y * (4.f * zz - xx - yy);
    Y[12] = kC3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy);
    Y[13] = kC3[4] * x * (4.f * zz - xx - yy);
    Y[14] = kC3[5] * z * (xx - yy);
    Y[15] = kC3[6] * x * (xx - 3.f * yy);
    if (GRAD) {
        dYx[9] = kC3[0] * 6.f * xy; dYy[9] = kC3[0] * (3.f * xx - 3.f * yy); dYz[9] = 0;
        dYx[10] = kC3[1] * yz; dYy[10] = kC3[1] * xz; dYz[10] = kC3[1] * xy;
        dYx[11] = kC3[2] * -2.f * xy; dYy[11] = kC3[2] * (4.f * zz - xx - 3.f * yy); dYz[11] = kC3[2] * 8.f * yz;
        dYx[12] = kC3[3] * -6.f * xz; dYy[12] = kC3[3] * -6.f * yz; dYz[12] = kC3[3] * (6.f * zz - 3.f * xx - 3.f * yy);
        dYx[13] = kC3[4] * (4.f * zz - 3.f * xx - yy); dYy[13] = kC3[4] * -2.f * xy; dYz[13] = kC3[4] * 8.f * xz;
        dYx[14] = kC3[5] * 2.f * xz; dYy[14] = kC3[5] * -2.f * yz; dYz[14] = kC3[5] * (xx - yy);
        dYx[15] = kC3[6] * (3.f * xx - 3.f * yy); dYy[15] = kC3[6] * -6.f * xy; dYz[15] = 0;
    }
    if (deg < 4) return;
    Y[16] = kC4[0] * xy * (xx - yy);
    Y[17] = kC4[1] * yz * (3.f * xx - yy);
    Y[18] = kC4[2] * xy * (7.f * zz - 1.f);
    Y[19] = kC4[3] * yz * (7.f * zz - 3.f);
    Y[20] = kC4[4] * (zz * (35.f * zz - 30.f) + 3.f);
    Y[21] = kC4[5] * xz * (7.f * zz - 3.f);
    Y[22] = kC4[6] * (xx - yy) * (7.f * zz - 1.f);
    Y[23] = kC4[7] * xz * (xx - 3.f * yy);
    Y[24] = kC4[8] * (xx * (xx - 3.f * yy) - yy * (3.f * xx - yy));
    if (GRAD) {
        dYx[16] = kC4[0] * (3.f * xx * y - yy * y); dYy[16] = kC4[0] * (xx * x - 3.f * x * yy); dYz[16] = 0;
        dYx[17] = kC4[1] * 6.f * xy * z; dYy[17] = kC4[1] * z * (3.f * xx - 3.f * yy); dYz[17] = kC4[1] * y * (3.f * xx - yy);
        dYx[18] = kC4[2] * y * (7.f * zz - 1.f); dYy[18] = kC4[2] * x * (7.f * zz - 1.f); dYz[18] = kC4[2] * 14.f * xy * z;
        dYx[19] = 0; dYy[19] = kC4[3] * z * (7.f * zz - 3.f); dYz[19] = kC4[3] * y * (21.f * zz - 3.f);
        dYx[20] = 0; dYy[20] = 0; dYz[20] = kC4[4] * (140.f * zz * z - 60.f * z);
        dYx[21] = kC4[5] * z * (7.f * zz - 3.f); dYy[21] = 0; dYz[21] = kC4[5] * x * (21.f * zz - 3.f);
        dYx[22] = kC4[6] * 2.f * x * (7.f * zz - 1.f); dYy[22] = kC4[6] * -2.f * y * (7.f * zz - 1.f); dYz[22] = kC4[6] * (xx - yy) * 14.f * z;
        dYx[23] = kC4[7] * z * (3.f * xx - 3.f * yy); dYy[23] = kC4[7] * -6.f * xy * z; dYz[23] = kC4[7] * x * (xx - 3.f * yy);
        dYx[24] = kC4[8] * (4.f * xx * x - 12.f * x * yy); dYy[24] = kC4[8] * (-12.f * xx * y + 4.f * yy * y); dYz[24] = 0;
    }
}

__global__ void __launch_bounds__(kPts)
sh_fwd_kernel(const float* __restrict__ sh, int sh_stride, int deg, const float* __restrict__ pos,
              const float* __restrict__ campos, float* __restrict__ rgb, uint8_t* __restrict__ clamped,
              int64_t N, int aligned) {
    extern __shared__ __align__(16) float smem[];
    const int row = sh_stride * 3;
    float* sS = smem;                       // [kPts*row]
    float* sP = sS + ((kPts * row + 3) & ~3);   // [kPts*3] positions in, rgb out
    const int64_t n0 = (int64_t)blockIdx.x * kPts;
    const int n = (int)min((int64_t)kPts, N - n0);
    cta_load_floats(sS, sh + n0 * row, n * row, aligned);
    cta_load_floats(sP, pos + n0 * 3, n * 3, aligned);
    __syncthreads();
    const int t = threadIdx.x;
    float c[3];
    uint8_t mask = 0;
    if (t < n) {
        float vx = sP[t * 3] - campos[0], vy = sP[t * 3 + 1] - campos[1], vz = sP[t * 3 + 2] - campos[2];
        const float inv = 1.0f / fmaxf(sqrtf(vx * vx + vy * vy + vz * vz), 1e-12f);
        vx *= inv; vy *= inv; vz *= inv;
        float Y[25];
        sh_basis<false>(deg, vx, vy, vz, Y, nullptr, nullptr, nullptr);
        const int K = (deg + 1) * (deg + 1);
        c[0] = c[1] = c[2] = 0.f;
        const float* s = sS + t * row;
        for (int k = 0; k < K; k++) {
            c[0] = fmaf(Y[k], s[k * 3], c[0]); c[1] = fmaf(Y[k], s[k * 3 + 1], c[1]); c[2] = fmaf(Y[k], s[k * 3 + 2], c[2]);
        }
#pragma unroll
        for (int ch = 0; ch < 3; ch++) {
            c[ch] += 0.5f;
            if (c[ch] < 0.f) { c[ch] = 0.f; mask |= (1u << ch); }
        }
    }
    __syncthreads();
    if (t < n) {
        sP[t * 3] = c[0]; sP[t * 3 + 1] = c[1]; sP[t * 3 + 2] = c[2];
        if (clamped) clamped[n0 + t] = mask;
    }
    __syncthreads();
    cta_store_floats(rgb + n0 * 3, sP, n * 3, aligned);
}

__global__ void __launch_bounds__(kPts)
sh_bwd_kernel(const float* __restrict__ sh, int sh_stride, int deg, const float* __restrict__ pos,
              const float* __restrict__ campos, const uint8_t* __restrict__ clamped,
              const float* __restrict__ g_rgb, float* __restrict__ g_sh, float* __restrict__ g_pos,
              int64_t N, int aligned) {
    extern __shared__ __align__(16) float smem[];
    const int row = sh_stride * 3;
    float* sS = smem;                           // sh in, g_sh out
    float* sP = sS + ((kPts * row + 3) & ~3);   // pos in, g_pos out
    float* sG = sP + kPts * 3;                  // g_rgb
    const int64_t n0 = (int64_t)blockIdx.x * kPts;
    const int n = (int)min((int64_t)kPts, N - n0);
    cta_load_floats(sS, sh + n0 * row, n * row, aligned);
    cta_load_floats(sP, pos + n0 * 3, n * 3, aligned);
    cta_load_floats(sG, g_rgb + n0 * 3, n * 3, aligned);
    __syncthreads();
    const int t = threadIdx.x;
    if (t < n) {
        float vx = sP[t * 3] - campos[0], vy = sP[t * 3 + 1] - campos[1], vz = sP[t * 3 + 2] - campos[2];
        const float len = sqrtf(vx * vx + vy * vy + vz * vz);
        const float inv = 1.0f / fmaxf(len, 1e-12f);
        vx *= inv; vy *= inv; vz *= inv;
        float Y[25], dx[25], dy[25], dz[25];
        sh_basis<true>(deg, vx, vy, vz, Y, dx, dy, dz);
        const int K = (deg + 1) * (deg + 1);
        const uint8_t mask = clamped ? clamped[n0 + t] : 0;
        const float g[3] = {(mask & 1) ? 0.f : sG[t * 3], (mask & 2) ? 0.f : sG[t * 3 + 1], (mask & 4) ? 0.f : sG[t * 3 + 2]};
        float* s = sS + t * row;
        float gdx = 0.f, gdy = 0.f, gdz = 0.f;
        for (int k = 0; k < K; k++) {
            const float dot = s[k * 3] * g[0] + s[k * 3 + 1] * g[1] + s[k * 3 + 2] * g[2];
            gdx = fmaf(dx[k], dot, gdx); gdy = fmaf(dy[k], dot, gdy); gdz = fmaf(dz[k], dot, gdz);
            s[k * 3] = Y[k] * g[0]; s[k * 3 + 1] = Y[k] * g[1]; s[k * 3 + 2] = Y[k] * g[2];
        }
        for (int k = K * 3; k < row; k++) s[k] = 0.f;
        // normalisation backward: v/|v|
        const float dd = vx * gdx + vy * gdy + vz * gdz;
        sP[t * 3] = (gdx - vx * dd) * inv; sP[t * 3 + 1] = (gdy - vy * dd) * inv; sP[t * 3 + 2] = (gdz - vz * dd) * inv;
    }
    __syncthreads();
    cta_store_floats(g_sh + n0 * row, sS, n * row, aligned);
    if (g_pos) cta_store_floats(g_pos + n0 * 3, sP, n * 3, aligned);
}

}  // namespace
}  // namespace dwg

using namespace dwg;

static inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

extern "C" int dwg_sh_eval_fwd(const float* sh, int sh_stride, int deg, const float* pos, const float* campos,
                               float* rgb, uint8_t* clamped, int64_t N, void* stream) {
    DWG_REQUIRE(sh && pos && campos && rgb, "null pointer");
    DWG_REQUIRE(deg >= 0 && deg <= 4 && sh_stride >= (deg + 1) * (deg + 1) && sh_stride <= 32, "bad degree / stride");
    if (N == 0) return DWG_OK;
    const int aligned = al16(sh) && al16(pos) && al16(rgb);
    const size_t smem = sizeof(float) * (size_t)(((kPts * sh_stride * 3 + 3) & ~3) + kPts * 3);
    cudaFuncSetAttribute(sh_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    sh_fwd_kernel<<<(unsigned)ceil_div(N, kPts), kPts, smem, (cudaStream_t)stream>>>(sh, sh_stride, deg, pos, campos, rgb, clamped, N, aligned);
    return check_launch("dwg_sh_eval_fwd");
}

extern "C" int dwg_sh_eval_bwd(const float* sh, int sh_stride, int deg, const float* pos, const float* campos,
                               const uint8_t* clamped, const float* g_rgb, float* g_sh, float* g_pos,
                               int64_t N, void* stream) {
    DWG_REQUIRE(sh && pos && campos && g_rgb && g_sh, "null pointer");
    DWG_REQUIRE(deg >= 0 && deg <= 4 && sh_stride >= (deg + 1) * (deg + 1) && sh_stride <= 32, "bad degree / stride");
    if (N == 0) return DWG_OK;
    const int aligned = al16(sh) && al16(pos) && al16(g_rgb) && al16(g_sh) && al16(g_pos);
    const size_t smem = sizeof(float) * (size_t)(((kPts * sh_stride * 3 + 3) & ~3) + kPts * 6);
    cudaFuncSetAttribute(sh_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    sh_bwd_kernel<<<(unsigned)ceil_div(N, kPts), kPts, smem, (cudaStream_t)stream>>>(sh, sh_stride, deg, pos, campos, clamped, g_rgb, g_sh, g_pos, N, aligned);
    return check_launch("dwg_sh_eval_bwd");
}
