// R6: multi-resolution tiled/hash grid encoder (D=3, C=2), forward + backward.
// Reference: core/nerf/gridencoder/src/gridencoder.cu:66-366 and grid.py:149-166.
//
// Mapping: 16 consecutive lanes = the 16 levels of ONE point (a warp covers two points), so
//   * the 128-byte encoded row of a point is written by 16 lanes as one coalesced line
//     (the reference writes [L,B,C] and then permutes, grid.py:51,61);
//   * every lane has 8 independent 8-byte gathers in flight (table is 48 MB -> L2-resident
//     on B200's 126 MB L2);
//   * the backward reduces dL/dx over the 16 levels with shuffles instead of a second kernel
//     over a stored dy_dx buffer (the reference stores 384 B/point of dy_dx, grid.py:54).
// Index arithmetic is uint32 and bit-exact with the oracle and with the reference kernel.  The per-level scale /
// resolution (gridencoder.cu:138-139: exp2f(level * S) * H - 1, ceil(scale) + 1) are evaluated ONCE by
// dwg_grid_level_table ON THE DEVICE with the same exp2f the reference kernel calls per thread: a host libm exp2 is
// correctly rounded while the device function is not (<= 2 ulp), and one ulp of scale at the 4096-wide level moves
// the interpolation position enough to change outputs by 1e-4 (found when the oracle was pinned to the reference
// kernel's own outputs, tests/test_grid_golden.py).
#include "common.cuh"

namespace dwg {
namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ uint32_t grid_index(uint32_t gridtype, bool align_corners, uint32_t hashmap_size,
                                               uint32_t resolution, uint32_t px, uint32_t py, uint32_t pz) {
    uint32_t stride = 1, index = 0;
    const uint32_t step = align_corners ? resolution : resolution + 1;
    const uint32_t pg[3] = {px, py, pz};
#pragma unroll
    for (int d = 0; d < 3; d++) {
        if (stride <= hashmap_size) {
            index += pg[d] * stride;
            stride *= step;
        }
    }
    if (gridtype == 0 && stride > hashmap_size) index = (px * 1u) ^ (py * 2654435761u) ^ (pz * 805459861u);
    return index % hashmap_size;
}

struct LevelCtx {
    float pos[3], deriv[3];
    uint32_t pg[3];
    uint32_t hs, res;
    float scale;
    bool oob;
};

__device__ __forceinline__ void level_setup(LevelCtx& c, const float* __restrict__ x, int64_t b, float bound,
                                            const int32_t* __restrict__ offsets, const float* __restrict__ level_scale,
                                            const uint32_t* __restrict__ level_res, int l, bool align_corners, int interp) {
    c.oob = false;
    c.hs = (uint32_t)(offsets[l + 1] - offsets[l]);
    c.scale = level_scale[l];
    c.res = level_res[l];
    const float two_b = 2.0f * bound;
#pragma unroll
    for (int d = 0; d < 3; d++) {
        const float x01 = bound > 0.f ? __fdiv_rn(__fadd_rn(x[b * 3 + d], bound), two_b)      // grid.py:153
                                      : x[b * 3 + d];                                    // already in [0,1]
        if (x01 < 0.f || x01 > 1.f) c.oob = true;
        float p = __fmaf_rn(x01, c.scale, align_corners ? 0.0f : 0.5f);
        const float fl = floorf(p);
        c.pg[d] = (uint32_t)fl;
        p -= fl;
        // reference quirk, reproduced for parity (found by the golden vectors of the reference's own kernel): with LINEAR
        // interpolation `float pos_deriv[D] = {1.0f}` (gridencoder.cu:143) initialises only element 0, so dy/dx_1 and dy/dx_2
        // are multiplied by 0.  The avatar grid uses smoothstep, where every element is overwritten.
        c.deriv[d] = d == 0 ? 1.0f : 0.0f;
        if (interp == 1) {
            c.deriv[d] = 6.f * p * (1.0f - p);
            p = p * p * (3.0f - 2.0f * p);
        }
        c.pos[d] = p;
    }
}

__global__ void __launch_bounds__(kThreads)
grid_fwd_kernel(const float* __restrict__ x, float bound, const float2* __restrict__ table,
                const int32_t* __restrict__ offsets, const float* __restrict__ level_scale,
                const uint32_t* __restrict__ level_res, float* __restrict__ out, int64_t osb, int64_t osl,
                float* __restrict__ dy_dx, int64_t B, int L, int LP, int gridtype, bool align_corners, int interp) {
    const int64_t gid = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    const int64_t b = gid / LP;                 // LP = lanes per point = L rounded up to a power of two (<= 32)
    const int l = (int)(gid - b * LP);
    if (b >= B || l >= L) return;
    LevelCtx c;
    level_setup(c, x, b, bound, offsets, level_scale, level_res, l, align_corners, interp);
    float2* o = reinterpret_cast<float2*>(out + b * osb + (int64_t)l * osl);
    float* dd = dy_dx ? dy_dx + b * (int64_t)L * 6 + l * 6 : nullptr;
    if (c.oob) {
        *o = make_float2(0.f, 0.f);
        if (dd) {
#pragma unroll
            for (int k = 0; k < 6; k++) dd[k] = 0.f;
        }
        return;
    }
    const float2* g = table + offsets[l];
    float2 v[8];
#pragma unroll
    for (int idx = 0; idx < 8; idx++) {
        const uint32_t gi = grid_index(gridtype, align_corners, c.hs, c.res, c.pg[0] + (idx & 1),
                                       c.pg[1] + ((idx >> 1) & 1), c.pg[2] + ((idx >> 2) & 1));
        v[idx] = __ldg(g + gi);
    }
    float r0 = 0.f, r1 = 0.f;
#pragma unroll
    for (int idx = 0; idx < 8; idx++) {
        float w = 1.f;
#pragma unroll
        for (int d = 0; d < 3; d++) w *= ((idx >> d) & 1) ? c.pos[d] : 1.f - c.pos[d];
        r0 += w * v[idx].x;
        r1 += w * v[idx].y;
    }
    *o = make_float2(r0, r1);
    if (dd) {
#pragma unroll
        for (int gd = 0; gd < 3; gd++) {
            float g0 = 0.f, g1 = 0.f;
#pragma unroll
            for (int idx = 0; idx < 4; idx++) {
                float w = c.scale;
                int corner = 0;
#pragma unroll
                for (int nd = 0; nd < 2; nd++) {
                    const int d = (nd >= gd) ? nd + 1 : nd;
                    const int bit = (idx >> nd) & 1;
                    w *= bit ? c.pos[d] : 1.f - c.pos[d];
                    corner |= bit << d;
                }
                const float2 lo = v[corner], hi = v[corner | (1 << gd)];
                g0 += w * (hi.x - lo.x) * c.deriv[gd];
                g1 += w * (hi.y - lo.y) * c.deriv[gd];
            }
            dd[gd * 2] = g0; dd[gd * 2 + 1] = g1;
        }
    }
}

template <bool WANT_GX>
__global__ void __launch_bounds__(kThreads)
grid_bwd_kernel(const float* __restrict__ grad, int64_t gsb, int64_t gsl, const float* __restrict__ x, float bound,
                const float2* __restrict__ table, const int32_t* __restrict__ offsets,
                const float* __restrict__ level_scale, const uint32_t* __restrict__ level_res,
                float2* __restrict__ g_table, float* __restrict__ g_x, int64_t B, int L, int LP, int gridtype,
                bool align_corners, int interp) {
    const int64_t gid = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    int64_t b = gid / LP;
    int l = (int)(gid - b * LP);
    const bool active = b < B && l < L;
    if (b >= B) b = B - 1;           // keep the lane for the shuffles below
    if (l >= L) l = L - 1;
    LevelCtx c;
    level_setup(c, x, b, bound, offsets, level_scale, level_res, l, align_corners, interp);
    float gxv[3] = {0.f, 0.f, 0.f};
    if (active && !c.oob) {
        const float2 gr = *reinterpret_cast<const float2*>(grad + b * gsb + (int64_t)l * gsl);
        float2* gt = g_table + offsets[l];
        const float2* g = table + offsets[l];
        float2 v[8];
#pragma unroll
        for (int idx = 0; idx < 8; idx++) {
            const uint32_t gi = grid_index(gridtype, align_corners, c.hs, c.res, c.pg[0] + (idx & 1),
                                           c.pg[1] + ((idx >> 1) & 1), c.pg[2] + ((idx >> 2) & 1));
            float w = 1.f;
#pragma unroll
            for (int d = 0; d < 3; d++) w *= ((idx >> d) & 1) ? c.pos[d] : 1.f - c.pos[d];
            atomicAdd(gt + gi, make_float2(w * gr.x, w * gr.y));     // 8-byte vector atomic (sm_90+)
            if (WANT_GX) v[idx] = __ldg(g + gi);
        }
        if (WANT_GX) {
#pragma unroll
            for (int gd = 0; gd < 3; gd++) {
                float acc = 0.f;
#pragma unroll
                for (int idx = 0; idx < 4; idx++) {
                    float w = c.scale;
                    int corner = 0;
#pragma unroll
                    for (int nd = 0; nd < 2; nd++) {
                        const int d = (nd >= gd) ? nd + 1 : nd;
                        const int bit = (idx >> nd) & 1;
                        w *= bit ? c.pos[d] : 1.f - c.pos[d];
                        corner |= bit << d;
                    }
                    const float2 lo = v[corner], hi = v[corner | (1 << gd)];
                    acc += gr.x * (w * (hi.x - lo.x) * c.deriv[gd]) + gr.y * (w * (hi.y - lo.y) * c.deriv[gd]);
                }
                gxv[gd] = acc;
            }
        }
    }
    if (WANT_GX) {
        // sum over the LP lanes of this point (LP is a power of two <= 32; lanes >= L hold zeros)
#pragma unroll
        for (int d = 0; d < 3; d++) {
            float vsum = gxv[d];
            for (int off = LP >> 1; off > 0; off >>= 1) vsum += __shfl_xor_sync(0xffffffffu, vsum, off);
            gxv[d] = vsum;
        }
        if (active && l == 0) {
            const float s = bound > 0.f ? 1.0f / (2.0f * bound) : 1.0f;
            g_x[b * 3] = gxv[0] * s; g_x[b * 3 + 1] = gxv[1] * s; g_x[b * 3 + 2] = gxv[2] * s;
        }
    }
}

}  // namespace
}  // namespace dwg

using namespace dwg;

namespace dwg {
namespace {
__global__ void level_table_kernel(float S, uint32_t H, int L, float* __restrict__ level_scale, uint32_t* __restrict__ level_res) {
    const uint32_t level = threadIdx.x;
    if ((int)level >= L) return;
    const float scale = exp2f(level * S) * H - 1.0f;              // gridencoder.cu:138, same expression, same device function
    level_scale[level] = scale;
    level_res[level] = (uint32_t)ceil(scale) + 1;                 // gridencoder.cu:139
}
}  // namespace
}  // namespace dwg

extern "C" int dwg_grid_level_table(float S, int H, int L, float* level_scale, uint32_t* level_res, void* stream) {
    DWG_REQUIRE(level_scale && level_res && L > 0 && L <= 32 && H > 0, "bad arguments");
    dwg::level_table_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(S, (uint32_t)H, L, level_scale, level_res);
    return dwg::check_launch("dwg_grid_level_table");
}

extern "C" int dwg_grid_encode_fwd(const float* x, float bound, const float* table, const int32_t* offsets,
                                   const float* level_scale, const uint32_t* level_res, float* out,
                                   int64_t out_stride_b, int64_t out_stride_l, float* dy_dx, int64_t B, int L,
                                   int gridtype, int align_corners, int interp, void* stream) {
    DWG_REQUIRE(x && table && offsets && level_scale && level_res && out, "null pointer");
    DWG_REQUIRE(L >= 1 && L <= 32, "1 <= L <= 32");
    DWG_REQUIRE((out_stride_b % 2) == 0 && (out_stride_l % 2) == 0 && (reinterpret_cast<uintptr_t>(out) & 7) == 0,
                "out must be 8-byte aligned with even strides");
    DWG_REQUIRE(gridtype == 0 || gridtype == 1, "gridtype must be 0 (hash) or 1 (tiled)");
    DWG_REQUIRE(interp == 0 || interp == 1, "interp must be 0 (linear) or 1 (smoothstep)");
    if (B == 0) return DWG_OK;
    int LP = 1;
    while (LP < L) LP <<= 1;
    const int64_t threads = B * LP;
    grid_fwd_kernel<<<(unsigned)ceil_div(threads, kThreads), kThreads, 0, (cudaStream_t)stream>>>(
        x, bound, reinterpret_cast<const float2*>(table), offsets, level_scale, level_res, out, out_stride_b,
        out_stride_l, dy_dx, B, L, LP, gridtype, align_corners != 0, interp);
    return check_launch("dwg_grid_encode_fwd");
}

extern "C" int dwg_grid_encode_bwd(const float* grad, int64_t g_stride_b, int64_t g_stride_l, const float* x,
                                   float bound, const float* table, const int32_t* offsets,
                                   const float* level_scale, const uint32_t* level_res, float* g_table, float* g_x,
                                   int64_t B, int L, int gridtype, int align_corners, int interp, void* stream) {
    DWG_REQUIRE(grad && x && table && offsets && level_scale && level_res && g_table, "null pointer");
    DWG_REQUIRE(L >= 1 && L <= 32, "1 <= L <= 32");
    DWG_REQUIRE((g_stride_b % 2) == 0 && (g_stride_l % 2) == 0 && (reinterpret_cast<uintptr_t>(grad) & 7) == 0,
                "grad must be 8-byte aligned with even strides");
    DWG_REQUIRE(gridtype == 0 || gridtype == 1, "gridtype must be 0 (hash) or 1 (tiled)");
    if (B == 0) return DWG_OK;
    int LP = 1;
    while (LP < L) LP <<= 1;
    const int64_t threads = B * LP;
    const unsigned grid = (unsigned)ceil_div(threads, kThreads);
    cudaStream_t st = (cudaStream_t)stream;
    if (g_x)
        grid_bwd_kernel<true><<<grid, kThreads, 0, st>>>(grad, g_stride_b, g_stride_l, x, bound,
            reinterpret_cast<const float2*>(table), offsets, level_scale, level_res,
            reinterpret_cast<float2*>(g_table), g_x, B, L, LP, gridtype, align_corners != 0, interp);
    else
        grid_bwd_kernel<false><<<grid, kThreads, 0, st>>>(grad, g_stride_b, g_stride_l, x, bound,
            reinterpret_cast<const float2*>(table), offsets, level_scale, level_res,
            reinterpret_cast<float2*>(g_table), g_x, B, L, LP, gridtype, align_corners != 0, interp);
    return check_launch("dwg_grid_encode_bwd");
}
