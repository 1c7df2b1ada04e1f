// R2/R3: fused linear-blend skinning of Gaussian positions + quaternions (fwd + bwd).
//
// One pass over W[N,J] per direction: the per-Gaussian blended 3x4 transform is built in
// registers and applied to the position and (optionally) the quaternion in the same thread;
// the reference issues >= 5 einsum passes over W (core/human/inverse_lbs.py:208-209,235).
// HBM-bound: 276 B/Gaussian fwd (220 W + 28 in + 28 out).  Each CTA stages its contiguous
// W / x / q slab through shared memory with coalesced 128-bit streaming loads; per-thread
// rows are then read at stride J (J = 55 is odd -> conflict-free banks).
#include <initializer_list>

#include "common.cuh"

namespace dwg {
namespace {

constexpr int kPts = 128;       // Gaussians per CTA == threads per CTA

struct Mat3 { float m[9]; };

// pytorch3d.transforms.quaternion_to_matrix (real-first, two_s = 2/|q|^2)
__device__ __forceinline__ void quat_to_mat(const float q[4], float Q[9]) {
    const float r = q[0], i = q[1], j = q[2], k = q[3];
    const float two_s = 2.0f / (r * r + i * i + j * j + k * k);
    Q[0] = 1.f - two_s * (j * j + k * k); Q[1] = two_s * (i * j - k * r); Q[2] = two_s * (i * k + j * r);
    Q[3] = two_s * (i * j + k * r); Q[4] = 1.f - two_s * (i * i + k * k); Q[5] = two_s * (j * k - i * r);
    Q[6] = two_s * (i * k - j * r); Q[7] = two_s * (j * k + i * r); Q[8] = 1.f - two_s * (i * i + j * j);
}

// pytorch3d 0.7.5 matrix_to_quaternion: best-conditioned candidate (argmax q_abs, first on ties)
__device__ __forceinline__ int mat_to_quat(const float m[9], float out[4], float* a_sel, float num[4]) {
    const float e[4] = {1.f + m[0] + m[4] + m[8], 1.f + m[0] - m[4] - m[8],
                        1.f - m[0] + m[4] - m[8], 1.f - m[0] - m[4] + m[8]};
    float a[4];
#pragma unroll
    for (int c = 0; c < 4; c++) a[c] = e[c] > 0.f ? sqrtf(e[c]) : 0.f;
    int s = 0;
#pragma unroll
    for (int c = 1; c < 4; c++) if (a[c] > a[s]) s = c;
    const float a2 = a[s] * a[s];
    switch (s) {
        case 0: num[0] = a2; num[1] = m[7] - m[5]; num[2] = m[2] - m[6]; num[3] = m[3] - m[1]; break;
        case 1: num[0] = m[7] - m[5]; num[1] = a2; num[2] = m[3] + m[1]; num[3] = m[2] + m[6]; break;
        case 2: num[0] = m[2] - m[6]; num[1] = m[3] + m[1]; num[2] = a2; num[3] = m[5] + m[7]; break;
        default: num[0] = m[3] - m[1]; num[1] = m[6] + m[2]; num[2] = m[7] + m[5]; num[3] = a2; break;
    }
    const float d = 2.0f * fmaxf(a[s], 0.1f);
#pragma unroll
    for (int c = 0; c < 4; c++) out[c] = num[c] / d;
    *a_sel = a[s];
    return s;
}

// Blend the 3x4 joint transforms: M[r*4+c] = sum_j w[j] * A[j][r][c]
__device__ __forceinline__ void blend(const float* __restrict__ w, int w_stride, const float* __restrict__ sA, int J, float M[12]) {
#pragma unroll
    for (int k = 0; k < 12; k++) M[k] = 0.f;
    for (int j = 0; j < J; j++) {
        const float wj = w[j * w_stride];
        const float* a = sA + j * 12;
#pragma unroll
        for (int k = 0; k < 12; k++) M[k] = fmaf(wj, a[k], M[k]);
    }
}

template <bool HAS_Q>
__global__ void __launch_bounds__(kPts)
lbs_skin_fwd_kernel(const float* __restrict__ W, const float* __restrict__ A, const float* __restrict__ x,
                    const float* __restrict__ q, float* __restrict__ x_out, float* __restrict__ q_out,
                    int64_t N, int J, int aligned) {
    extern __shared__ __align__(16) float smem[];
    float* sW = smem;                    // [kPts*J]
    float* sA = sW + kPts * J;           // [J*12]
    float* sX = sA + ((J * 12 + 3) & ~3);  // [kPts*3]
    float* sQ = sX + kPts * 3;           // [kPts*4]
    const int64_t n0 = (int64_t)blockIdx.x * kPts;
    const int n = (int)min((int64_t)kPts, N - n0);
    cta_load_floats(sW, W + n0 * J, n * J, aligned);
    cta_load_floats(sX, x + n0 * 3, n * 3, aligned);
    if (HAS_Q) cta_load_floats(sQ, q + n0 * 4, n * 4, aligned);
    for (int i = threadIdx.x; i < J * 12; i += blockDim.x) sA[i] = A[(i / 12) * 16 + (i % 12)];
    __syncthreads();
    const int t = threadIdx.x;
    float xo[3], qo[4];
    if (t < n) {
        float M[12];
        blend(sW + t * J, 1, sA, J, M);
        const float px = sX[t * 3], py = sX[t * 3 + 1], pz = sX[t * 3 + 2];
        xo[0] = M[0] * px + M[1] * py + M[2] * pz + M[3];
        xo[1] = M[4] * px + M[5] * py + M[6] * pz + M[7];
        xo[2] = M[8] * px + M[9] * py + M[10] * pz + M[11];
        if (HAS_Q) {
            float qq[4] = {sQ[t * 4], sQ[t * 4 + 1], sQ[t * 4 + 2], sQ[t * 4 + 3]};
            float Q[9], D[9];
            quat_to_mat(qq, Q);
            // B = F Q (rows 1,2 negated); C = R B; D = F C
#pragma unroll
            for (int r = 0; r < 3; r++)
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    float acc = M[r * 4 + 0] * Q[c] - M[r * 4 + 1] * Q[3 + c] - M[r * 4 + 2] * Q[6 + c];
                    D[r * 3 + c] = (r == 0) ? acc : -acc;
                }
            float a_sel, num[4];
            mat_to_quat(D, qo, &a_sel, num);
        }
    }
    __syncthreads();                                    // everyone done reading sX / sQ
    if (t < n) {
        sX[t * 3] = xo[0]; sX[t * 3 + 1] = xo[1]; sX[t * 3 + 2] = xo[2];
        if (HAS_Q) { sQ[t * 4] = qo[0]; sQ[t * 4 + 1] = qo[1]; sQ[t * 4 + 2] = qo[2]; sQ[t * 4 + 3] = qo[3]; }
    }
    __syncthreads();
    cta_store_floats(x_out + n0 * 3, sX, n * 3, aligned);
    if (HAS_Q) cta_store_floats(q_out + n0 * 4, sQ, n * 4, aligned);
}

// Backward.  Recomputes M from W (one more pass over W: 304 B/Gaussian).
template <bool HAS_Q, bool WANT_GW, bool WANT_GA>
__global__ void __launch_bounds__(kPts)
lbs_skin_bwd_kernel(const float* __restrict__ W, const float* __restrict__ A, const float* __restrict__ x,
                    const float* __restrict__ q, const float* __restrict__ gxo, const float* __restrict__ gqo,
                    float* __restrict__ gx, float* __restrict__ gq, float* __restrict__ gW, float* __restrict__ gA,
                    int64_t N, int J, int aligned) {
    extern __shared__ __align__(16) float smem[];
    float* sW = smem;                      // [kPts*J]  (re-used for gW)
    float* sA = sW + kPts * J;             // [J*12]
    float* sX = sA + ((J * 12 + 3) & ~3);  // [kPts*3]
    float* sQ = sX + kPts * 3;             // [kPts*4]
    float* sGX = sQ + kPts * 4;            // [kPts*3]
    float* sGQ = sGX + kPts * 3;           // [kPts*4]
    float* sGA = sGQ + kPts * 4;           // [J*12] block-local accumulation of g_A
    const int64_t n0 = (int64_t)blockIdx.x * kPts;
    const int n = (int)min((int64_t)kPts, N - n0);
    cta_load_floats(sW, W + n0 * J, n * J, aligned);
    cta_load_floats(sX, x + n0 * 3, n * 3, aligned);
    cta_load_floats(sGX, gxo + n0 * 3, n * 3, aligned);
    if (HAS_Q) {
        cta_load_floats(sQ, q + n0 * 4, n * 4, aligned);
        cta_load_floats(sGQ, gqo + n0 * 4, n * 4, aligned);
    }
    for (int i = threadIdx.x; i < J * 12; i += blockDim.x) {
        sA[i] = A[(i / 12) * 16 + (i % 12)];
        if (WANT_GA) sGA[i] = 0.f;
    }
    __syncthreads();
    const int t = threadIdx.x;
    float gM[12];
    float gxv[3], gqv[4];
    if (t < n) {
        float M[12];
        blend(sW + t * J, 1, sA, J, M);
        const float px = sX[t * 3], py = sX[t * 3 + 1], pz = sX[t * 3 + 2];
        const float g0 = sGX[t * 3], g1 = sGX[t * 3 + 1], g2 = sGX[t * 3 + 2];
        // x' = R x + T
        gxv[0] = M[0] * g0 + M[4] * g1 + M[8] * g2;
        gxv[1] = M[1] * g0 + M[5] * g1 + M[9] * g2;
        gxv[2] = M[2] * g0 + M[6] * g1 + M[10] * g2;
        gM[0] = g0 * px; gM[1] = g0 * py; gM[2] = g0 * pz; gM[3] = g0;
        gM[4] = g1 * px; gM[5] = g1 * py; gM[6] = g1 * pz; gM[7] = g1;
        gM[8] = g2 * px; gM[9] = g2 * py; gM[10] = g2 * pz; gM[11] = g2;
        if (HAS_Q) {
            float qq[4] = {sQ[t * 4], sQ[t * 4 + 1], sQ[t * 4 + 2], sQ[t * 4 + 3]};
            float Q[9], D[9], qo[4], a_sel, num[4];
            quat_to_mat(qq, Q);
#pragma unroll
            for (int r = 0; r < 3; r++)
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    float acc = M[r * 4 + 0] * Q[c] - M[r * 4 + 1] * Q[3 + c] - M[r * 4 + 2] * Q[6 + c];
                    D[r * 3 + c] = (r == 0) ? acc : -acc;
                }
            const int s = mat_to_quat(D, qo, &a_sel, num);
            const float go[4] = {sGQ[t * 4], sGQ[t * 4 + 1], sGQ[t * 4 + 2], sGQ[t * 4 + 3]};
            // ---- mat_to_quat backward (selected candidate only) ----
            const float d = 2.0f * fmaxf(a_sel, 0.1f);
            float dnum[4], dd = 0.f;
#pragma unroll
            for (int c = 0; c < 4; c++) { dnum[c] = go[c] / d; dd -= go[c] * num[c] / (d * d); }
            // d = 2*max(a,0.1); diagonal numerator a^2 = e_s (when e_s > 0)
            float de = 0.f;
            if (a_sel > 0.f) {
                de = dnum[s];
                if (a_sel > 0.1f) de += dd * 2.0f / (2.0f * a_sel);
            }
            float dD[9];
#pragma unroll
            for (int c = 0; c < 9; c++) dD[c] = 0.f;
            // e_s = 1 + s0*m00 + s1*m11 + s2*m22
            const float sg0 = (s == 0 || s == 1) ? 1.f : -1.f;
            const float sg1 = (s == 0 || s == 2) ? 1.f : -1.f;
            const float sg2 = (s == 0 || s == 3) ? 1.f : -1.f;
            dD[0] += sg0 * de; dD[4] += sg1 * de; dD[8] += sg2 * de;
            switch (s) {
                case 0: dD[7] += dnum[1]; dD[5] -= dnum[1]; dD[2] += dnum[2]; dD[6] -= dnum[2]; dD[3] += dnum[3]; dD[1] -= dnum[3]; break;
                case 1: dD[7] += dnum[0]; dD[5] -= dnum[0]; dD[3] += dnum[2]; dD[1] += dnum[2]; dD[2] += dnum[3]; dD[6] += dnum[3]; break;
                case 2: dD[2] += dnum[0]; dD[6] -= dnum[0]; dD[3] += dnum[1]; dD[1] += dnum[1]; dD[5] += dnum[3]; dD[7] += dnum[3]; break;
                default: dD[3] += dnum[0]; dD[1] -= dnum[0]; dD[6] += dnum[1]; dD[2] += dnum[1]; dD[7] += dnum[2]; dD[5] += dnum[2]; break;
            }
            // D = F C -> dC = F dD ; C = R B -> dR = dC B^T, dB = R^T dC ; B = F Q -> dQ = F dB
            float dC[9], Bm[9];
#pragma unroll
            for (int c = 0; c < 3; c++) {
                dC[c] = dD[c]; dC[3 + c] = -dD[3 + c]; dC[6 + c] = -dD[6 + c];
                Bm[c] = Q[c]; Bm[3 + c] = -Q[3 + c]; Bm[6 + c] = -Q[6 + c];
            }
            float dQ[9];
#pragma unroll
            for (int r = 0; r < 3; r++)
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    gM[r * 4 + c] += dC[r * 3 + 0] * Bm[c * 3 + 0] + dC[r * 3 + 1] * Bm[c * 3 + 1] + dC[r * 3 + 2] * Bm[c * 3 + 2];
                    float db = M[0 * 4 + r] * dC[0 * 3 + c] + M[1 * 4 + r] * dC[1 * 3 + c] + M[2 * 4 + r] * dC[2 * 3 + c];
                    dQ[r * 3 + c] = (r == 0) ? db : -db;
                }
            // ---- quat_to_mat backward: Q = I + s P(q), s = 2/|q|^2 ----
            const float r_ = qq[0], i_ = qq[1], j_ = qq[2], k_ = qq[3];
            const float nn = r_ * r_ + i_ * i_ + j_ * j_ + k_ * k_;
            const float sc = 2.0f / nn;
            const float P[9] = {-(j_ * j_ + k_ * k_), i_ * j_ - k_ * r_, i_ * k_ + j_ * r_,
                                i_ * j_ + k_ * r_, -(i_ * i_ + k_ * k_), j_ * k_ - i_ * r_,
                                i_ * k_ - j_ * r_, j_ * k_ + i_ * r_, -(i_ * i_ + j_ * j_)};
            float ds = 0.f;
#pragma unroll
            for (int c = 0; c < 9; c++) ds += dQ[c] * P[c];
            const float dPr = -k_ * dQ[1] + j_ * dQ[2] + k_ * dQ[3] - i_ * dQ[5] - j_ * dQ[6] + i_ * dQ[7];
            const float dPi = j_ * dQ[1] + k_ * dQ[2] + j_ * dQ[3] - 2.f * i_ * dQ[4] - r_ * dQ[5] + k_ * dQ[6] + r_ * dQ[7] - 2.f * i_ * dQ[8];
            const float dPj = -2.f * j_ * dQ[0] + i_ * dQ[1] + r_ * dQ[2] + i_ * dQ[3] + k_ * dQ[5] - r_ * dQ[6] + k_ * dQ[7] - 2.f * j_ * dQ[8];
            const float dPk = -2.f * k_ * dQ[0] - r_ * dQ[1] + i_ * dQ[2] + r_ * dQ[3] - 2.f * k_ * dQ[4] + j_ * dQ[5] + i_ * dQ[6] + j_ * dQ[7];
            const float dsn = -ds * sc * sc;        // d s / d q_c = -s^2 q_c
            gqv[0] = sc * dPr + dsn * r_;
            gqv[1] = sc * dPi + dsn * i_;
            gqv[2] = sc * dPj + dsn * j_;
            gqv[3] = sc * dPk + dsn * k_;
        }
        if (WANT_GA) {
            for (int j = 0; j < J; j++) {
                const float wj = sW[t * J + j];
                if (wj != 0.f) {
#pragma unroll
                    for (int k = 0; k < 12; k++) atomicAdd(&sGA[j * 12 + k], wj * gM[k]);
                }
            }
        }
    }
    __syncthreads();        // all reads of sW / sX / sQ / sGX / sGQ complete
    if (t < n) {
        sX[t * 3] = gxv[0]; sX[t * 3 + 1] = gxv[1]; sX[t * 3 + 2] = gxv[2];
        if (HAS_Q) { sQ[t * 4] = gqv[0]; sQ[t * 4 + 1] = gqv[1]; sQ[t * 4 + 2] = gqv[2]; sQ[t * 4 + 3] = gqv[3]; }
        if (WANT_GW) {
            for (int j = 0; j < J; j++) {
                const float* a = sA + j * 12;
                float acc = 0.f;
#pragma unroll
                for (int k = 0; k < 12; k++) acc = fmaf(gM[k], a[k], acc);
                sW[t * J + j] = acc;
            }
        }
    }
    __syncthreads();
    cta_store_floats(gx + n0 * 3, sX, n * 3, aligned);
    if (HAS_Q) cta_store_floats(gq + n0 * 4, sQ, n * 4, aligned);
    if (WANT_GW) cta_store_floats(gW + n0 * J, sW, n * J, aligned);
    if (WANT_GA) {
        for (int i = threadIdx.x; i < J * 12; i += blockDim.x) {
            const float v = sGA[i];
            if (v != 0.f) atomicAdd(&gA[(i / 12) * 16 + (i % 12)], v);
        }
    }
}

inline bool all_aligned16(std::initializer_list<const void*> ps) {
    for (const void* p : ps) if (p && (reinterpret_cast<uintptr_t>(p) & 15)) return false;
    return true;
}

}  // namespace
}  // namespace dwg

using namespace dwg;

extern "C" int dwg_lbs_skin_fwd(const float* W, const float* A, const float* x, const float* q,
                                float* x_out, float* q_out, int64_t N, int J, void* stream) {
    DWG_REQUIRE(W && A && x && x_out, "null pointer");
    DWG_REQUIRE((q == nullptr) == (q_out == nullptr), "q and q_out must both be given or both NULL");
    DWG_REQUIRE(N >= 0 && J > 0 && J <= 128, "bad N / J");
    if (N == 0) return DWG_OK;
    const int aligned = all_aligned16({W, x, q, x_out, q_out}) ? 1 : 0;
    const size_t smem = sizeof(float) * (size_t)(kPts * J + ((J * 12 + 3) & ~3) + kPts * 7);
    const unsigned grid = (unsigned)ceil_div(N, kPts);
    cudaStream_t st = (cudaStream_t)stream;
    if (q) {
        cudaFuncSetAttribute(lbs_skin_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        lbs_skin_fwd_kernel<true><<<grid, kPts, smem, st>>>(W, A, x, q, x_out, q_out, N, J, aligned);
    } else {
        cudaFuncSetAttribute(lbs_skin_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        lbs_skin_fwd_kernel<false><<<grid, kPts, smem, st>>>(W, A, x, q, x_out, q_out, N, J, aligned);
    }
    return check_launch("dwg_lbs_skin_fwd");
}

template <bool HAS_Q, bool GW, bool GA>
static int launch_bwd(const float* W, const float* A, const float* x, const float* q, const float* gxo,
                      const float* gqo, float* gx, float* gq, float* gW, float* gA, int64_t N, int J,
                      int aligned, cudaStream_t st) {
    const size_t smem = sizeof(float) * (size_t)(kPts * J + ((J * 12 + 3) & ~3) + kPts * 14 + J * 12);
    cudaFuncSetAttribute(lbs_skin_bwd_kernel<HAS_Q, GW, GA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    lbs_skin_bwd_kernel<HAS_Q, GW, GA><<<(unsigned)ceil_div(N, kPts), kPts, smem, st>>>(
        W, A, x, q, gxo, gqo, gx, gq, gW, gA, N, J, aligned);
    return check_launch("dwg_lbs_skin_bwd");
}

extern "C" int dwg_lbs_skin_bwd(const float* W, const float* A, const float* x, const float* q,
                                const float* g_x_out, const float* g_q_out, float* g_x, float* g_q,
                                float* g_W, float* g_A, int64_t N, int J, void* stream) {
    DWG_REQUIRE(W && A && x && g_x_out && g_x, "null pointer");
    DWG_REQUIRE((q == nullptr) == (g_q == nullptr) && (q == nullptr) == (g_q_out == nullptr),
                "q, g_q_out, g_q must all be given or all NULL");
    DWG_REQUIRE(N >= 0 && J > 0 && J <= 128, "bad N / J");
    if (N == 0) return DWG_OK;
    const int aligned = all_aligned16({W, x, q, g_x_out, g_q_out, g_x, g_q, g_W}) ? 1 : 0;
    cudaStream_t st = (cudaStream_t)stream;
#define DWG_BWD(HQ, GW_, GA_) return launch_bwd<HQ, GW_, GA_>(W, A, x, q, g_x_out, g_q_out, g_x, g_q, g_W, g_A, N, J, aligned, st)
    if (q) {
        if (g_W && g_A) DWG_BWD(true, true, true);
        if (g_W) DWG_BWD(true, true, false);
        if (g_A) DWG_BWD(true, false, true);
        DWG_BWD(true, false, false);
    } else {
        if (g_W && g_A) DWG_BWD(false, true, true);
        if (g_W) DWG_BWD(false, true, false);
        if (g_A) DWG_BWD(false, false, true);
        DWG_BWD(false, false, false);
    }
#undef DWG_BWD
}
