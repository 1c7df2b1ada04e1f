// R7 / R8 / R9: the two per-Gaussian MLPs of DreamWaltzG.animate and their activations, fused.
//
//   nerf_opacity_and_color_net   MLP 32 -> 64 -> 64 -> 4, ReLU            (core/nerf/nerf_model.py:12-33,
//                                out[:,0] -> sigmoid opacity, out[:,1:4] -> sigmoid colour,
//                                core/system/avatar.py:1283-1290)
//   nerf_scale_and_quaternion_net DeformNetwork [enc(32) | body_pose(63)] -> 4 x (64, leaky_relu 0.01)
//                                -> heads warp(3) / scaling(3)            (core/deformation/deform_model.py:102-143;
//                                the rotation head is not consumed with use_non_rigid_rotations=False)
//   non_rigid_transform          pos = positions + 0.01 * warp, scales = clamp_max(exp(scaling) * 1e-3, 0.01)
//                                (core/system/avatar.py:1464-1498, shipped flags)
//
// The reference runs these as ~40 eager kernels forward (Linear = cuBLAS SGEMM + elementwise) and as
// many again in autograd, moving every [N,64] activation through HBM several times.  Here one
// forward and one backward kernel do all of it in plain fp32 FMA arithmetic -- exactly the
// reference's numeric type (the hidden width is 64: the work is 21 kMAC per Gaussian, 5.7 GFLOP per
// pass, far too small to be tensor-bound, and fp32 keeps parity with the reference tight):
//   * the 63 pose inputs are identical for every Gaussian: folded into the first-layer bias once
//     per CTA (their weight gradient is the outer product bias-gradient x pose);
//   * weights live in shared memory for the whole kernel (86 KB); a thread owns one Gaussian and
//     walks the layer chain with its activation column in shared memory (layout [feature][point],
//     conflict-free), weights arriving as broadcast 128-bit shared loads: 64 FMAs per 17 loads;
//   * forward saves the six hidden activations (fp32, [layer][feature][point], coalesced);
//   * backward: per 256-point tile and layer, dL/dW is a register-tiled 64x64xP product over the
//     tile (4x4 block per thread, accumulated in registers across ALL tiles of the CTA, written as
//     per-CTA partials and summed by a second tiny kernel: deterministic, no atomics), dL/dx is
//     again thread-per-point and in place.
#include "common.cuh"

namespace dwg {
namespace mlp {

constexpr int P = 256;            // points per tile == threads per CTA
constexpr int S = P + 4;          // backward smem row stride in words: 16-byte aligned rows, 4 banks apart
constexpr int HID = 64, ENC = 32, POSE = 63;

// flat fp32 parameter vector (also the layout of the gradient vector)
constexpr int S1W = 0, S1B = 2048, S2W = 2112, S2B = 6208, S3W = 6272, S3B = 6528;
constexpr int D0W = 6532, D0B = 8580, D1W = 8644, D1B = 12740, D2W = 12804, D2B = 16900, D3W = 16964, D3B = 21060;
constexpr int WPW = 21124, WPB = 21316, SCW = 21319, SCB = 21511, NPARAM = 21514;

struct FwdArgs {
    const float* enc;         // [N,32]
    const float* positions;   // [Nu,3]
    const float* params;      // [NPARAM]
    const float* w_pose;      // [64,63]  (layers.0.weight[:, 32:95])
    const float* pose;        // [63]
    float* colors;            // [N,3]
    float* opac;              // [N]
    float* pos_out;           // [Nu,3]
    float* scales;            // [Nu,3]
    float* acts_s;            // [2][64][Np]   or null
    float* acts_d;            // [4][64][Nup]  or null
    int64_t N, Nu, Np, Nup;
    float init_offset, init_scale, max_scale;
};

struct BwdArgs {
    const float* enc; const float* params;
    const float* colors; const float* opac; const float* scales;
    const float* acts_s; const float* acts_d;
    const float* g_colors; const float* g_opac; const float* g_pos; const float* g_scales;   // any may be null (= zero)
    float* g_enc;             // [N,32]
    float* partial;           // [gridDim.x][NPARAM]
    int64_t N, Nu, Np, Nup;
    float init_offset, max_scale;
};

enum { ACT_NONE = 0, ACT_RELU = 1, ACT_LRELU = 2 };

__device__ __forceinline__ float act_fn(float x, int act) {
    if (act == ACT_RELU) return fmaxf(x, 0.f);
    if (act == ACT_LRELU) return x > 0.f ? x : 0.01f * x;
    return x;
}
__device__ __forceinline__ float act_grad(float a, int act) {      // from the POST-activation value (same sign as the input)
    if (act == ACT_RELU) return a > 0.f ? 1.f : 0.f;
    if (act == ACT_LRELU) return a > 0.f ? 1.f : 0.01f;
    return 1.f;
}
__device__ __forceinline__ float sigmoidf(float x) { return 1.0f / (1.0f + expf(-x)); }

// thread-per-point layer: out[o] = act(b[o] + sum_k Wt[k][o] * col[k*stride]);  Wt = [K][O] in smem
template <int K, int O, int ACT>
__device__ __forceinline__ void layer_fwd(const float* __restrict__ Wt, const float* __restrict__ b, float* col, int stride,
                                          float (&out)[O]) {
#pragma unroll
    for (int o = 0; o < O; o++) out[o] = b[o];
#pragma unroll 2
    for (int k = 0; k < K; k++) {
        const float a = col[k * stride];
        const float4* w = reinterpret_cast<const float4*>(Wt + k * O);
#pragma unroll
        for (int o4 = 0; o4 < O / 4; o4++) {
            const float4 ww = w[o4];
            out[4 * o4] = fmaf(a, ww.x, out[4 * o4]); out[4 * o4 + 1] = fmaf(a, ww.y, out[4 * o4 + 1]);
            out[4 * o4 + 2] = fmaf(a, ww.z, out[4 * o4 + 2]); out[4 * o4 + 3] = fmaf(a, ww.w, out[4 * o4 + 3]);
        }
    }
#pragma unroll
    for (int o = 0; o < O; o++) out[o] = act_fn(out[o], ACT);
}

// hidden layer in place on the thread's activation column (+ optional save to global [O][Np])
template <int K, int ACT>
__device__ __forceinline__ void hidden_fwd(const float* __restrict__ Wt, const float* __restrict__ b, float* col, float* gsave, int64_t Np, bool valid) {
    float h[HID];
    layer_fwd<K, HID, ACT>(Wt, b, col, P, h);
#pragma unroll
    for (int o = 0; o < HID; o++) col[o * P] = h[o];
    if (gsave && valid) {
#pragma unroll
        for (int o = 0; o < HID; o++) gsave[o * Np] = h[o];
    }
}

__device__ __forceinline__ void load_transposed(float* dst, const float* __restrict__ src, int O, int K) {     // dst[k][o] = src[o][k]
    for (int i = threadIdx.x; i < O * K; i += blockDim.x) {
        const int o = i / K, k = i - o * K;
        dst[k * O + o] = src[i];
    }
}

__device__ __forceinline__ void load_enc_column(const float* __restrict__ enc, int64_t p, bool valid, float* col, int stride) {
    if (valid) {
        const float4* e = reinterpret_cast<const float4*>(enc + p * ENC);
#pragma unroll
        for (int j = 0; j < ENC / 4; j++) {
            const float4 v = __ldg(e + j);
            col[(4 * j) * stride] = v.x; col[(4 * j + 1) * stride] = v.y; col[(4 * j + 2) * stride] = v.z; col[(4 * j + 3) * stride] = v.w;
        }
    } else {
#pragma unroll
        for (int k = 0; k < ENC; k++) col[k * stride] = 0.f;
    }
}

__global__ void __launch_bounds__(P, 1)
mlp_fwd_kernel(const FwdArgs a) {
    extern __shared__ __align__(16) float sm[];
    float* Ws1 = sm;                       // [32][64]
    float* Ws2 = Ws1 + ENC * HID;          // [64][64]
    float* Ws3 = Ws2 + HID * HID;          // [64][4]
    float* Wd0 = Ws3 + HID * 4;            // [32][64]
    float* Wd1 = Wd0 + ENC * HID;
    float* Wd2 = Wd1 + HID * HID;
    float* Wd3 = Wd2 + HID * HID;
    float* Wdh = Wd3 + HID * HID;          // [64][8]: warp xyz, scaling xyz, 0, 0
    float* bs1 = Wdh + HID * 8;
    float* bs2 = bs1 + HID;
    float* bs3 = bs2 + HID;                // [4]
    float* bd0 = bs3 + 4;                  // effective: bias + W_pose @ pose
    float* bd1 = bd0 + HID;
    float* bd2 = bd1 + HID;
    float* bd3 = bd2 + HID;
    float* bdh = bd3 + HID;                // [8]
    float* act = bdh + 8;                  // [64][P]
    const int tid = threadIdx.x;
    const float* W = a.params;
    load_transposed(Ws1, W + S1W, HID, ENC);
    load_transposed(Ws2, W + S2W, HID, HID);
    load_transposed(Ws3, W + S3W, 4, HID);
    load_transposed(Wd0, W + D0W, HID, ENC);
    load_transposed(Wd1, W + D1W, HID, HID);
    load_transposed(Wd2, W + D2W, HID, HID);
    load_transposed(Wd3, W + D3W, HID, HID);
    for (int i = tid; i < HID * 8; i += P) {
        const int k = i >> 3, o = i & 7;
        Wdh[i] = o < 3 ? W[WPW + o * HID + k] : (o < 6 ? W[SCW + (o - 3) * HID + k] : 0.f);
    }
    if (tid < HID) {
        bs1[tid] = W[S1B + tid]; bs2[tid] = W[S2B + tid];
        bd1[tid] = W[D1B + tid]; bd2[tid] = W[D2B + tid]; bd3[tid] = W[D3B + tid];
        float b = W[D0B + tid];
        for (int j = 0; j < POSE; j++) b = fmaf(a.w_pose[tid * POSE + j], a.pose[j], b);
        bd0[tid] = b;
    }
    if (tid < 4) bs3[tid] = W[S3B + tid];
    if (tid < 8) bdh[tid] = tid < 3 ? W[WPB + tid] : (tid < 6 ? W[SCB + tid - 3] : 0.f);
    __syncthreads();

    float* col = act + tid;
    const int64_t ntiles = (a.N + P - 1) / P;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t p = tile * P + tid;
        const bool valid = p < a.N;
        // ---- opacity / colour net
        load_enc_column(a.enc, p, valid, col, P);
        hidden_fwd<ENC, ACT_RELU>(Ws1, bs1, col, a.acts_s ? a.acts_s + p : nullptr, a.Np, valid);
        hidden_fwd<HID, ACT_RELU>(Ws2, bs2, col, a.acts_s ? a.acts_s + HID * a.Np + p : nullptr, a.Np, valid);
        float o4[4];
        layer_fwd<HID, 4, ACT_NONE>(Ws3, bs3, col, P, o4);
        if (valid) {
            a.opac[p] = p < a.Nu ? sigmoidf(o4[0]) : 1.0f;           // mesh-bound Gaussians: opacity fixed to 1 (avatar.py:1287-1288)
            a.colors[p * 3] = sigmoidf(o4[1]); a.colors[p * 3 + 1] = sigmoidf(o4[2]); a.colors[p * 3 + 2] = sigmoidf(o4[3]);
        }
        // ---- deformation net (unconstrained Gaussians only)
        if (tile * P < a.Nu) {
            const bool vu = p < a.Nu;
            load_enc_column(a.enc, p, vu, col, P);
            hidden_fwd<ENC, ACT_LRELU>(Wd0, bd0, col, a.acts_d ? a.acts_d + p : nullptr, a.Nup, vu);
            hidden_fwd<HID, ACT_LRELU>(Wd1, bd1, col, a.acts_d ? a.acts_d + 1 * HID * a.Nup + p : nullptr, a.Nup, vu);
            hidden_fwd<HID, ACT_LRELU>(Wd2, bd2, col, a.acts_d ? a.acts_d + 2 * HID * a.Nup + p : nullptr, a.Nup, vu);
            hidden_fwd<HID, ACT_LRELU>(Wd3, bd3, col, a.acts_d ? a.acts_d + 3 * HID * a.Nup + p : nullptr, a.Nup, vu);
            float h8[8];
            layer_fwd<HID, 8, ACT_NONE>(Wdh, bdh, col, P, h8);
            if (vu) {
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    a.pos_out[p * 3 + c] = a.positions[p * 3 + c] + h8[c] * a.init_offset;
                    a.scales[p * 3 + c] = fminf(expf(h8[3 + c]) * a.init_scale, a.max_scale);
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------- backward
// cooperative load of a saved activation tile [rows][P] (global row stride Nrow) into A[rows][S]; points >= Nvalid -> 0
__device__ __forceinline__ void load_act_tile(float* A, const float* __restrict__ g, int rows, int64_t Nrow, int64_t p0, int64_t Nvalid) {
    for (int i = threadIdx.x; i < rows * (P / 4); i += P) {
        const int k = i / (P / 4), j = i - k * (P / 4);
        const int64_t p = p0 + 4 * j;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p + 3 < Nvalid) v = __ldg(reinterpret_cast<const float4*>(g + k * Nrow + p));
        else if (p < Nvalid) {
            v.x = g[k * Nrow + p];
            if (p + 1 < Nvalid) v.y = g[k * Nrow + p + 1];
            if (p + 2 < Nvalid) v.z = g[k * Nrow + p + 2];
        }
        *reinterpret_cast<float4*>(A + k * S + 4 * j) = v;
    }
}

__device__ __forceinline__ float dot4(const float4& a, const float4& b) { return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w))); }

// dW[o][i] += sum_p G[o][p] * A[i][p]   (O = 64 rows of G, KIN rows of A); thread (ty, tx) owns o = ty + 16a, i = tx + 16b
template <int KIN>
__device__ __forceinline__ void wgrad_tile(const float* __restrict__ G, const float* __restrict__ A, float (&acc)[4 * (KIN / 16)]) {
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    constexpr int NB = KIN / 16;
#pragma unroll 2
    for (int p4 = 0; p4 < P / 4; p4++) {
        float4 g[4], x[NB];
#pragma unroll
        for (int q = 0; q < 4; q++) g[q] = *reinterpret_cast<const float4*>(G + (ty + 16 * q) * S + 4 * p4);
#pragma unroll
        for (int b = 0; b < NB; b++) x[b] = *reinterpret_cast<const float4*>(A + (tx + 16 * b) * S + 4 * p4);
#pragma unroll
        for (int q = 0; q < 4; q++)
#pragma unroll
            for (int b = 0; b < NB; b++) acc[q * NB + b] += dot4(g[q], x[b]);
    }
}

// sum over the tile of row `row` of G
__device__ __forceinline__ float row_sum(const float* __restrict__ G, int row) {
    float s0 = 0.f, s1 = 0.f;
    const float4* r = reinterpret_cast<const float4*>(G + row * S);
#pragma unroll 4
    for (int j = 0; j < P / 4; j += 2) {
        const float4 u = r[j], v = r[j + 1];
        s0 += (u.x + u.y) + (u.z + u.w);
        s1 += (v.x + v.y) + (v.z + v.w);
    }
    return s0 + s1;
}

// thread-per-point input gradient, in place: G[i][p] <- (sum_{o<O} W[o][i] * G[o][p]) * act'(A[i][p]);  W = [O][KIN] row-major in smem
template <int O, int KIN, int ACT, bool W_ALIGNED>
__device__ __forceinline__ void dgrad_tile(const float* __restrict__ W, float* Gcol, const float* Acol, float (&gin)[KIN]) {
#pragma unroll
    for (int i = 0; i < KIN; i++) gin[i] = 0.f;
#pragma unroll 2
    for (int o = 0; o < O; o++) {
        const float g = Gcol[o * S];
        if (W_ALIGNED) {
            const float4* w = reinterpret_cast<const float4*>(W + o * KIN);
#pragma unroll
            for (int i4 = 0; i4 < KIN / 4; i4++) {
                const float4 ww = w[i4];
                gin[4 * i4] = fmaf(g, ww.x, gin[4 * i4]); gin[4 * i4 + 1] = fmaf(g, ww.y, gin[4 * i4 + 1]);
                gin[4 * i4 + 2] = fmaf(g, ww.z, gin[4 * i4 + 2]); gin[4 * i4 + 3] = fmaf(g, ww.w, gin[4 * i4 + 3]);
            }
        } else {
#pragma unroll
            for (int i = 0; i < KIN; i++) gin[i] = fmaf(g, W[o * KIN + i], gin[i]);
        }
    }
    if (ACT != ACT_NONE) {
#pragma unroll
        for (int i = 0; i < KIN; i++) gin[i] *= act_grad(Acol[i * S], ACT);
    }
}

__global__ void __launch_bounds__(P, 1)
mlp_bwd_kernel(const BwdArgs a) {
    extern __shared__ __align__(16) float sm[];
    float* Wsm = sm;                                   // flat parameter vector (original [out][in] layouts)
    float* G = Wsm + ((NPARAM + 3) & ~3);              // [64][S]
    float* A = G + HID * S;                            // [64][S]
    const int tid = threadIdx.x;
    for (int i = tid; i < NPARAM; i += P) Wsm[i] = a.params[i];
    __syncthreads();

    // register-resident weight-gradient accumulators (whole CTA lifetime)
    float dWs1[8] = {}, dWs2[16] = {}, dWs3 = 0.f, dWd0[8] = {}, dWd1[16] = {}, dWd2[16] = {}, dWd3[16] = {}, dWdh0 = 0.f, dWdh1 = 0.f;
    float dbs1 = 0.f, dbs2 = 0.f, dbs3 = 0.f, dbd0 = 0.f, dbd1 = 0.f, dbd2 = 0.f, dbd3 = 0.f, dbdh = 0.f;

    float* Gcol = G + tid;
    float* Acol = A + tid;
    const int64_t ntiles = (a.N + P - 1) / P;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t p0 = tile * P, p = p0 + tid;
        const bool valid = p < a.N;
        // =============================================================== opacity / colour net
        {
            float go[4] = {0.f, 0.f, 0.f, 0.f};
            if (valid) {
                if (a.g_opac && p < a.Nu) { const float o = a.opac[p]; go[0] = a.g_opac[p] * o * (1.f - o); }
                if (a.g_colors) {
#pragma unroll
                    for (int c = 0; c < 3; c++) { const float v = a.colors[p * 3 + c]; go[1 + c] = a.g_colors[p * 3 + c] * v * (1.f - v); }
                }
            }
            __syncthreads();                                       // previous tile is done with G / A
#pragma unroll
            for (int o = 0; o < 4; o++) Gcol[o * S] = go[o];
            load_act_tile(A, a.acts_s + HID * a.Np, HID, a.Np, p0, a.N);          // h2
            __syncthreads();
            {   // head weight gradient: 4 x 64 entries, one per thread
                const int o = tid >> 6, i = tid & 63;
                const float4* gr = reinterpret_cast<const float4*>(G + o * S);
                const float4* ar = reinterpret_cast<const float4*>(A + i * S);
                float s = 0.f;
#pragma unroll 4
                for (int j = 0; j < P / 4; j++) s += dot4(gr[j], ar[j]);
                dWs3 += s;
                if (tid < 4) dbs3 += row_sum(G, tid);
            }
            __syncthreads();
            float gin[HID];
            dgrad_tile<4, HID, ACT_RELU, true>(Wsm + S3W, Gcol, Acol, gin);
#pragma unroll
            for (int i = 0; i < HID; i++) Gcol[i * S] = gin[i];
            __syncthreads();
            load_act_tile(A, a.acts_s, HID, a.Np, p0, a.N);                        // h1
            __syncthreads();
            wgrad_tile<HID>(G, A, dWs2);
            if (tid < HID) dbs2 += row_sum(G, tid);
            __syncthreads();
            dgrad_tile<HID, HID, ACT_RELU, true>(Wsm + S2W, Gcol, Acol, gin);
#pragma unroll
            for (int i = 0; i < HID; i++) Gcol[i * S] = gin[i];
            __syncthreads();                                       // every thread is done with A (h1) before the columns are rewritten
            load_enc_column(a.enc, p, valid, Acol, S);
            __syncthreads();
            wgrad_tile<ENC>(G, A, dWs1);
            if (tid < HID) dbs1 += row_sum(G, tid);
            __syncthreads();
            float ge[ENC];
            dgrad_tile<HID, ENC, ACT_NONE, true>(Wsm + S1W, Gcol, Acol, ge);
            if (valid) {
                float4* dst = reinterpret_cast<float4*>(a.g_enc + p * ENC);
#pragma unroll
                for (int j = 0; j < ENC / 4; j++) dst[j] = make_float4(ge[4 * j], ge[4 * j + 1], ge[4 * j + 2], ge[4 * j + 3]);
            }
        }
        // =============================================================== deformation net
        if (p0 < a.Nu) {
            const bool vu = p < a.Nu;
            float gh[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if (vu) {
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    if (a.g_pos) gh[c] = a.g_pos[p * 3 + c] * a.init_offset;
                    if (a.g_scales) { const float s = a.scales[p * 3 + c]; gh[3 + c] = s < a.max_scale ? a.g_scales[p * 3 + c] * s : 0.f; }
                }
            }
            __syncthreads();
#pragma unroll
            for (int o = 0; o < 6; o++) Gcol[o * S] = gh[o];
            load_act_tile(A, a.acts_d + 3 * HID * a.Nup, HID, a.Nup, p0, a.Nu);     // d4
            __syncthreads();
            {   // head weight gradients: 6 x 64 entries: e = tid (o = tid/64 < 4) and e = tid + 256 (o = 4 + tid/64, tid < 128)
                const int i = tid & 63;
                const float4* ar = reinterpret_cast<const float4*>(A + i * S);
                const float4* g0 = reinterpret_cast<const float4*>(G + (tid >> 6) * S);
                const float4* g1 = reinterpret_cast<const float4*>(G + (4 + ((tid >> 6) & 1)) * S);
                float s0 = 0.f, s1 = 0.f;
#pragma unroll 4
                for (int j = 0; j < P / 4; j++) { const float4 x = ar[j]; s0 += dot4(g0[j], x); s1 += dot4(g1[j], x); }
                dWdh0 += s0;
                if (tid < 128) dWdh1 += s1;
                if (tid < 6) dbdh += row_sum(G, tid);
            }
            __syncthreads();
            float gin[HID];
            {   // heads -> d4: rows of the two head matrices (SCW is not 16-byte aligned: scalar loads)
#pragma unroll
                for (int i = 0; i < HID; i++) gin[i] = 0.f;
#pragma unroll
                for (int o = 0; o < 6; o++) {
                    const float g = Gcol[o * S];
                    const float* w = o < 3 ? Wsm + WPW + o * HID : Wsm + SCW + (o - 3) * HID;
#pragma unroll
                    for (int i = 0; i < HID; i++) gin[i] = fmaf(g, w[i], gin[i]);
                }
#pragma unroll
                for (int i = 0; i < HID; i++) gin[i] *= act_grad(Acol[i * S], ACT_LRELU);
            }
#pragma unroll
            for (int i = 0; i < HID; i++) Gcol[i * S] = gin[i];
#pragma unroll 1
            for (int l = 3; l >= 1; l--) {
                __syncthreads();
                load_act_tile(A, a.acts_d + (l - 1) * HID * a.Nup, HID, a.Nup, p0, a.Nu);
                __syncthreads();
                if (l == 3) { wgrad_tile<HID>(G, A, dWd3); if (tid < HID) dbd3 += row_sum(G, tid); }
                else if (l == 2) { wgrad_tile<HID>(G, A, dWd2); if (tid < HID) dbd2 += row_sum(G, tid); }
                else { wgrad_tile<HID>(G, A, dWd1); if (tid < HID) dbd1 += row_sum(G, tid); }
                __syncthreads();
                dgrad_tile<HID, HID, ACT_LRELU, true>(Wsm + (l == 3 ? D3W : (l == 2 ? D2W : D1W)), Gcol, Acol, gin);
#pragma unroll
                for (int i = 0; i < HID; i++) Gcol[i * S] = gin[i];
            }
            __syncthreads();
            load_enc_column(a.enc, p, vu, Acol, S);
            __syncthreads();
            wgrad_tile<ENC>(G, A, dWd0);
            if (tid < HID) dbd0 += row_sum(G, tid);
            __syncthreads();
            float ge[ENC];
            dgrad_tile<HID, ENC, ACT_NONE, true>(Wsm + D0W, Gcol, Acol, ge);
            if (vu) {
                float4* dst = reinterpret_cast<float4*>(a.g_enc + p * ENC);
#pragma unroll
                for (int j = 0; j < ENC / 4; j++) {
                    float4 v = dst[j];
                    v.x += ge[4 * j]; v.y += ge[4 * j + 1]; v.z += ge[4 * j + 2]; v.w += ge[4 * j + 3];
                    dst[j] = v;
                }
            }
        }
    }
    // ---- per-CTA partial gradient vector
    float* out = a.partial + (int64_t)blockIdx.x * NPARAM;
    const int tx = tid & 15, ty = tid >> 4;
#pragma unroll
    for (int q = 0; q < 4; q++) {
#pragma unroll
        for (int b = 0; b < 4; b++) {
            const int o = ty + 16 * q, i = tx + 16 * b;
            out[S2W + o * HID + i] = dWs2[q * 4 + b];
            out[D1W + o * HID + i] = dWd1[q * 4 + b];
            out[D2W + o * HID + i] = dWd2[q * 4 + b];
            out[D3W + o * HID + i] = dWd3[q * 4 + b];
        }
#pragma unroll
        for (int b = 0; b < 2; b++) {
            const int o = ty + 16 * q, i = tx + 16 * b;
            out[S1W + o * ENC + i] = dWs1[q * 2 + b];
            out[D0W + o * ENC + i] = dWd0[q * 2 + b];
        }
    }
    out[S3W + tid] = dWs3;                                          // [4][64] row-major == tid
    {
        const int o = tid >> 6, i = tid & 63;
        if (o < 3) out[WPW + o * HID + i] = dWdh0; else out[SCW + i] = dWdh0;
        if (tid < 128) out[SCW + (1 + (tid >> 6)) * HID + i] = dWdh1;
    }
    if (tid < HID) {
        out[S1B + tid] = dbs1; out[S2B + tid] = dbs2;
        out[D0B + tid] = dbd0; out[D1B + tid] = dbd1; out[D2B + tid] = dbd2; out[D3B + tid] = dbd3;
    }
    if (tid < 4) out[S3B + tid] = dbs3;
    if (tid < 6) { if (tid < 3) out[WPB + tid] = dbdh; else out[SCB + tid - 3] = dbdh; }
}

// g_params[j] = sum over CTAs of partial[c][j]; g_w_pose[o][j] = g_bias0[o] * pose[j]
__global__ void __launch_bounds__(256)
mlp_reduce_kernel(const float* __restrict__ partial, int nparts, const float* __restrict__ pose, float* __restrict__ g_params,
                  float* __restrict__ g_w_pose) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < NPARAM) {
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
        int c = 0;
        for (; c + 3 < nparts; c += 4) {
            s0 += partial[(int64_t)c * NPARAM + j]; s1 += partial[(int64_t)(c + 1) * NPARAM + j];
            s2 += partial[(int64_t)(c + 2) * NPARAM + j]; s3 += partial[(int64_t)(c + 3) * NPARAM + j];
        }
        for (; c < nparts; c++) s0 += partial[(int64_t)c * NPARAM + j];
        const float s = (s0 + s1) + (s2 + s3);
        g_params[j] = s;
        if (g_w_pose && j >= D0B && j < D0B + HID) {
            const int o = j - D0B;
            for (int k = 0; k < POSE; k++) g_w_pose[o * POSE + k] = s * pose[k];
        }
    }
}

// =====================================================================================================================
// Tensor-core forward (tcgen05, sm_100a).  The hidden width is 64: one layer of a 128-point tile is a 128 x 64 x 64 GEMM.
//   * fp32 accuracy on fp16 tensor cores: every operand is split x = hi + lo (two fp16 values, 22 mantissa bits) and a
//     layer is the three products  A_hi W_hi + A_hi W_lo + A_lo W_hi  accumulated in fp32 in TMEM (the dropped lo x lo
//     term is 2^-22 relative): 12 MMAs (M128 N64 K16) per 64-wide layer.  The 32 grid features (|x| can be ~1e-4 right
//     after initialisation, where an fp16 lo part would be subnormal) are scaled by 2^10 before the split and the
//     accumulator by 2^-10 -- both exact.
//   * all weights of both MLPs live in shared memory for the whole kernel as 128B-swizzled K-major fp16 tiles (hi + lo,
//     104 KB); a thread owns one point (= one TMEM lane): it reads its accumulator row with tcgen05.ld, applies bias +
//     activation, stores the fp32 activations for the backward (same [layer][feature][point] layout as the SIMT kernel)
//     and writes the split row straight into the swizzled A-operand tile of the next layer.
//   * two independent 128-point tiles per CTA (warps 0-3 / 4-7): while one tile is
//     in its epilogue the tensor pipe works on the other.  No dedicated issuing warp: the first thread of a group issues
//     its MMAs once all 128 rows of the A tile have arrived (256 threads per CTA: 255 registers per thread, no spills).
constexpr int PT = 128;
constexpr float kEncScale = 1024.0f, kEncInv = 1.0f / 1024.0f;
constexpr uint32_t TC_TILE64 = 64 * 128, TC_TILE16 = 16 * 128, TC_ATILE = PT * 128;
// byte offsets of the weight tiles inside one precision half
constexpr uint32_t TW_S1 = 0, TW_S2 = TW_S1 + TC_TILE64, TW_S3 = TW_S2 + TC_TILE64, TW_D0 = TW_S3 + TC_TILE16, TW_D1 = TW_D0 + TC_TILE64,
                   TW_D2 = TW_D1 + TC_TILE64, TW_D3 = TW_D2 + TC_TILE64, TW_DH = TW_D3 + TC_TILE64, TW_HALF = TW_DH + TC_TILE16;
constexpr size_t kTcFwdSmem = 1024 + 2 * (size_t)TW_HALF + 2 * 2 * (size_t)TC_ATILE + sizeof(float) * (6 * HID + 32) + 64;

__device__ __forceinline__ uint32_t tc_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tc_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(tc_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void tc_mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(tc_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(tc_smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(tc_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
                 "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int j = 0; j < 32; j++) v[j] = __uint_as_float(r[j]);
}
// K-major, 128B-swizzled operand tile [rows][64 fp16]: 8-row groups 1024 B apart (same descriptor as csrc/gemm_tcgen05.cu)
__device__ __forceinline__ uint64_t tc_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ uint32_t tc_idesc(int N) { return (1u << 4) | kIdescFmtAB | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(PT >> 4) << 24); }

// byte offset of element (row, k) inside a swizzled tile
__device__ __forceinline__ uint32_t tc_off(int row, int k) {
    return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((((k >> 3) ^ (row & 7)) & 7) << 4) + (k & 7) * 2);
}
// W [O][K] fp32 row-major -> hi / lo tiles with `rows` rows (rows >= O, zero padded; K <= 64, zero padded)
__device__ __forceinline__ void tc_fill_weight(uint8_t* hi, uint8_t* lo, const float* __restrict__ W, int O, int K, int rows, int tid, int nthr) {
    for (int i = tid; i < rows * 64; i += nthr) {
        const int o = i >> 6, k = i & 63;
        const float w = (o < O && k < K) ? W[o * K + k] : 0.f;
        const __half h = __float2half_rn(w);
        const __half l = __float2half_rn(w - __half2float(h));
        const uint32_t off = tc_off(o, k);
        *reinterpret_cast<__half*>(hi + off) = h;
        *reinterpret_cast<__half*>(lo + off) = l;
    }
}
// one thread's row of NV values (NV = 32 or 64) -> split hi / lo -> its 128-byte row of the swizzled A tiles
template <int NV>
__device__ __forceinline__ void tc_store_row(uint8_t* Ahi, uint8_t* Alo, int r, const float (&v)[64]) {
    uint8_t* rh = Ahi + (r >> 3) * 1024 + (r & 7) * 128;
    uint8_t* rl = Alo + (r >> 3) * 1024 + (r & 7) * 128;
#pragma unroll
    for (int ch = 0; ch < NV / 8; ch++) {
        uint32_t ph[4], pl[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const float x0 = v[ch * 8 + 2 * i], x1 = v[ch * 8 + 2 * i + 1];
            ph[i] = pack_act2(x0, x1);
            const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&ph[i]));
            pl[i] = pack_act2(x0 - hf.x, x1 - hf.y);
        }
        const int sw = ((ch ^ (r & 7)) & 7) << 4;
        *reinterpret_cast<uint4*>(rh + sw) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
        *reinterpret_cast<uint4*>(rl + sw) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
    }
}
// the issuing lane: D[tmem] = A_hi W_hi + A_hi W_lo + A_lo W_hi over `ksteps` K = 16 steps, N output columns
__device__ __forceinline__ void tc_issue_layer(uint32_t tmem_d, uint32_t a_hi, uint32_t a_lo, uint32_t w_hi, uint32_t w_lo, int ksteps, int N) {
    const uint32_t idesc = tc_idesc(N);
    const uint64_t dah = tc_desc(a_hi), dal = tc_desc(a_lo), dwh = tc_desc(w_hi), dwl = tc_desc(w_lo);
    uint32_t acc = 0;
    for (int ks = 0; ks < ksteps; ks++) { tc_mma(tmem_d, dah + 2u * ks, dwh + 2u * ks, idesc, acc); acc = 1; }
    for (int ks = 0; ks < ksteps; ks++) tc_mma(tmem_d, dah + 2u * ks, dwl + 2u * ks, idesc, 1);
    for (int ks = 0; ks < ksteps; ks++) tc_mma(tmem_d, dal + 2u * ks, dwh + 2u * ks, idesc, 1);
}

__global__ void __launch_bounds__(256, 1)
mlp_fwd_tc_kernel(const FwdArgs a) {
    extern __shared__ uint8_t tc_smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tc_smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* Whi = smem;
    uint8_t* Wlo = Whi + TW_HALF;
    uint8_t* Abase = Wlo + TW_HALF;                               // group g: hi at g * 2 * TC_ATILE, lo right behind
    float* bias = reinterpret_cast<float*>(Abase + 4 * TC_ATILE);  // bs1, bs2, bd0 (effective), bd1, bd2, bd3 [64 each], bs3 [16], bdh [16]
    uint64_t* bars = reinterpret_cast<uint64_t*>(bias + 6 * HID + 32);
    uint64_t* acc_full = bars;                                     // [2]
    uint64_t* a_ready = bars + 2;                                  // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const float* W = a.params;

    tc_fill_weight(Whi + TW_S1, Wlo + TW_S1, W + S1W, HID, ENC, 64, tid, 256);
    tc_fill_weight(Whi + TW_S2, Wlo + TW_S2, W + S2W, HID, HID, 64, tid, 256);
    tc_fill_weight(Whi + TW_S3, Wlo + TW_S3, W + S3W, 4, HID, 16, tid, 256);
    tc_fill_weight(Whi + TW_D0, Wlo + TW_D0, W + D0W, HID, ENC, 64, tid, 256);
    tc_fill_weight(Whi + TW_D1, Wlo + TW_D1, W + D1W, HID, HID, 64, tid, 256);
    tc_fill_weight(Whi + TW_D2, Wlo + TW_D2, W + D2W, HID, HID, 64, tid, 256);
    tc_fill_weight(Whi + TW_D3, Wlo + TW_D3, W + D3W, HID, HID, 64, tid, 256);
    tc_fill_weight(Whi + TW_DH, Wlo + TW_DH, W + WPW, 3, HID, 3, tid, 256);                  // rows 0-2: warp head
    for (int i = tid; i < 13 * 64; i += 256) {                                               // rows 3-5: scaling head, 6-15: zero
        const int o = 3 + (i >> 6), k = i & 63;
        const float w = o < 6 ? W[SCW + (o - 3) * HID + k] : 0.f;
        const __half h = __float2half_rn(w);
        const uint32_t off = tc_off(o, k);
        *reinterpret_cast<__half*>(Whi + TW_DH + off) = h;
        *reinterpret_cast<__half*>(Wlo + TW_DH + off) = __float2half_rn(w - __half2float(h));
    }
    if (tid < HID) {
        bias[tid] = W[S1B + tid]; bias[HID + tid] = W[S2B + tid];
        float b = W[D0B + tid];
        for (int j = 0; j < POSE; j++) b = fmaf(a.w_pose[tid * POSE + j], a.pose[j], b);
        bias[2 * HID + tid] = b;
        bias[3 * HID + tid] = W[D1B + tid]; bias[4 * HID + tid] = W[D2B + tid]; bias[5 * HID + tid] = W[D3B + tid];
    }
    if (tid < 16) {
        bias[6 * HID + tid] = tid < 4 ? W[S3B + tid] : 0.f;
        bias[6 * HID + 16 + tid] = tid < 3 ? W[WPB + tid] : (tid < 6 ? W[SCB + tid - 3] : 0.f);
    }
    if (tid == 0) {
        tc_mbar_init(&acc_full[0], 1); tc_mbar_init(&acc_full[1], 1);
        tc_mbar_init(&a_ready[0], PT); tc_mbar_init(&a_ready[1], PT);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(tc_smem_u32(tmem_slot)), "n"(128));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");    // the weight tiles were written through the generic proxy
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int64_t ntiles = (a.N + PT - 1) / PT;
    {
        // ---------------- point threads: group g = warp / 4, TMEM lane = row = 32 * (warp % 4) + lane
        const int g = warp >> 2, r = ((warp & 3) << 5) + lane;
        uint8_t* Ahi = Abase + g * 2 * TC_ATILE;
        uint8_t* Alo = Ahi + TC_ATILE;
        const uint32_t tm = tmem_base + (uint32_t)(g * 64) + ((uint32_t)((warp & 3) * 32) << 16);
        uint32_t par = 0;
        float v[64];
        // publish the A tile of layer l; the first thread of the group is also its issuing thread: once all 128 rows have
        // arrived it issues the layer's MMAs and commits them to acc_full (every thread then waits on that barrier)
        const uint32_t a_hi_s = tc_smem_u32(Ahi), a_lo_s = tc_smem_u32(Alo), w_hi_s = tc_smem_u32(Whi), w_lo_s = tc_smem_u32(Wlo);
        const uint32_t tm_acc = tmem_base + (uint32_t)(g * 64);
        uint32_t par_a = 0;
        auto publish = [&](int l) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            tc_fence_before();
            tc_mbar_arrive(&a_ready[g]);
            if (r == 0) {
                tc_mbar_wait(&a_ready[g], par_a);
                par_a ^= 1u;
                tc_fence_after();
                uint32_t off; int ks, N;
                switch (l) {
                    case 0: off = TW_S1; ks = 2; N = 64; break;
                    case 1: off = TW_S2; ks = 4; N = 64; break;
                    case 2: off = TW_S3; ks = 4; N = 16; break;
                    case 3: off = TW_D0; ks = 2; N = 64; break;
                    case 4: off = TW_D1; ks = 4; N = 64; break;
                    case 5: off = TW_D2; ks = 4; N = 64; break;
                    case 6: off = TW_D3; ks = 4; N = 64; break;
                    default: off = TW_DH; ks = 4; N = 16; break;
                }
                tc_issue_layer(tm_acc, a_hi_s, a_lo_s, w_hi_s + off, w_lo_s + off, ks, N);
                tc_commit(&acc_full[g]);
            }
        };
        auto load_enc = [&](int64_t p, bool ok) {
            if (ok) {
                const float4* e = reinterpret_cast<const float4*>(a.enc + p * ENC);
#pragma unroll
                for (int j = 0; j < ENC / 4; j++) {
                    const float4 t = __ldg(e + j);
                    v[4 * j] = t.x * kEncScale; v[4 * j + 1] = t.y * kEncScale; v[4 * j + 2] = t.z * kEncScale; v[4 * j + 3] = t.w * kEncScale;
                }
            } else {
#pragma unroll
                for (int j = 0; j < ENC; j++) v[j] = 0.f;
            }
            tc_store_row<ENC>(Ahi, Alo, r, v);
        };
        auto wait_acc = [&](bool both) {
            tc_mbar_wait(&acc_full[g], par);
            par ^= 1u;
            tc_fence_after();
            float t0[32];
            tc_ld32(tm, t0);
#pragma unroll
            for (int j = 0; j < 32; j++) v[j] = t0[j];
            if (both) {
                tc_ld32(tm + 32u, t0);
#pragma unroll
                for (int j = 0; j < 32; j++) v[32 + j] = t0[j];
            }
        };
        // hidden layer epilogue: v <- act(v * scale + b); save; write the next A tile
        // (the activation stores for the backward are issued AFTER the A tile is published: the proxy fence would otherwise
        //  wait for them, and they overlap the next layer's MMAs instead)
        auto hidden = [&](const float* b, float scale, int act, float* gsave, int64_t Nrow, bool ok, int next_layer) {
#pragma unroll
            for (int o = 0; o < HID; o++) v[o] = act_fn(fmaf(v[o], scale, b[o]), act);
            tc_store_row<HID>(Ahi, Alo, r, v);
            publish(next_layer);
            if (gsave && ok) {
#pragma unroll
                for (int o = 0; o < HID; o++) gsave[o * Nrow] = v[o];
            }
        };
        for (int64_t it = blockIdx.x; 2 * it + g < ntiles; it += gridDim.x) {
            const int64_t p0 = (2 * it + g) * PT, p = p0 + r;
            const bool valid = p < a.N;
            // saved activations: tile-blocked [layer][tile][feature][128 points] (one contiguous 32 KB block per tile and layer)
            const int64_t tile_off = (2 * it + g) * (int64_t)(HID * PT) + r;
            const int64_t ngs = (a.N + PT - 1) / PT * PT, ngd = (a.Nu + PT - 1) / PT * PT;
            load_enc(p, valid);
            publish(0);
            wait_acc(true);
            hidden(bias, kEncInv, ACT_RELU, a.acts_s ? a.acts_s + tile_off : nullptr, PT, valid, 1);
            wait_acc(true);
            hidden(bias + HID, 1.0f, ACT_RELU, a.acts_s ? a.acts_s + HID * ngs + tile_off : nullptr, PT, valid, 2);
            wait_acc(false);
            if (valid) {
                const float* b3 = bias + 6 * HID;
                a.opac[p] = p < a.Nu ? sigmoidf(v[0] + b3[0]) : 1.0f;      // mesh-bound Gaussians: opacity fixed to 1 (avatar.py:1287-1288)
                a.colors[p * 3] = sigmoidf(v[1] + b3[1]); a.colors[p * 3 + 1] = sigmoidf(v[2] + b3[2]); a.colors[p * 3 + 2] = sigmoidf(v[3] + b3[3]);
            }
            if (p0 < a.Nu) {
                const bool vu = p < a.Nu;
                load_enc(p, vu);
                publish(3);
                wait_acc(true);
                hidden(bias + 2 * HID, kEncInv, ACT_LRELU, a.acts_d ? a.acts_d + tile_off : nullptr, PT, vu, 4);
#pragma unroll 1
                for (int l = 1; l <= 3; l++) {
                    wait_acc(true);
                    hidden(bias + (2 + l) * HID, 1.0f, ACT_LRELU, a.acts_d ? a.acts_d + (int64_t)l * HID * ngd + tile_off : nullptr, PT, vu, 4 + l);
                }
                wait_acc(false);
                if (vu) {
                    const float* bh = bias + 6 * HID + 16;
#pragma unroll
                    for (int c = 0; c < 3; c++) {
                        a.pos_out[p * 3 + c] = a.positions[p * 3 + c] + (v[c] + bh[c]) * a.init_offset;
                        a.scales[p * 3 + c] = fminf(expf(v[3 + c] + bh[3 + c]) * a.init_scale, a.max_scale);
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "n"(128));
}

// =====================================================================================================================
// Tensor-core backward, two kernels.
//   K1  mlp_bwd_tc_dgrad_kernel: the input-gradient chain, same structure as the forward (thread = point = TMEM lane,
//       two tiles in flight): G_l = dL/dH_l * act'(H_l) is built per point, written (a) as the swizzled A operand of
//       dL/dH_{l-1} = G_l W_l (B operand = W_l^T tiles resident in shared memory) and (b) to global memory as packed
//       (hi, lo) fp16 pairs in [feature][point] layout for K2.  Ends with dL/denc.
//   K2  mlp_bwd_tc_wgrad_kernel: dW_l = G_l^T X_{l-1}, a reduction over POINTS: both operands are K-major exactly as they
//       lie in memory ([feature][point]); eight producer warps convert / split them into double-buffered swizzled tiles,
//       one lane issues the MMAs, and every weight matrix keeps ONE accumulator in TMEM for the whole kernel (496 of the
//       512 columns).  A row of ones appended to X^T yields the bias gradients in the same MMA; the two layers fed by
//       the grid features (sigma layer 1, deform layer 0) share one 128-row A tile.  Per-CTA partials + the existing
//       ordered reduction kernel: deterministic.
constexpr int GR_S3 = 0, GR_S2 = 4, GR_S1 = 68, GR_DH = 132, GR_D3 = 138, GR_D2 = 202, GR_D1 = 266, GR_D0 = 330, GR_ROWS = 394;
// W^T tiles ([in rows][out = K]) inside one precision half
constexpr uint32_t TT_S3 = 0, TT_S2 = TT_S3 + TC_TILE64, TT_S1 = TT_S2 + TC_TILE64, TT_DH = TT_S1 + 32 * 128, TT_D3 = TT_DH + TC_TILE64,
                   TT_D2 = TT_D3 + TC_TILE64, TT_D1 = TT_D2 + TC_TILE64, TT_D0 = TT_D1 + TC_TILE64, TT_HALF = TT_D0 + 32 * 128;
constexpr size_t kTcDgradSmem = 1024 + 2 * (size_t)TT_HALF + 2 * 2 * (size_t)TC_ATILE + sizeof(float) * (256 * 10) + 64;

struct BwdTcArgs {
    BwdArgs b;
    uint32_t* G;              // [GR_ROWS][Ng] packed half2 (hi, lo)
    float* partial1;          // [gridDim.x][16]: db of the two head layers (S3: 0-3, warp: 4-6, scaling: 7-9)
    int64_t Ng;
};

__device__ __forceinline__ uint32_t tc_split_pack(float x) {           // (hi, lo) of one value as half2
    const __half h = __float2half_rn(fminf(fmaxf(x, -65504.f), 65504.f));
    const __half l = __float2half_rn(x - __half2float(h));
    return (uint32_t)__half_as_ushort(h) | ((uint32_t)__half_as_ushort(l) << 16);
}
// B tile of W^T: rows = input feature i (rows_pad rows), K = output o: element (i, o) = W[o][i]
__device__ __forceinline__ void tc_fill_weight_t(uint8_t* hi, uint8_t* lo, const float* __restrict__ W, int O, int K, int rows_pad, int o_off,
                                                  int tid, int nthr) {
    for (int idx = tid; idx < rows_pad * 64; idx += nthr) {
        const int i = idx >> 6, o = idx & 63;
        const float w = (i < K && o >= o_off && o < o_off + O) ? W[(o - o_off) * K + i] : 0.f;
        const __half h = __float2half_rn(w);
        const uint32_t off = tc_off(i, o);
        *reinterpret_cast<__half*>(hi + off) = h;
        *reinterpret_cast<__half*>(lo + off) = __float2half_rn(w - __half2float(h));
    }
}

__global__ void __launch_bounds__(256, 1)
mlp_bwd_tc_dgrad_kernel(const BwdTcArgs t) {
    const BwdArgs& a = t.b;
    extern __shared__ uint8_t tc_smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tc_smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* Whi = smem;
    uint8_t* Wlo = Whi + TT_HALF;
    uint8_t* Abase = Wlo + TT_HALF;
    float* s_red = reinterpret_cast<float*>(Abase + 4 * TC_ATILE);   // [256][10]
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_red + 256 * 10);
    uint64_t* acc_full = bars;
    uint64_t* a_ready = bars + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const float* W = a.params;
    tc_fill_weight_t(Whi + TT_S3, Wlo + TT_S3, W + S3W, 4, HID, 64, 0, tid, 256);
    tc_fill_weight_t(Whi + TT_S2, Wlo + TT_S2, W + S2W, HID, HID, 64, 0, tid, 256);
    tc_fill_weight_t(Whi + TT_S1, Wlo + TT_S1, W + S1W, HID, ENC, 32, 0, tid, 256);
    // head tile: K columns 0-2 = warp head, 3-5 = scaling head
    for (int idx = tid; idx < 64 * 64; idx += 256) {
        const int i = idx >> 6, o = idx & 63;
        const float w = o < 3 ? W[WPW + o * HID + i] : (o < 6 ? W[SCW + (o - 3) * HID + i] : 0.f);
        const __half h = __float2half_rn(w);
        const uint32_t off = tc_off(i, o);
        *reinterpret_cast<__half*>(Whi + TT_DH + off) = h;
        *reinterpret_cast<__half*>(Wlo + TT_DH + off) = __float2half_rn(w - __half2float(h));
    }
    tc_fill_weight_t(Whi + TT_D3, Wlo + TT_D3, W + D3W, HID, HID, 64, 0, tid, 256);
    tc_fill_weight_t(Whi + TT_D2, Wlo + TT_D2, W + D2W, HID, HID, 64, 0, tid, 256);
    tc_fill_weight_t(Whi + TT_D1, Wlo + TT_D1, W + D1W, HID, HID, 64, 0, tid, 256);
    tc_fill_weight_t(Whi + TT_D0, Wlo + TT_D0, W + D0W, HID, ENC, 32, 0, tid, 256);
    if (tid == 0) {
        tc_mbar_init(&acc_full[0], 1); tc_mbar_init(&acc_full[1], 1);
        tc_mbar_init(&a_ready[0], PT); tc_mbar_init(&a_ready[1], PT);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(tc_smem_u32(tmem_slot)), "n"(128));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int64_t ntiles = (a.N + PT - 1) / PT;
    float bsum[10];
#pragma unroll
    for (int i = 0; i < 10; i++) bsum[i] = 0.f;

    {
        const int g = warp >> 2, r = ((warp & 3) << 5) + lane;
        uint8_t* Ahi = Abase + g * 2 * TC_ATILE;
        uint8_t* Alo = Ahi + TC_ATILE;
        const uint32_t tm = tmem_base + (uint32_t)(g * 64) + ((uint32_t)((warp & 3) * 32) << 16);
        uint32_t par = 0;
        float v[64];
        // publish the A tile of layer l; the first thread of the group is also its issuing thread: once all 128 rows have
        // arrived it issues the layer's MMAs and commits them to acc_full (every thread then waits on that barrier)
        const uint32_t a_hi_s = tc_smem_u32(Ahi), a_lo_s = tc_smem_u32(Alo), w_hi_s = tc_smem_u32(Whi), w_lo_s = tc_smem_u32(Wlo);
        const uint32_t tm_acc = tmem_base + (uint32_t)(g * 64);
        uint32_t par_a = 0;
        auto publish = [&](int l) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            tc_fence_before();
            tc_mbar_arrive(&a_ready[g]);
            if (r == 0) {
                tc_mbar_wait(&a_ready[g], par_a);
                par_a ^= 1u;
                tc_fence_after();
                uint32_t off; int ks, N;
                switch (l) {
                    case 0: off = TT_S3; ks = 1; N = 64; break;
                    case 1: off = TT_S2; ks = 4; N = 64; break;
                    case 2: off = TT_S1; ks = 4; N = 32; break;
                    case 3: off = TT_DH; ks = 1; N = 64; break;
                    case 4: off = TT_D3; ks = 4; N = 64; break;
                    case 5: off = TT_D2; ks = 4; N = 64; break;
                    case 6: off = TT_D1; ks = 4; N = 64; break;
                    default: off = TT_D0; ks = 4; N = 32; break;
                }
                tc_issue_layer(tm_acc, a_hi_s, a_lo_s, w_hi_s + off, w_lo_s + off, ks, N);
                tc_commit(&acc_full[g]);
            }
        };
        auto wait_acc = [&](bool both) {
            tc_mbar_wait(&acc_full[g], par);
            par ^= 1u;
            tc_fence_after();
            float t0[32];
            tc_ld32(tm, t0);
#pragma unroll
            for (int j = 0; j < 32; j++) v[j] = t0[j];
            if (both) {
                tc_ld32(tm + 32u, t0);
#pragma unroll
                for (int j = 0; j < 32; j++) v[32 + j] = t0[j];
            }
        };
        // v (dL/dH of a hidden layer) -> G = v * act'(H).  The 64 saved activations of the point are fetched BEFORE the wait
        // for the accumulator (they do not depend on it), the A tile is published, and only then G goes to global memory
        // (packed hi / lo, tile-blocked [tile][row][128 points]) for the weight-gradient kernel.
        float hbuf[HID];
        auto fetch_h = [&](const float* __restrict__ H, bool ok) {
#pragma unroll
            for (int i = 0; i < HID; i++) hbuf[i] = ok ? __ldg(H + i * PT) : 0.f;
        };
        auto hidden_bwd = [&](uint32_t* gp, bool ok, int act, int next_layer) {
#pragma unroll
            for (int i = 0; i < HID; i++) v[i] = ok ? v[i] * act_grad(hbuf[i], act) : 0.f;
            tc_store_row<HID>(Ahi, Alo, r, v);
            publish(next_layer);
#pragma unroll
            for (int i = 0; i < HID; i++) gp[i * PT] = tc_split_pack(v[i]);
        };
        for (int64_t it = blockIdx.x; 2 * it + g < ntiles; it += gridDim.x) {
            const int64_t p0 = (2 * it + g) * PT, p = p0 + r;
            const bool valid = p < a.N;
            const int64_t tile_off = (2 * it + g) * (int64_t)(HID * PT) + r;
            const int64_t ngs = (a.N + PT - 1) / PT * PT, ngd = (a.Nu + PT - 1) / PT * PT;
            uint32_t* Gt = t.G + (2 * it + g) * (int64_t)(GR_ROWS * PT) + r;       // this tile's block of G, this point's column
            // ---- opacity / colour net
#pragma unroll
            for (int i = 0; i < 16; i++) v[i] = 0.f;
            if (valid) {
                if (a.g_opac && p < a.Nu) { const float o = a.opac[p]; v[0] = a.g_opac[p] * o * (1.f - o); }
                if (a.g_colors) {
#pragma unroll
                    for (int c = 0; c < 3; c++) { const float x = a.colors[p * 3 + c]; v[1 + c] = a.g_colors[p * 3 + c] * x * (1.f - x); }
                }
            }
#pragma unroll
            for (int o = 0; o < 4; o++) { Gt[(GR_S3 + o) * PT] = tc_split_pack(v[o]); bsum[o] += v[o]; }
            tc_store_row<16>(Ahi, Alo, r, v);
            publish(0);
            fetch_h(a.acts_s + HID * ngs + tile_off, valid);
            wait_acc(true);
            hidden_bwd(Gt + GR_S2 * PT, valid, ACT_RELU, 1);
            fetch_h(a.acts_s + tile_off, valid);
            wait_acc(true);
            hidden_bwd(Gt + GR_S1 * PT, valid, ACT_RELU, 2);
            wait_acc(false);                                          // dL/denc of the colour net (32 columns)
            if (valid) {
                float4* dst = reinterpret_cast<float4*>(a.g_enc + p * ENC);
#pragma unroll
                for (int j = 0; j < ENC / 4; j++) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            }
            // ---- deformation net
            if (p0 < a.Nu) {
                const bool vu = p < a.Nu;
#pragma unroll
                for (int i = 0; i < 16; i++) v[i] = 0.f;
                if (vu) {
#pragma unroll
                    for (int c = 0; c < 3; c++) {
                        if (a.g_pos) v[c] = a.g_pos[p * 3 + c] * a.init_offset;
                        if (a.g_scales) { const float sc = a.scales[p * 3 + c]; v[3 + c] = sc < a.max_scale ? a.g_scales[p * 3 + c] * sc : 0.f; }
                    }
                }
#pragma unroll
                for (int o = 0; o < 6; o++) { Gt[(GR_DH + o) * PT] = tc_split_pack(v[o]); bsum[4 + o] += v[o]; }
                tc_store_row<16>(Ahi, Alo, r, v);
                publish(3);
                fetch_h(a.acts_d + 3 * HID * ngd + tile_off, vu);
                wait_acc(true);
                hidden_bwd(Gt + GR_D3 * PT, vu, ACT_LRELU, 4);
                fetch_h(a.acts_d + 2 * HID * ngd + tile_off, vu);
                wait_acc(true);
                hidden_bwd(Gt + GR_D2 * PT, vu, ACT_LRELU, 5);
                fetch_h(a.acts_d + 1 * HID * ngd + tile_off, vu);
                wait_acc(true);
                hidden_bwd(Gt + GR_D1 * PT, vu, ACT_LRELU, 6);
                fetch_h(a.acts_d + tile_off, vu);
                wait_acc(true);
                hidden_bwd(Gt + GR_D0 * PT, vu, ACT_LRELU, 7);
                wait_acc(false);
                if (vu) {                                             // same thread wrote these 128 bytes a few layers ago
                    float4* dst = reinterpret_cast<float4*>(a.g_enc + p * ENC);
#pragma unroll
                    for (int j = 0; j < ENC / 4; j++) {
                        float4 q = dst[j];
                        q.x += v[4 * j]; q.y += v[4 * j + 1]; q.z += v[4 * j + 2]; q.w += v[4 * j + 3];
                        dst[j] = q;
                    }
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 10; i++) s_red[tid * 10 + i] = bsum[i];
    }
    tc_fence_before();
    __syncthreads();
    if (tid < 10) {                                                   // fixed-order sum over the 256 point threads
        float sacc = 0.f;
        for (int k = 0; k < 256; k++) sacc += s_red[k * 10 + tid];
        t.partial1[blockIdx.x * 16 + tid] = sacc;
    }
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "n"(128));
}

// ---- K2: weight gradients.  Work item = (tile of 128 points, layer group); stage = 104 KB of operand tiles.
constexpr uint32_t WG_AATOM = 128 * 128, WG_BATOM = 80 * 128;                 // [128 rows][64 pts], [80 rows][64 pts]
constexpr uint32_t WG_A = 2 * WG_AATOM, WG_B = 2 * WG_BATOM;                   // one precision half (two 64-point atoms)
constexpr uint32_t WG_STAGE = 2 * WG_A + 2 * WG_B;                             // hi + lo of both operands = 106496 B
constexpr size_t kTcWgradSmem = 1024 + 2 * (size_t)WG_STAGE + 128;
// TMEM columns of the seven accumulators
constexpr int WC_L0 = 0, WC_S2 = 48, WC_S3 = 128, WC_D1 = 192, WC_D2 = 272, WC_D3 = 352, WC_DH = 432;

struct WgItem { int a_row0, a_rows, a2_row0, b_kind, b_layer, n, col; };
// b_kind: 0 = enc^T (32 rows + ones), 1 = acts_s[b_layer] (64 rows + ones), 2 = acts_d[b_layer] (+ ones), 3 / 4 = the same without the ones row
__device__ __forceinline__ WgItem wg_item(int l) {
    switch (l) {
        case 0: return {GR_S1, 64, GR_D0, 0, 0, 48, WC_L0};
        case 1: return {GR_S2, 64, -1, 1, 0, 80, WC_S2};
        case 2: return {GR_S3, 4, -1, 3, 1, 64, WC_S3};
        case 3: return {GR_D1, 64, -1, 2, 0, 80, WC_D1};
        case 4: return {GR_D2, 64, -1, 2, 1, 80, WC_D2};
        case 5: return {GR_D3, 64, -1, 2, 2, 80, WC_D3};
        default: return {GR_DH, 6, -1, 4, 3, 64, WC_DH};
    }
}

__global__ void __launch_bounds__(288, 1)
mlp_bwd_tc_wgrad_kernel(const BwdTcArgs t, float* __restrict__ partial) {
    const BwdArgs& a = t.b;
    extern __shared__ uint8_t tc_smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tc_smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * WG_STAGE);
    uint64_t* full = bars;            // [2] producers -> issuer
    uint64_t* empty = bars + 2;       // [2] tensor pipe -> producers
    uint64_t* done = bars + 4;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);
    int* s_used = reinterpret_cast<int*>(bars + 6);                 // [0]: this CTA had a tile, [1]: ... a deform tile
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        tc_mbar_init(&full[0], 256); tc_mbar_init(&full[1], 256);
        tc_mbar_init(&empty[0], 1); tc_mbar_init(&empty[1], 1);
        tc_mbar_init(done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        s_used[0] = 0; s_used[1] = 0;
    }
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(tc_smem_u32(tmem_slot)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int64_t ntiles = (a.N + PT - 1) / PT;

    if (warp == 8) {
        if (lane == 0) {
            uint32_t seen = 0;                                      // bit l: accumulator l already holds data
            uint32_t n_item = 0;
            for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                const int nl = tile * PT < a.Nu ? 7 : 3;
                for (int l = 0; l < nl; l++, n_item++) {
                    const uint32_t st = n_item & 1u;
                    tc_mbar_wait(&full[st], (n_item >> 1) & 1u);
                    tc_fence_after();
                    const WgItem w = wg_item(l);
                    const uint32_t base = tc_smem_u32(smem + st * WG_STAGE);
                    const uint32_t idesc = tc_idesc(w.n);
                    const uint32_t d = tmem_base + (uint32_t)w.col;
                    uint32_t acc = (seen >> l) & 1u;
                    for (int at = 0; at < 2; at++) {
                        const uint64_t ah = tc_desc(base + at * WG_AATOM), al = tc_desc(base + WG_A + at * WG_AATOM);
                        const uint64_t bh = tc_desc(base + 2 * WG_A + at * WG_BATOM), bl = tc_desc(base + 2 * WG_A + WG_B + at * WG_BATOM);
                        for (int ks = 0; ks < 4; ks++) { tc_mma(d, ah + 2u * ks, bh + 2u * ks, idesc, acc); acc = 1; }
                        for (int ks = 0; ks < 4; ks++) tc_mma(d, ah + 2u * ks, bl + 2u * ks, idesc, 1);
                        for (int ks = 0; ks < 4; ks++) tc_mma(d, al + 2u * ks, bh + 2u * ks, idesc, 1);
                    }
                    seen |= 1u << l;
                    tc_commit(&empty[st]);
                }
            }
            tc_commit(done);
            s_used[0] = (seen & 1u) ? 1 : 0;
            s_used[1] = (seen & 8u) ? 1 : 0;
        }
    } else {
        // ---------------- producers: 256 threads fill the operand tiles of one item
        uint32_t n_item = 0;
        for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int64_t p0 = tile * PT;
            const bool deform = p0 < a.Nu;
            const int nl = deform ? 7 : 3;
            for (int l = 0; l < nl; l++, n_item++) {
                const uint32_t st = n_item & 1u;
                tc_mbar_wait(&empty[st], ((n_item >> 1) & 1u) ^ 1u);
                const WgItem w = wg_item(l);
                uint8_t* Ah = smem + st * WG_STAGE;
                uint8_t* Al = Ah + WG_A;
                uint8_t* Bh = Ah + 2 * WG_A;
                uint8_t* Bl = Bh + WG_B;
                // ---- A = G^T: 128 rows x 16 groups of 8 points; 8 units per thread, every load issued before the first use
                {
                    uint4 u0[8], u1[8];
#pragma unroll
                    for (int it = 0; it < 8; it++) {
                        const int idx = tid + it * 256, f = idx >> 4, q = idx & 15;
                        int grow = -1;
                        if (f < w.a_rows) grow = w.a_row0 + f;
                        else if (f >= 64 && w.a2_row0 >= 0 && deform) grow = w.a2_row0 + (f - 64);
                        u0[it] = make_uint4(0, 0, 0, 0); u1[it] = u0[it];
                        if (grow >= 0) {
                            const uint4* src = reinterpret_cast<const uint4*>(t.G + tile * (int64_t)(GR_ROWS * PT) + grow * PT + 8 * q);
                            u0[it] = __ldg(src); u1[it] = __ldg(src + 1);
                        }
                    }
#pragma unroll
                    for (int it = 0; it < 8; it++) {
                        const int idx = tid + it * 256, f = idx >> 4, q = idx & 15;
                        uint4 hi, lo;
                        hi.x = __byte_perm(u0[it].x, u0[it].y, 0x5410); hi.y = __byte_perm(u0[it].z, u0[it].w, 0x5410);
                        hi.z = __byte_perm(u1[it].x, u1[it].y, 0x5410); hi.w = __byte_perm(u1[it].z, u1[it].w, 0x5410);
                        lo.x = __byte_perm(u0[it].x, u0[it].y, 0x7632); lo.y = __byte_perm(u0[it].z, u0[it].w, 0x7632);
                        lo.z = __byte_perm(u1[it].x, u1[it].y, 0x7632); lo.w = __byte_perm(u1[it].z, u1[it].w, 0x7632);
                        const uint32_t off = (uint32_t)((q >> 3) * WG_AATOM + (f >> 3) * 1024 + (f & 7) * 128 + ((((q & 7) ^ (f & 7)) & 7) << 4));
                        *reinterpret_cast<uint4*>(Ah + off) = hi;
                        *reinterpret_cast<uint4*>(Al + off) = lo;
                    }
                }
                // ---- B = X^T (+ a row of ones): 80 rows x 16 groups of 8 points
                const int kin = w.b_kind == 0 ? ENC : HID;
                const bool ones = w.b_kind <= 2;
                if (w.b_kind == 0) {
                    // grid features are point-major [N][32]: transposing scatter (2-byte stores)
                    for (int idx = tid; idx < 128 * 8; idx += 256) {
                        const int k = idx >> 3, j = idx & 7;                           // point k of the tile, features 4j .. 4j+3
                        const int64_t p = p0 + k;
                        float4 e = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (p < a.N) e = __ldg(reinterpret_cast<const float4*>(a.enc + p * ENC) + j);
                        const float ev[4] = {e.x, e.y, e.z, e.w};
#pragma unroll
                        for (int c = 0; c < 4; c++) {
                            const int i = 4 * j + c;
                            const uint32_t pk = tc_split_pack(ev[c] * kEncScale);
                            const uint32_t off = (uint32_t)((k >> 6) * WG_BATOM + (i >> 3) * 1024 + (i & 7) * 128 + (((((k & 63) >> 3) ^ (i & 7)) & 7) << 4) + (k & 7) * 2);
                            *reinterpret_cast<uint16_t*>(Bh + off) = (uint16_t)(pk & 0xffffu);
                            *reinterpret_cast<uint16_t*>(Bl + off) = (uint16_t)(pk >> 16);
                        }
                    }
                }
                const int row_begin = w.b_kind == 0 ? ENC : 0;
                const float* X = nullptr;
                int64_t Nvalid = 0;
                if (w.b_kind == 1 || w.b_kind == 3) { X = a.acts_s + (int64_t)w.b_layer * HID * ((a.N + PT - 1) / PT * PT) + tile * (int64_t)(HID * PT); Nvalid = a.N; }
                if (w.b_kind == 2 || w.b_kind == 4) { X = a.acts_d + (int64_t)w.b_layer * HID * ((a.Nu + PT - 1) / PT * PT) + tile * (int64_t)(HID * PT); Nvalid = a.Nu; }
                {
                    // 80 rows x 16 groups = 1280 units, 5 per thread; loads first (rows below row_begin were written above)
                    float4 x0[5], x1[5];
#pragma unroll
                    for (int it = 0; it < 5; it++) {
                        const int idx = tid + it * 256, i = idx >> 4, q = idx & 15;
                        x0[it] = make_float4(0.f, 0.f, 0.f, 0.f); x1[it] = x0[it];
                        if (i >= row_begin && i < kin) {
                            const int64_t p = p0 + 8 * q;
                            const float* src = X + i * PT + 8 * q;
                            if (p + 7 < Nvalid) { x0[it] = __ldg(reinterpret_cast<const float4*>(src)); x1[it] = __ldg(reinterpret_cast<const float4*>(src) + 1); }
                            else {
                                float x[8];
#pragma unroll
                                for (int c = 0; c < 8; c++) x[c] = (p + c < Nvalid) ? src[c] : 0.f;
                                x0[it] = make_float4(x[0], x[1], x[2], x[3]); x1[it] = make_float4(x[4], x[5], x[6], x[7]);
                            }
                        }
                    }
#pragma unroll
                    for (int it = 0; it < 5; it++) {
                        const int idx = tid + it * 256, i = idx >> 4, q = idx & 15;
                        if (i < row_begin) continue;
                        uint4 hi = make_uint4(0, 0, 0, 0), lo = hi;
                        if (i < kin) {
                            const float x[8] = {x0[it].x, x0[it].y, x0[it].z, x0[it].w, x1[it].x, x1[it].y, x1[it].z, x1[it].w};
                            uint32_t ph[4], pl[4];
#pragma unroll
                            for (int c = 0; c < 4; c++) {
                                ph[c] = pack_act2(x[2 * c], x[2 * c + 1]);
                                const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&ph[c]));
                                pl[c] = pack_act2(x[2 * c] - hf.x, x[2 * c + 1] - hf.y);
                            }
                            hi = make_uint4(ph[0], ph[1], ph[2], ph[3]);
                            lo = make_uint4(pl[0], pl[1], pl[2], pl[3]);
                        } else if (i == kin && ones) {
                            hi = make_uint4(0x3C003C00u, 0x3C003C00u, 0x3C003C00u, 0x3C003C00u);     // 1.0h x 8
                        }
                        const uint32_t off = (uint32_t)((q >> 3) * WG_BATOM + (i >> 3) * 1024 + (i & 7) * 128 + ((((q & 7) ^ (i & 7)) & 7) << 4));
                        *reinterpret_cast<uint4*>(Bh + off) = hi;
                        *reinterpret_cast<uint4*>(Bl + off) = lo;
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                tc_fence_before();
                tc_mbar_arrive(&full[st]);
            }
        }
    }
    // ---------------- every accumulator -> this CTA's partial gradient vector
    tc_mbar_wait(done, 0);
    tc_fence_after();
    __syncthreads();
    float* out = partial + (int64_t)blockIdx.x * NPARAM;
    const bool used = s_used[0] != 0, used_d = s_used[1] != 0;
    if (warp < 4) {
        const int o = (warp << 5) + lane;                            // TMEM lane = row of the accumulator
        const uint32_t tl = tmem_base + ((uint32_t)(warp * 32) << 16);
        float v[32];
        // L0: lanes 0-63 sigma layer 1, lanes 64-127 deform layer 0; columns 0-31 weights (x 2^-10), column 32 bias
        tc_ld32(tl + WC_L0, v);
        {
            const bool ok = o < 64 ? used : used_d;
            float* wdst = out + (o < 64 ? S1W + o * ENC : D0W + (o - 64) * ENC);
#pragma unroll
            for (int i = 0; i < ENC; i++) wdst[i] = ok ? v[i] * kEncInv : 0.f;
        }
        float bcol[32];
        tc_ld32(tl + WC_L0 + 32, bcol);
        out[o < 64 ? S1B + o : D0B + (o - 64)] = (o < 64 ? used : used_d) ? bcol[0] : 0.f;
        if (o < 64) {
            auto dump64 = [&](int col, int wofs, int bofs, bool ok) {
                tc_ld32(tl + col, v);
#pragma unroll
                for (int i = 0; i < 32; i++) out[wofs + o * HID + i] = ok ? v[i] : 0.f;
                tc_ld32(tl + col + 32, v);
#pragma unroll
                for (int i = 0; i < 32; i++) out[wofs + o * HID + 32 + i] = ok ? v[i] : 0.f;
                if (bofs >= 0) {
                    tc_ld32(tl + col + 64, v);
                    out[bofs + o] = ok ? v[0] : 0.f;
                }
            };
            dump64(WC_S2, S2W, S2B, used);
            dump64(WC_D1, D1W, D1B, used_d);
            dump64(WC_D2, D2W, D2B, used_d);
            dump64(WC_D3, D3W, D3B, used_d);
        }
        // head layers: a handful of rows (all lanes of the warp execute the collective loads)
        {
            float h0[32], h1[32];
            tc_ld32(tl + WC_S3, h0); tc_ld32(tl + WC_S3 + 32, h1);
            if (o < 4) {
#pragma unroll
                for (int i = 0; i < 32; i++) { out[S3W + o * HID + i] = used ? h0[i] : 0.f; out[S3W + o * HID + 32 + i] = used ? h1[i] : 0.f; }
            }
            tc_ld32(tl + WC_DH, h0); tc_ld32(tl + WC_DH + 32, h1);
            if (o < 6) {
                float* wdst = out + (o < 3 ? WPW + o * HID : SCW + (o - 3) * HID);
#pragma unroll
                for (int i = 0; i < 32; i++) { wdst[i] = used_d ? h0[i] : 0.f; wdst[32 + i] = used_d ? h1[i] : 0.f; }
            }
        }
    }
    if (tid < 10) out[tid < 4 ? S3B + tid : (tid < 7 ? WPB + tid - 4 : SCB + tid - 7)] = 0.f;      // head biases come from K1 (partial1)
    tc_fence_before();
    __syncthreads();
    if (warp == 8) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "n"(512));
}

// g_params[j] = sum over CTAs of partial[c][j] (+ the head-bias partials of K1); g_w_pose[o][j] = g_bias0[o] * pose[j]
__global__ void __launch_bounds__(256)
mlp_reduce_tc_kernel(const float* __restrict__ partial, int nparts, const float* __restrict__ partial1, int nparts1,
                     const float* __restrict__ pose, float* __restrict__ g_params, float* __restrict__ g_w_pose) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= NPARAM) return;
    float s = 0.f;
    for (int c = 0; c < nparts; c++) s += partial[(int64_t)c * NPARAM + j];
    int hb = -1;
    if (j >= S3B && j < S3B + 4) hb = j - S3B;
    else if (j >= WPB && j < WPB + 3) hb = 4 + j - WPB;
    else if (j >= SCB && j < SCB + 3) hb = 7 + j - SCB;
    if (hb >= 0) for (int c = 0; c < nparts1; c++) s += partial1[c * 16 + hb];
    g_params[j] = s;
    if (g_w_pose && j >= D0B && j < D0B + HID) {
        const int o = j - D0B;
        for (int k = 0; k < POSE; k++) g_w_pose[o * POSE + k] = s * pose[k];
    }
}

static int g_mlp_tc = 1;      // debug switch (dwg_avatar_mlp_set_tc): 0 = the fp32 SIMT forward

static int g_sms = 0;
static int num_sms() {
    if (!g_sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev);
        if (g_sms <= 0) g_sms = kNumSMs;
    }
    return g_sms;
}
constexpr size_t kFwdSmem = sizeof(float) * (size_t)(2 * ENC * HID + 4 * HID * HID + HID * 4 + HID * 8 + 7 * HID + 4 + 8 + HID * P) + 16;
constexpr size_t kBwdSmem = sizeof(float) * (size_t)(((NPARAM + 3) & ~3) + 2 * HID * S) + 16;

}  // namespace mlp
}  // namespace dwg

using namespace dwg;
using namespace dwg::mlp;

extern "C" int64_t dwg_avatar_mlp_param_count(void) { return NPARAM; }
/* Debug / A-B switch: 1 (default) = tensor-core kernels, 0 = the fp32 SIMT kernels. */
extern "C" int dwg_avatar_mlp_set_tc(int on) { g_mlp_tc = on ? 1 : 0; return DWG_OK; }
extern "C" int64_t dwg_avatar_mlp_scratch_bytes(void) { return (int64_t)sizeof(float) * NPARAM * num_sms(); }
/* Scratch of dwg_avatar_mlp_bwd for N Gaussians: per-CTA partial gradient vectors + (tensor-core path) the packed
 * per-layer gradients G [394][ceil(N / 128) * 128] that the input-gradient kernel hands to the weight-gradient kernel. */
extern "C" int64_t dwg_avatar_mlp_bwd_scratch_bytes(int64_t N) {
    const int64_t Ng = (N + PT - 1) / PT * PT;
    return (int64_t)sizeof(float) * (NPARAM + 16) * num_sms() + (int64_t)sizeof(uint32_t) * GR_ROWS * Ng + 256;
}

// Forward of both MLPs + activations for N Gaussians (the first Nu are unconstrained: both nets; the
// rest are mesh-bound: colour only, opacity 1).  acts_s [2][64][Np], acts_d [4][64][Nup] (Np, Nup =
// N, Nu rounded up to a multiple of 4) receive the hidden activations for the backward (pass null
// for inference).
extern "C" int dwg_avatar_mlp_fwd(const float* enc, const float* positions, const float* params, const float* w_pose, const float* body_pose,
                                  float* colors, float* opac, float* pos_out, float* scales, float* acts_s, float* acts_d,
                                  int64_t N, int64_t Nu, float init_offset, float init_scale, float max_scale, void* stream) {
    DWG_REQUIRE(enc && params && w_pose && body_pose && colors && opac, "null pointer");
    DWG_REQUIRE(N > 0 && Nu >= 0 && Nu <= N, "bad sizes");
    DWG_REQUIRE(Nu == 0 || (positions && pos_out && scales), "null pointer (unconstrained outputs)");
    DWG_REQUIRE(((uintptr_t)enc & 15) == 0, "enc must be 16-byte aligned");
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(mlp_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFwdSmem); attr = true; }
    FwdArgs a;
    a.enc = enc; a.positions = positions; a.params = params; a.w_pose = w_pose; a.pose = body_pose;
    a.colors = colors; a.opac = opac; a.pos_out = pos_out; a.scales = scales; a.acts_s = acts_s; a.acts_d = acts_d;
    a.N = N; a.Nu = Nu; a.Np = (N + 3) & ~(int64_t)3; a.Nup = (Nu + 3) & ~(int64_t)3;
    a.init_offset = init_offset; a.init_scale = init_scale; a.max_scale = max_scale;
    if (g_mlp_tc) {
        static bool attr_tc = false;
        if (!attr_tc) { cudaFuncSetAttribute(mlp_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcFwdSmem); attr_tc = true; }
        const int64_t npairs = ((N + PT - 1) / PT + 1) / 2;
        const int grid = (int)(npairs < num_sms() ? npairs : num_sms());
        mlp_fwd_tc_kernel<<<grid, 256, kTcFwdSmem, (cudaStream_t)stream>>>(a);
        return check_launch("dwg_avatar_mlp_fwd");
    }
    const int64_t ntiles = (N + P - 1) / P;
    const int grid = (int)(ntiles < num_sms() ? ntiles : num_sms());
    mlp_fwd_kernel<<<grid, P, kFwdSmem, (cudaStream_t)stream>>>(a);
    return check_launch("dwg_avatar_mlp_fwd");
}

// Backward: g_enc [N,32] (overwritten), g_params [NPARAM] and g_w_pose [64,63] (overwritten); `scratch`
// = dwg_avatar_mlp_scratch_bytes() bytes.  Null output-gradient pointers mean zero.
extern "C" int dwg_avatar_mlp_bwd(const float* enc, const float* params, const float* body_pose,
                                  const float* colors, const float* opac, const float* scales, const float* acts_s, const float* acts_d,
                                  const float* g_colors, const float* g_opac, const float* g_pos, const float* g_scales,
                                  float* g_enc, float* g_params, float* g_w_pose, void* scratch,
                                  int64_t N, int64_t Nu, float init_offset, float max_scale, void* stream) {
    DWG_REQUIRE(enc && params && body_pose && colors && opac && acts_s && g_enc && g_params && scratch, "null pointer");
    DWG_REQUIRE(N > 0 && Nu >= 0 && Nu <= N, "bad sizes");
    DWG_REQUIRE(Nu == 0 || (scales && acts_d), "null pointer (unconstrained inputs)");
    DWG_REQUIRE(((uintptr_t)enc & 15) == 0 && ((uintptr_t)g_enc & 15) == 0 && ((uintptr_t)acts_s & 15) == 0 && ((uintptr_t)acts_d & 15) == 0,
                "enc / g_enc / activations must be 16-byte aligned");
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(mlp_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBwdSmem); attr = true; }
    BwdArgs a;
    a.enc = enc; a.params = params; a.colors = colors; a.opac = opac; a.scales = scales; a.acts_s = acts_s; a.acts_d = acts_d;
    a.g_colors = g_colors; a.g_opac = g_opac; a.g_pos = g_pos; a.g_scales = g_scales;
    a.g_enc = g_enc; a.partial = reinterpret_cast<float*>(scratch);
    a.N = N; a.Nu = Nu; a.Np = (N + 3) & ~(int64_t)3; a.Nup = (Nu + 3) & ~(int64_t)3;
    a.init_offset = init_offset; a.max_scale = max_scale;
    cudaStream_t st = (cudaStream_t)stream;
    if (g_mlp_tc) {
        static bool attr_tc = false;
        if (!attr_tc) {
            cudaFuncSetAttribute(mlp_bwd_tc_dgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcDgradSmem);
            cudaFuncSetAttribute(mlp_bwd_tc_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcWgradSmem);
            attr_tc = true;
        }
        const int sms = num_sms();
        const int64_t nt = (N + PT - 1) / PT;
        BwdTcArgs t;
        t.b = a;
        t.Ng = nt * PT;
        float* partial2 = reinterpret_cast<float*>(scratch);
        t.partial1 = partial2 + (int64_t)NPARAM * sms;
        t.G = reinterpret_cast<uint32_t*>((reinterpret_cast<uintptr_t>(t.partial1 + 16 * (int64_t)sms) + 255) & ~(uintptr_t)255);
        const int grid1 = (int)((nt + 1) / 2 < sms ? (nt + 1) / 2 : sms);
        const int grid2 = (int)(nt < sms ? nt : sms);
        mlp_bwd_tc_dgrad_kernel<<<grid1, 256, kTcDgradSmem, st>>>(t);
        mlp_bwd_tc_wgrad_kernel<<<grid2, 288, kTcWgradSmem, st>>>(t, partial2);
        mlp_reduce_tc_kernel<<<(NPARAM + 255) / 256, 256, 0, st>>>(partial2, grid2, t.partial1, grid1, body_pose, g_params, g_w_pose);
        return check_launch("dwg_avatar_mlp_bwd");
    }
    const int64_t ntiles = (N + P - 1) / P;
    const int grid = (int)(ntiles < num_sms() ? ntiles : num_sms());
    mlp_bwd_kernel<<<grid, P, kBwdSmem, st>>>(a);
    mlp_reduce_kernel<<<(NPARAM + 255) / 256, 256, 0, st>>>(a.partial, grid, body_pose, g_params, g_w_pose);
    return check_launch("dwg_avatar_mlp_bwd");
}
