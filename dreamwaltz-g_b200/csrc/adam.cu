// (f2) Fused multi-tensor Adam over the flat gradient bucket: ONE kernel updates every trainable avatar parameter.
//
// Replaces the four torch.optim.Adam instances the reference steps every iteration (core/trainer.py:880-882): 'avatar'
// (GaussianOptimizer, core/gaussian/gaussian_optimizer.py:49-141: per-name learning rates, eps 1e-15, positions on the
// exponential schedule of core/optim/optim_utils.py:5-40 times the spatial scale), 'nerf' (core/system/avatar.py:1619-1626:
// betas (0.9, 0.99), eps 1e-15, grid encoder at 10x lr) and 'mesh_*' (avatar.py:1081-1094).  Parameters, gradients and both
// moments live in flat fp32 buffers with 16-byte aligned segments (dwg/parallel.py GradBucket), so the update is a pure
// HBM stream: 16 B/element read (p, g, m, v) + 12 B/element written, 128-bit accesses, one segment lookup per float4.
// The step counter stays on the device (CUDA-graph replay needs no host value); the per-group learning rates are a
// small device table the host refreshes asynchronously (position schedule).
#include "common.cuh"

namespace dwg {
namespace {

constexpr int kMaxSeg = 128, kMaxGroup = 16;

struct AdamTable {
    int n_seg, n_group;
    int64_t seg_end[kMaxSeg];        // exclusive end offset (elements) of each segment in the flat buffers
    int seg_group[kMaxSeg];
    float beta1[kMaxGroup], beta2[kMaxGroup], eps[kMaxGroup];
};

// torch.optim.Adam (no amsgrad, no weight decay, maximize=False), fp32:
//   m = lerp(m, g, 1-b1); v = b2 v + (1-b2) g^2; p -= (lr / (1-b1^t)) * m / (sqrt(v) / sqrt(1-b2^t) + eps)
__global__ void __launch_bounds__(256)
adam_kernel(float4* __restrict__ p, const float4* __restrict__ g, float4* __restrict__ m, float4* __restrict__ v,
            const float* __restrict__ lr, long long* __restrict__ step, const __grid_constant__ AdamTable tab, int64_t n4) {
    __shared__ float s_ss[kMaxGroup], s_bc2[kMaxGroup];
    const long long t = *step + 1;                               // every thread reads the pre-increment value (bumped by the last block below)
    if (threadIdx.x < tab.n_group) {
        const int k = threadIdx.x;
        const double bc1 = 1.0 - pow((double)tab.beta1[k], (double)t), bc2 = 1.0 - pow((double)tab.beta2[k], (double)t);
        s_ss[k] = (float)((double)lr[k] / bc1);
        s_bc2[k] = (float)sqrt(bc2);
    }
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t e = i * 4;
        int lo = 0, hi = tab.n_seg - 1;                          // first segment whose end > e
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (tab.seg_end[mid] > e) hi = mid; else lo = mid + 1; }
        const int k = tab.seg_group[lo];
        const float b1 = tab.beta1[k], b2 = tab.beta2[k], eps = tab.eps[k], ss = s_ss[k], bc2 = s_bc2[k];
        const float4 gg = g[i];
        float4 pp = p[i], mm = m[i], vv = v[i];
        auto upd = [&](float& pw, float gw, float& mw, float& vw) {
            mw = fmaf(gw - mw, 1.0f - b1, mw);
            vw = fmaf(vw, b2, (1.0f - b2) * gw * gw);
            pw -= ss * (mw / (sqrtf(vw) / bc2 + eps));
        };
        upd(pp.x, gg.x, mm.x, vv.x); upd(pp.y, gg.y, mm.y, vv.y); upd(pp.z, gg.z, mm.z, vv.z); upd(pp.w, gg.w, mm.w, vv.w);
        p[i] = pp; m[i] = mm; v[i] = vv;
    }
}

__global__ void adam_bump_kernel(long long* step) { *step += 1; }

}  // namespace
}  // namespace dwg

using namespace dwg;

extern "C" int dwg_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                             int n_seg, const int64_t* seg_end, const int32_t* seg_group,
                             int n_group, const float* beta1, const float* beta2, const float* eps,
                             const float* lr_dev, int64_t* step_dev, void* stream) {
    DWG_REQUIRE(params && grads && exp_avg && exp_avg_sq && seg_end && seg_group && beta1 && beta2 && eps && lr_dev && step_dev, "null pointer");
    DWG_REQUIRE(n > 0 && (n % 4) == 0, "flat length must be a positive multiple of 4 (16-byte aligned segments)");
    DWG_REQUIRE(n_seg > 0 && n_seg <= kMaxSeg && n_group > 0 && n_group <= kMaxGroup, "too many segments / groups");
    DWG_REQUIRE(((reinterpret_cast<uintptr_t>(params) | reinterpret_cast<uintptr_t>(grads) | reinterpret_cast<uintptr_t>(exp_avg) |
                  reinterpret_cast<uintptr_t>(exp_avg_sq)) & 15) == 0, "flat buffers must be 16-byte aligned");
    AdamTable tab = {};
    tab.n_seg = n_seg; tab.n_group = n_group;
    for (int i = 0; i < n_seg; i++) {
        DWG_REQUIRE(seg_end[i] % 4 == 0 && seg_group[i] >= 0 && seg_group[i] < n_group && (i == 0 || seg_end[i] > seg_end[i - 1]), "bad segment table");
        tab.seg_end[i] = seg_end[i]; tab.seg_group[i] = seg_group[i];
    }
    DWG_REQUIRE(seg_end[n_seg - 1] == n, "segments must cover the flat buffer");
    for (int k = 0; k < n_group; k++) { tab.beta1[k] = beta1[k]; tab.beta2[k] = beta2[k]; tab.eps[k] = eps[k]; }
    const int64_t n4 = n / 4;
    int64_t blocks = (n4 + 255) / 256;
    if (blocks > 8 * kNumSMs) blocks = 8 * kNumSMs;
    cudaStream_t st = (cudaStream_t)stream;
    adam_kernel<<<(unsigned)blocks, 256, 0, st>>>(reinterpret_cast<float4*>(params), reinterpret_cast<const float4*>(grads),
                                                  reinterpret_cast<float4*>(exp_avg), reinterpret_cast<float4*>(exp_avg_sq), lr_dev,
                                                  reinterpret_cast<long long*>(step_dev), tab, n4);
    adam_bump_kernel<<<1, 1, 0, st>>>(reinterpret_cast<long long*>(step_dev));
    return check_launch("dwg_adam_step");
}
