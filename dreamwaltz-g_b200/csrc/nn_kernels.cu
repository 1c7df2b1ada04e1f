// Normalisation / activation / softmax kernels of the diffusion blocks (R14/R15), NHWC bf16.
// All are HBM-bound streaming kernels: 128-bit loads (8 bf16), fp32 math, 128-bit stores.
//   groupnorm (+SiLU) forward: stats pass (per-(image,group) sum / sumsq via block reduction and
//   a few atomics) + apply pass; backward (VAE encoder input-gradient) the same two-pass shape;
//   layernorm over channels (one warp per token); row softmax (bf16 in place); GEGLU; SiLU;
//   the fused CFG + SDS-gradient epilogue (core/guidance/basic.py:595-603,642).

#include "common.cuh"

namespace dwg {
namespace nn {

struct h8 { act2_t v[4]; };
static_assert(sizeof(h8) == 16, "h8");

__device__ __forceinline__ void unpack8(const h8& p, float f[8]) {
#pragma unroll
    for (int i = 0; i < 4; i++) { const float2 t = act2_to_f2(p.v[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
}
__device__ __forceinline__ h8 pack8(const float f[8]) {
    h8 p;
#pragma unroll
    for (int i = 0; i < 4; i++) p.v[i] = f2_to_act2(f[2 * i], f[2 * i + 1]);
    return p;
}
// These kernels are issue-bound on the big VAE tensors: sigmoid(x) = 0.5 + 0.5 tanh(x/2) costs ONE MUFU op
// (tanh.approx.f32, max rel err 2^-11, below the bf16 rounding of every result) instead of exp + reciprocal.
__device__ __forceinline__ float tanh_fast(float x) { float y; asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float silu(float x) { const float h = 0.5f * x; return fmaf(h, tanh_fast(h), h); }
__device__ __forceinline__ float silu_grad(float x) { const float s = fmaf(0.5f, tanh_fast(0.5f * x), 0.5f); return s * fmaf(x, 1.0f - s, 1.0f); }

// ---------------------------------------------------------------------------- GroupNorm
// Run-to-run DETERMINISTIC statistics: every per-thread fp32 partial sum is converted to 64-bit fixed point and
// accumulated with INTEGER atomics (shared, then global) -- integer addition is associative, so the result does not
// depend on the arrival order, unlike fp32 atomics whose last-bit differences are blown up to the full 16-bit rounding
// noise by the next few layers (tools/determinism_probe.py).  Forward moments use 2^-20 resolution (|sum| < 2^43);
// the backward sums (gradients, much smaller) 2^-36 (|sum| < 2^27).  Mean / variance are then formed in double, which
// also removes the one-pass E[x^2] - E[x]^2 cancellation on large-mean activations.
typedef unsigned long long fix_t;
constexpr float kFixFwd = 1048576.0f;               // 2^20
constexpr float kFixBwd = 68719476736.0f;           // 2^36
__device__ __forceinline__ fix_t to_fix(float v, float scale) { return (fix_t)__float2ll_rn(v * scale); }
__device__ __forceinline__ double from_fix(fix_t v, double inv_scale) { return (double)(long long)v * inv_scale; }
// (mean, rstd) of group idx = n * G + g from the fixed-point (sum, sumsq)
__device__ __forceinline__ void gn_moments(const fix_t* __restrict__ stats, size_t idx, double inv_cnt, float eps, float& mean, float& rstd) {
    const double m = from_fix(stats[2 * idx], inv_cnt * (1.0 / 1048576.0));
    const double q = from_fix(stats[2 * idx + 1], inv_cnt * (1.0 / 1048576.0));
    mean = (float)m;
    rstd = rsqrtf((float)fmax(q - m * m, 0.0) + eps);
}

// x [N, HW, C] fp16, C % 8 == 0, (C/G) % 8 == 0 or 8 % (C/G) == 0 handled generically via smem bins.
// stats[n][g] = fixed-point (sum, sumsq) accumulated with integer atomics (zeroed by the entry point).
__global__ void __launch_bounds__(256)
gn_stats_kernel(const act_t* __restrict__ x, fix_t* __restrict__ stats, int HW, int C, int G, int rows_per_cta) {
    pdl_wait();
    pdl_trigger();
    extern __shared__ fix_t s_bins[];        // [G][2]
    const int n = blockIdx.y;
    const int cpg = C / G;
    for (int i = threadIdx.x; i < 2 * G; i += blockDim.x) s_bins[i] = 0ull;
    __syncthreads();
    const int c8 = C / 8;                    // 16-byte vectors per row
    const int row0 = blockIdx.x * rows_per_cta;
    const int row1 = min(HW, row0 + rows_per_cta);
    const h8* xp = reinterpret_cast<const h8*>(x + (size_t)n * HW * C);
    // thread -> (row lane, fixed vector column): per-thread register accumulation over rows, a
    // handful of shared-memory atomics at the end
    const int rp = c8 <= 256 ? 256 / c8 : 1;
    const int rl = c8 <= 256 ? threadIdx.x / c8 : 0;
    for (int cv = c8 <= 256 ? threadIdx.x % c8 : threadIdx.x; cv < c8 && rl < rp; cv += 256) {
        float s[8], ss[8];
#pragma unroll
        for (int i = 0; i < 8; i++) { s[i] = 0.f; ss[i] = 0.f; }
        int r = row0 + rl;
        for (; r + 3 * rp < row1; r += 4 * rp) {          // 4 independent 16-byte loads in flight
            h8 v0 = xp[(size_t)r * c8 + cv], v1 = xp[(size_t)(r + rp) * c8 + cv];
            h8 v2 = xp[(size_t)(r + 2 * rp) * c8 + cv], v3 = xp[(size_t)(r + 3 * rp) * c8 + cv];
            float f[8];
            unpack8(v0, f);
#pragma unroll
            for (int i = 0; i < 8; i++) { s[i] += f[i]; ss[i] += f[i] * f[i]; }
            unpack8(v1, f);
#pragma unroll
            for (int i = 0; i < 8; i++) { s[i] += f[i]; ss[i] += f[i] * f[i]; }
            unpack8(v2, f);
#pragma unroll
            for (int i = 0; i < 8; i++) { s[i] += f[i]; ss[i] += f[i] * f[i]; }
            unpack8(v3, f);
#pragma unroll
            for (int i = 0; i < 8; i++) { s[i] += f[i]; ss[i] += f[i] * f[i]; }
        }
        for (; r < row1; r += rp) {
            float f[8];
            unpack8(xp[(size_t)r * c8 + cv], f);
#pragma unroll
            for (int i = 0; i < 8; i++) { s[i] += f[i]; ss[i] += f[i] * f[i]; }
        }
        if (cpg % 8 == 0) {
            float a = 0.f, b = 0.f;
#pragma unroll
            for (int i = 0; i < 8; i++) { a += s[i]; b += ss[i]; }
            const int g = (cv * 8) / cpg;
            atomicAdd(&s_bins[2 * g], to_fix(a, kFixFwd));
            atomicAdd(&s_bins[2 * g + 1], to_fix(b, kFixFwd));
        } else {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const int g = (cv * 8 + i) / cpg;
                atomicAdd(&s_bins[2 * g], to_fix(s[i], kFixFwd));
                atomicAdd(&s_bins[2 * g + 1], to_fix(ss[i], kFixFwd));
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * G; i += blockDim.x) atomicAdd(&stats[(size_t)n * 2 * G + i], s_bins[i]);
}

// y = (x - mean) * rstd * gamma + beta, optional SiLU.  grid (row chunks, N); thread -> (row lane,
// fixed 8-channel column): scale/shift are computed once per thread, then one 16-byte load, 8 FMAs
// and one 16-byte store per row.
__global__ void __launch_bounds__(256)
gn_apply_kernel(const act_t* __restrict__ x, const fix_t* __restrict__ stats, const float* __restrict__ gamma,
                const float* __restrict__ beta, act_t* __restrict__ y, int HW, int C, int G, float eps, int do_silu,
                int rows_per_cta) {
    pdl_wait();
    pdl_trigger();
    const int n = blockIdx.y;
    const int c8 = C / 8, cpg = C / G;
    const double inv_cnt = 1.0 / ((double)HW * (double)cpg);
    const int row0 = blockIdx.x * rows_per_cta, row1 = min(HW, row0 + rows_per_cta);
    const h8* xp = reinterpret_cast<const h8*>(x + (size_t)n * HW * C);
    h8* yp = reinterpret_cast<h8*>(y + (size_t)n * HW * C);
    const int rp = c8 <= 256 ? 256 / c8 : 1;
    const int rl = c8 <= 256 ? threadIdx.x / c8 : 0;
    for (int cv = c8 <= 256 ? threadIdx.x % c8 : threadIdx.x; cv < c8 && rl < rp; cv += 256) {
        float a[8], b[8];
        int g_prev = -1;
        float mean = 0.f, rstd = 0.f;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int c = cv * 8 + i, g = c / cpg;
            if (g != g_prev) { gn_moments(stats, (size_t)n * G + g, inv_cnt, eps, mean, rstd); g_prev = g; }
            a[i] = rstd * gamma[c];
            b[i] = beta[c] - mean * a[i];
        }
        int r = row0 + rl;
        for (; r + 3 * rp < row1; r += 4 * rp) {          // 4 independent 16-byte loads in flight
            h8 v[4];
#pragma unroll
            for (int u = 0; u < 4; u++) v[u] = xp[(size_t)(r + u * rp) * c8 + cv];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                float f[8];
                unpack8(v[u], f);
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    float o = fmaf(f[i], a[i], b[i]);
                    f[i] = do_silu ? silu(o) : o;
                }
                yp[(size_t)(r + u * rp) * c8 + cv] = pack8(f);
            }
        }
        for (; r < row1; r += rp) {
            float f[8];
            unpack8(xp[(size_t)r * c8 + cv], f);
#pragma unroll
            for (int i = 0; i < 8; i++) {
                float o = fmaf(f[i], a[i], b[i]);
                f[i] = do_silu ? silu(o) : o;
            }
            yp[(size_t)r * c8 + cv] = pack8(f);
        }
    }
}

// GroupNorm (+SiLU) whose statistics were produced by the GEMM / conv epilogue of the layer that wrote x (per-column fixed-point
// sums, csrc/gemm_tcgen05.cu): no statistics pass and no memset -- the prologue folds the cpg column sums of each group
// (integer adds: exact and order-free), the body is the apply pass.  CTA 0 of every image also writes the group statistics
// for the backward.
__global__ void __launch_bounds__(256, 3)
gn_apply_cs_kernel(const act_t* __restrict__ x, const fix_t* __restrict__ colstats, const float* __restrict__ gamma,
                   const float* __restrict__ beta, act_t* __restrict__ y, fix_t* __restrict__ stats_out, int HW, int C, int G, float eps,
                   int do_silu, int rows_per_cta, const act_t* __restrict__ x2, const fix_t* __restrict__ colstats2, int C1) {
    // x2 != null: the input is the channel concatenation [x (C1 channels) | x2 (C - C1 channels)] of two tensors that are
    // never materialised side by side (the UNet decoder's skip concat): columns < C1 come from x / colstats, the rest from x2.
    pdl_wait();
    pdl_trigger();
    __shared__ float s_mean[64], s_rstd[64];
    const int n = blockIdx.y;
    const int c8 = C / 8, cpg = C / G;
    // The kernel is three dependent global round trips long (column statistics -> affine parameters -> first rows): the
    // affine parameters and the first rows of this thread do not depend on the statistics, so their loads are issued
    // BEFORE the fold and are in flight while it runs.
    const int row0 = blockIdx.x * rows_per_cta, row1 = min(HW, row0 + rows_per_cta);
    const int c8a = C1 / 8, c8b = c8 - c8a;
    const h8* xpa = reinterpret_cast<const h8*>(x + (size_t)n * HW * C1);
    const h8* xpb = x2 ? reinterpret_cast<const h8*>(x2 + (size_t)n * HW * (C - C1)) : nullptr;
    auto load_x = [&](int r, int cv) { return cv < c8a ? xpa[(size_t)r * c8a + cv] : xpb[(size_t)r * c8b + (cv - c8a)]; };
    h8* yp = reinterpret_cast<h8*>(y + (size_t)n * HW * C);
    const int rp = c8 <= 256 ? 256 / c8 : 1;
    const int rl = c8 <= 256 ? threadIdx.x / c8 : 0;
    const int cv0 = c8 <= 256 ? threadIdx.x % c8 : threadIdx.x;
    const bool first_ok = cv0 < c8 && rl < rp;
    float gpre[8], bpre[8];
    h8 vpre[4];
    const bool pre_rows = first_ok && (row0 + rl + 3 * rp < row1);
    if (first_ok) {
#pragma unroll
        for (int i = 0; i < 8; i++) { gpre[i] = gamma[cv0 * 8 + i]; bpre[i] = beta[cv0 * 8 + i]; }
    }
    if (pre_rows) {
#pragma unroll
        for (int u = 0; u < 4; u++) vpre[u] = load_x(row0 + rl + u * rp, cv0);
    }
    {
        // 8 threads per group (G <= 32) fold the group's columns
        const int g = threadIdx.x >> 3, sub = threadIdx.x & 7;
        long long S = 0, Q = 0;
        if (g < G) {
            const int Cb = C - C1;
            for (int slot = 0; slot < 4; slot++) {                // the producer spreads its atomics over 4 slots (m_tile & 3)
                const fix_t* csa = colstats + (((size_t)slot * gridDim.y + n) * C1) * 2;
                const fix_t* csb = colstats2 ? colstats2 + (((size_t)slot * gridDim.y + n) * Cb) * 2 : nullptr;
                for (int k = sub; k < cpg; k += 8) {
                    const int c = g * cpg + k;
                    const fix_t* cs = c < C1 ? csa + 2 * (size_t)c : csb + 2 * (size_t)(c - C1);
                    S += (long long)cs[0]; Q += (long long)cs[1];
                }
            }
        }
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) { S += __shfl_xor_sync(0xffffffffu, S, o); Q += __shfl_xor_sync(0xffffffffu, Q, o); }
        if (g < G && sub == 0) {
            const double inv = 1.0 / ((double)HW * (double)cpg) * (1.0 / 1048576.0);
            const double m = (double)S * inv, q = (double)Q * inv;
            s_mean[g] = (float)m;
            s_rstd[g] = rsqrtf((float)fmax(q - m * m, 0.0) + eps);
            if (stats_out && blockIdx.x == 0) { stats_out[((size_t)n * G + g) * 2] = (fix_t)S; stats_out[((size_t)n * G + g) * 2 + 1] = (fix_t)Q; }
        }
    }
    __syncthreads();
    for (int cv = cv0; cv < c8 && rl < rp; cv += 256) {
        float a[8], b[8];
        const bool first = cv == cv0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int c = cv * 8 + i, g = c / cpg;
            a[i] = s_rstd[g] * (first ? gpre[i] : gamma[c]);
            b[i] = (first ? bpre[i] : beta[c]) - s_mean[g] * a[i];
        }
        int r = row0 + rl;
        for (; r + 3 * rp < row1; r += 4 * rp) {
            h8 v[4];
            if (first && pre_rows && r == row0 + rl) {
#pragma unroll
                for (int u = 0; u < 4; u++) v[u] = vpre[u];
            } else {
#pragma unroll
                for (int u = 0; u < 4; u++) v[u] = load_x(r + u * rp, cv);
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                float f[8];
                unpack8(v[u], f);
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    float o = fmaf(f[i], a[i], b[i]);
                    f[i] = do_silu ? silu(o) : o;
                }
                yp[(size_t)(r + u * rp) * c8 + cv] = pack8(f);
            }
        }
        for (; r < row1; r += rp) {
            float f[8];
            unpack8(load_x(r, cv), f);
#pragma unroll
            for (int i = 0; i < 8; i++) {
                float o = fmaf(f[i], a[i], b[i]);
                f[i] = do_silu ? silu(o) : o;
            }
            yp[(size_t)r * c8 + cv] = pack8(f);
        }
    }
}

// One-launch GroupNorm (+SiLU) for tensors that fit the shared memory of one cluster per (image, 4-group slab)
// -- every GroupNorm of the UNet / ControlNet at batch 2.  A cluster of GN_CS CTAs owns one image and a slab of
// 4 consecutive groups (4*cpg channels, always a multiple of 8); CTA r of the cluster takes the r-th share of the
// pixels: (1) load its [rows x slab] block once into shared memory while accumulating per-group sum / sumsq,
// (2) exchange the 8 partial sums through distributed shared memory (barrier.cluster + ld.shared::cluster),
// (3) normalise (+SiLU) from shared memory and store.  One read and one write of the tensor, one launch, no
// global atomics, no memset -- against stats kernel + memset + apply kernel (two reads, three launches).
constexpr int GN_CS = 8;
constexpr int GN_THREADS = 512;
__device__ __forceinline__ float ld_dsmem_f32(const float* p, uint32_t rank) {
    uint32_t a = (uint32_t)__cvta_generic_to_shared(p), ra;
    float v;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(a), "r"(rank));
    asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(ra) : "memory");
    return v;
}
__device__ __forceinline__ fix_t ld_dsmem_u64(const fix_t* p, uint32_t rank) {
    uint32_t a = (uint32_t)__cvta_generic_to_shared(p), ra;
    fix_t v;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(a), "r"(rank));
    asm volatile("ld.shared::cluster.u64 %0, [%1];" : "=l"(v) : "r"(ra) : "memory");
    return v;
}
__device__ __forceinline__ void cluster_sync_gn() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__global__ void __launch_bounds__(GN_THREADS)
gn_fused_kernel(const act_t* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                act_t* __restrict__ y, fix_t* __restrict__ stats_out, int HW, int C, int G, float eps, int do_silu,
                int rows_per_cta) {
    extern __shared__ __align__(16) uint8_t gsm[];
    __shared__ fix_t s_part[8];          // this CTA's fixed-point (sum, sumsq) of the slab's 4 groups
    __shared__ fix_t s_tot[8];           // cluster totals (private copy: peers keep reading s_part until the final barrier)
    __shared__ float s_ab[2 * 4];        // (mean, rstd) of the 4 groups
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const int slab = blockIdx.y, n = blockIdx.z;
    const int cpg = C / G, c8 = C / 8;
    const int sc8 = (4 * cpg) / 8;                           // 16-byte vectors per pixel in the slab
    const int cv0 = slab * sc8;                              // first vector column of the slab
    const int row0 = (int)rank * rows_per_cta, row1 = min(HW, row0 + rows_per_cta);
    const int nvec = max(0, row1 - row0) * sc8;
    if (threadIdx.x < 8) s_part[threadIdx.x] = 0ull;
    pdl_wait();
    pdl_trigger();
    __syncthreads();
    const h8* xp = reinterpret_cast<const h8*>(x + (size_t)n * HW * C);
    h8* sv = reinterpret_cast<h8*>(gsm);
    float s[4] = {0.f, 0.f, 0.f, 0.f}, ss[4] = {0.f, 0.f, 0.f, 0.f};
    auto accum = [&](const h8& v, int cv) {
        float f[8];
        unpack8(v, f);
        if (cpg % 8 == 0) {
            float a = 0.f, b = 0.f;
#pragma unroll
            for (int k = 0; k < 8; k++) { a += f[k]; b = fmaf(f[k], f[k], b); }
            const int g = (cv * 8) / cpg;
#pragma unroll
            for (int q = 0; q < 4; q++) if (g == q) { s[q] += a; ss[q] += b; }
        } else {
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const int g = (cv * 8 + k) / cpg;
#pragma unroll
                for (int q = 0; q < 4; q++) if (g == q) { s[q] += f[k]; ss[q] = fmaf(f[k], f[k], ss[q]); }
            }
        }
    };
    // 4 independent 16-byte loads in flight per thread (the block comes from L2: latency-, not bandwidth-bound)
    for (int i0 = threadIdx.x; i0 < nvec; i0 += 4 * GN_THREADS) {
        h8 v[4];
        int cvs[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int i = i0 + u * GN_THREADS;
            if (i < nvec) {
                const int r = i / sc8;
                cvs[u] = i - r * sc8;
                v[u] = xp[(size_t)(row0 + r) * c8 + cv0 + cvs[u]];
            }
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int i = i0 + u * GN_THREADS;
            if (i < nvec) { sv[i] = v[u]; accum(v[u], cvs[u]); }
        }
    }
#pragma unroll
    for (int q = 0; q < 4; q++) {
        float a = s[q], b = ss[q];
        for (int off = 16; off > 0; off >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, off); b += __shfl_xor_sync(0xffffffffu, b, off); }
        if ((threadIdx.x & 31) == 0) { atomicAdd(&s_part[2 * q], to_fix(a, kFixFwd)); atomicAdd(&s_part[2 * q + 1], to_fix(b, kFixFwd)); }
    }
    __syncthreads();
    cluster_sync_gn();                                       // every CTA's partials are complete and visible cluster-wide
    if (threadIdx.x < 8) {
        fix_t tot = 0ull;
#pragma unroll
        for (uint32_t r = 0; r < (uint32_t)GN_CS; r++) tot += ld_dsmem_u64(&s_part[threadIdx.x], r);
        if (stats_out && rank == 0) stats_out[((size_t)n * G + slab * 4) * 2 + threadIdx.x] = tot;
        s_tot[threadIdx.x] = tot;
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        const double inv = 1.0 / ((double)HW * (double)cpg) * (1.0 / 1048576.0);
        const double m = (double)(long long)s_tot[2 * threadIdx.x] * inv, q = (double)(long long)s_tot[2 * threadIdx.x + 1] * inv;
        s_ab[2 * threadIdx.x] = (float)m;
        s_ab[2 * threadIdx.x + 1] = rsqrtf((float)fmax(q - m * m, 0.0) + eps);
    }
    __syncthreads();
    // per-channel scale / shift of the slab -> shared memory (behind the data block)
    float* s_a = reinterpret_cast<float*>(gsm + (((size_t)rows_per_cta * sc8 * 16 + 15) & ~(size_t)15));
    float* s_b = s_a + 4 * cpg;
    const int c_slab0 = slab * 4 * cpg;
    {
        for (int cl = threadIdx.x; cl < 4 * cpg; cl += GN_THREADS) {
            const int g = cl / cpg;
            const float a = s_ab[2 * g + 1] * gamma[c_slab0 + cl];
            s_a[cl] = a;
            s_b[cl] = beta[c_slab0 + cl] - s_ab[2 * g] * a;
        }
    }
    __syncthreads();
    h8* yp = reinterpret_cast<h8*>(y + (size_t)n * HW * C);
    for (int i = threadIdx.x; i < nvec; i += GN_THREADS) {
        const int r = i / sc8, cv = i - r * sc8;
        float f[8];
        unpack8(sv[i], f);
        const float4 a0 = *reinterpret_cast<const float4*>(s_a + cv * 8), a1 = *reinterpret_cast<const float4*>(s_a + cv * 8 + 4);
        const float4 b0 = *reinterpret_cast<const float4*>(s_b + cv * 8), b1 = *reinterpret_cast<const float4*>(s_b + cv * 8 + 4);
        const float aa[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w}, bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const float o = fmaf(f[k], aa[k], bb[k]);
            f[k] = do_silu ? silu(o) : o;
        }
        yp[(size_t)(row0 + r) * c8 + cv0 + cv] = pack8(f);
    }
    cluster_sync_gn();                                       // peers may still be reading this CTA's partials
}

// Backward of y = act(GN(x)).  Both passes use the forward's work split (thread -> row lane x fixed
// 8-channel column, per-channel constants hoisted out of the row loop, 4 rows = 8 independent
// 16-byte loads in flight); with a = rstd*gamma, b = beta - mean*a:   z = a x + b,  dz = dy * act'(z).
// Pass 1: per-(n,g) sums of (gamma*dz) and (gamma*dz*xhat).
__global__ void __launch_bounds__(256)
gn_bwd_stats_kernel(const act_t* __restrict__ x, const act_t* __restrict__ dy, const fix_t* __restrict__ stats,
                    const float* __restrict__ gamma, const float* __restrict__ beta, fix_t* __restrict__ bstats,
                    int HW, int C, int G, float eps, int do_silu, int rows_per_cta) {
    pdl_wait();
    pdl_trigger();
    extern __shared__ fix_t s_bins[];
    const int n = blockIdx.y;
    const int cpg = C / G, c8 = C / 8;
    const double inv_cnt = 1.0 / ((double)HW * (double)cpg);
    for (int i = threadIdx.x; i < 2 * G; i += blockDim.x) s_bins[i] = 0ull;
    __syncthreads();
    const int row0 = blockIdx.x * rows_per_cta, row1 = min(HW, row0 + rows_per_cta);
    const h8* xp = reinterpret_cast<const h8*>(x + (size_t)n * HW * C);
    const h8* dp = reinterpret_cast<const h8*>(dy + (size_t)n * HW * C);
    const int rp = c8 <= 256 ? 256 / c8 : 1;
    const int rl = c8 <= 256 ? threadIdx.x / c8 : 0;
    for (int cv = c8 <= 256 ? threadIdx.x % c8 : threadIdx.x; cv < c8 && rl < rp; cv += 256) {
        float s1[8], s2[8], a[8], b[8], rs[8], mr[8], gm[8];
        int g_prev = -1;
        float mean = 0.f, rstd = 0.f;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int c = cv * 8 + i, g = c / cpg;
            s1[i] = 0.f; s2[i] = 0.f;
            if (g != g_prev) { gn_moments(stats, (size_t)n * G + g, inv_cnt, eps, mean, rstd); g_prev = g; }
            rs[i] = rstd; mr[i] = -mean * rs[i];
            gm[i] = gamma[c];
            a[i] = rs[i] * gm[i]; b[i] = beta[c] - mean * a[i];
        }
        auto accum = [&](const h8& xv, const h8& dv) {
            float f[8], d[8];
            unpack8(xv, f);
            unpack8(dv, d);
#pragma unroll
            for (int i = 0; i < 8; i++) {
                float dz = d[i];
                if (do_silu) dz *= silu_grad(fmaf(f[i], a[i], b[i]));
                const float gd = gm[i] * dz;
                s1[i] += gd; s2[i] = fmaf(gd, fmaf(f[i], rs[i], mr[i]), s2[i]);
            }
        };
        int r = row0 + rl;
        for (; r + 3 * rp < row1; r += 4 * rp) {
            h8 xv[4], dv[4];
#pragma unroll
            for (int u = 0; u < 4; u++) { xv[u] = xp[(size_t)(r + u * rp) * c8 + cv]; dv[u] = dp[(size_t)(r + u * rp) * c8 + cv]; }
#pragma unroll
            for (int u = 0; u < 4; u++) accum(xv[u], dv[u]);
        }
        for (; r < row1; r += rp) accum(xp[(size_t)r * c8 + cv], dp[(size_t)r * c8 + cv]);
        if (cpg % 8 == 0) {
            float u = 0.f, v = 0.f;
#pragma unroll
            for (int i = 0; i < 8; i++) { u += s1[i]; v += s2[i]; }
            const int g = (cv * 8) / cpg;
            atomicAdd(&s_bins[2 * g], to_fix(u, kFixBwd));
            atomicAdd(&s_bins[2 * g + 1], to_fix(v, kFixBwd));
        } else {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const int g = (cv * 8 + i) / cpg;
                atomicAdd(&s_bins[2 * g], to_fix(s1[i], kFixBwd));
                atomicAdd(&s_bins[2 * g + 1], to_fix(s2[i], kFixBwd));
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * G; i += blockDim.x) atomicAdd(&bstats[(size_t)n * 2 * G + i], s_bins[i]);
}

// Pass 2: dx = rstd * (gamma*dz - mean(gamma*dz) - xhat * mean(gamma*dz*xhat))  [+ dx_add]
//            = a*dz + k3*x + k4     with k3 = -rstd^2 m2,  k4 = rstd (rstd m2 mean - m1)
__device__ __forceinline__ uint4 gn_bwd_one(uint4 xv, uint4 dv, uint4 av, bool has_add, bool do_silu, const float (&a)[8],
                                            const float (&b)[8], const float (&k3)[8], const float (&k4)[8]) {
    const uint32_t xw[4] = {xv.x, xv.y, xv.z, xv.w}, dw[4] = {dv.x, dv.y, dv.z, dv.w}, aw[4] = {av.x, av.y, av.z, av.w};
    uint32_t ow[4];
#pragma unroll
    for (int h = 0; h < 4; h++) {
        const float2 f = act2_to_f2(*reinterpret_cast<const act2_t*>(&xw[h]));
        const float2 d = act2_to_f2(*reinterpret_cast<const act2_t*>(&dw[h]));
        const float2 e = act2_to_f2(*reinterpret_cast<const act2_t*>(&aw[h]));
        float dz0 = d.x, dz1 = d.y;
        if (do_silu) { dz0 *= silu_grad(fmaf(f.x, a[2 * h], b[2 * h])); dz1 *= silu_grad(fmaf(f.y, a[2 * h + 1], b[2 * h + 1])); }
        float o0 = fmaf(a[2 * h], dz0, fmaf(f.x, k3[2 * h], k4[2 * h]));
        float o1 = fmaf(a[2 * h + 1], dz1, fmaf(f.y, k3[2 * h + 1], k4[2 * h + 1]));
        if (has_add) { o0 += e.x; o1 += e.y; }
        const act2_t o = f2_to_act2(o0, o1);
        ow[h] = *reinterpret_cast<const uint32_t*>(&o);
    }
    return make_uint4(ow[0], ow[1], ow[2], ow[3]);
}

__global__ void __launch_bounds__(256)
gn_bwd_apply_kernel(const act_t* __restrict__ x, const act_t* __restrict__ dy, const fix_t* __restrict__ stats,
                    const fix_t* __restrict__ bstats, const float* __restrict__ gamma, const float* __restrict__ beta,
                    const act_t* __restrict__ dx_add, act_t* __restrict__ dx, int HW, int C, int G, float eps,
                    int do_silu, int rows_per_cta) {
    pdl_wait();
    pdl_trigger();
    const int n = blockIdx.y;
    const int c8 = C / 8, cpg = C / G;
    const double inv_cnt = 1.0 / ((double)HW * (double)cpg);
    const int row0 = blockIdx.x * rows_per_cta, row1 = min(HW, row0 + rows_per_cta);
    const size_t base = (size_t)n * HW * C;
    const uint4* xp = reinterpret_cast<const uint4*>(x + base);
    const uint4* dp = reinterpret_cast<const uint4*>(dy + base);
    const bool has_add = dx_add != nullptr;
    const uint4* ap = has_add ? reinterpret_cast<const uint4*>(dx_add + base) : xp;
    uint4* op = reinterpret_cast<uint4*>(dx + base);
    const int rp = c8 <= 256 ? 256 / c8 : 1;
    const int rl = c8 <= 256 ? threadIdx.x / c8 : 0;
    for (int cv = c8 <= 256 ? threadIdx.x % c8 : threadIdx.x; cv < c8 && rl < rp; cv += 256) {
        float a[8], b[8], k3[8], k4[8];
        int g_prev = -1;
        float mean = 0.f, rstd = 0.f, m1 = 0.f, m2 = 0.f;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int c = cv * 8 + i, g = c / cpg;
            if (g != g_prev) {
                gn_moments(stats, (size_t)n * G + g, inv_cnt, eps, mean, rstd);
                m1 = (float)from_fix(bstats[((size_t)n * G + g) * 2], inv_cnt * (1.0 / 68719476736.0));
                m2 = (float)from_fix(bstats[((size_t)n * G + g) * 2 + 1], inv_cnt * (1.0 / 68719476736.0));
                g_prev = g;
            }
            a[i] = rstd * gamma[c]; b[i] = beta[c] - mean * a[i];
            k3[i] = -rstd * rstd * m2; k4[i] = rstd * (rstd * m2 * mean - m1);
        }
        int r = row0 + rl;
        for (; r + 3 * rp < row1; r += 4 * rp) {
            const size_t o0 = (size_t)r * c8 + cv, o1 = o0 + (size_t)rp * c8, o2 = o1 + (size_t)rp * c8, o3 = o2 + (size_t)rp * c8;
            const uint4 x0 = xp[o0], x1 = xp[o1], x2 = xp[o2], x3 = xp[o3];
            const uint4 d0 = dp[o0], d1 = dp[o1], d2 = dp[o2], d3 = dp[o3];
            uint4 a0 = x0, a1 = x1, a2 = x2, a3 = x3;
            if (has_add) { a0 = ap[o0]; a1 = ap[o1]; a2 = ap[o2]; a3 = ap[o3]; }
            op[o0] = gn_bwd_one(x0, d0, a0, has_add, do_silu, a, b, k3, k4);
            op[o1] = gn_bwd_one(x1, d1, a1, has_add, do_silu, a, b, k3, k4);
            op[o2] = gn_bwd_one(x2, d2, a2, has_add, do_silu, a, b, k3, k4);
            op[o3] = gn_bwd_one(x3, d3, a3, has_add, do_silu, a, b, k3, k4);
        }
        for (; r < row1; r += rp) {
            const size_t o = (size_t)r * c8 + cv;
            const uint4 xv = xp[o];
            op[o] = gn_bwd_one(xv, dp[o], has_add ? ap[o] : xv, has_add, do_silu, a, b, k3, k4);
        }
    }
}

// ---------------------------------------------------------------------------- LayerNorm (one warp per row)
__global__ void __launch_bounds__(256)
layernorm_kernel(const act_t* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                 act_t* __restrict__ y, int64_t rows, int C, float eps) {
    pdl_wait();
    pdl_trigger();
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31, c8 = C / 8;
    const h8* xp = reinterpret_cast<const h8*>(x + row * C);
    h8* yp = reinterpret_cast<h8*>(y + row * C);
    constexpr int MAXV = 5;                            // C <= 1280 stays in registers (all loads issued up front)
    if (c8 <= MAXV * 32) {
        h8 v[MAXV];
#pragma unroll
        for (int u = 0; u < MAXV; u++) if (lane + u * 32 < c8) v[u] = xp[lane + u * 32];
        float s = 0.f, ss = 0.f;
#pragma unroll
        for (int u = 0; u < MAXV; u++) {
            if (lane + u * 32 < c8) {
                float f[8];
                unpack8(v[u], f);
#pragma unroll
                for (int i = 0; i < 8; i++) { s += f[i]; ss += f[i] * f[i]; }
            }
        }
        for (int off = 16; off > 0; off >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, off); ss += __shfl_xor_sync(0xffffffffu, ss, off); }
        const float mean = s / C, rstd = rsqrtf(fmaxf(ss / C - mean * mean, 0.f) + eps);
#pragma unroll
        for (int u = 0; u < MAXV; u++) {
            const int vv = lane + u * 32;
            if (vv < c8) {
                float f[8];
                unpack8(v[u], f);
                const float4 g0 = *reinterpret_cast<const float4*>(gamma + vv * 8), g1 = *reinterpret_cast<const float4*>(gamma + vv * 8 + 4);
                const float4 b0 = *reinterpret_cast<const float4*>(beta + vv * 8), b1 = *reinterpret_cast<const float4*>(beta + vv * 8 + 4);
                const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w}, bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                for (int i = 0; i < 8; i++) f[i] = (f[i] - mean) * rstd * gg[i] + bb[i];
                yp[vv] = pack8(f);
            }
        }
        return;
    }
    float s = 0.f, ss = 0.f;
    for (int v = lane; v < c8; v += 32) {
        float f[8];
        unpack8(xp[v], f);
#pragma unroll
        for (int i = 0; i < 8; i++) { s += f[i]; ss += f[i] * f[i]; }
    }
    for (int off = 16; off > 0; off >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, off); ss += __shfl_xor_sync(0xffffffffu, ss, off); }
    const float mean = s / C, rstd = rsqrtf(fmaxf(ss / C - mean * mean, 0.f) + eps);
    for (int v = lane; v < c8; v += 32) {
        float f[8];
        unpack8(xp[v], f);
#pragma unroll
        for (int i = 0; i < 8; i++) f[i] = (f[i] - mean) * rstd * gamma[v * 8 + i] + beta[v * 8 + i];
        yp[v] = pack8(f);
    }
}

// ---------------------------------------------------------------------------- row softmax (bf16, in place)
// rows x cols_pad (cols valid, the padding columns are written as 0); one CTA per row.
__global__ void __launch_bounds__(256)
softmax_kernel(act_t* __restrict__ s, int cols, int cols_pad) {
    pdl_wait();
    pdl_trigger();
    __shared__ float red[8];
    act_t* row = s + (size_t)blockIdx.x * cols_pad;
    const int c8 = cols_pad / 8;
    float mx = -INFINITY;
    for (int v = threadIdx.x; v < c8; v += blockDim.x) {
        float f[8];
        unpack8(reinterpret_cast<const h8*>(row)[v], f);
#pragma unroll
        for (int i = 0; i < 8; i++) if (v * 8 + i < cols) mx = fmaxf(mx, f[i]);
    }
    for (int off = 16; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    mx = red[0];
#pragma unroll
    for (int w = 1; w < 8; w++) mx = fmaxf(mx, red[w]);
    __syncthreads();
    float sum = 0.f;
    for (int v = threadIdx.x; v < c8; v += blockDim.x) {
        float f[8];
        unpack8(reinterpret_cast<const h8*>(row)[v], f);
#pragma unroll
        for (int i = 0; i < 8; i++) if (v * 8 + i < cols) sum += __expf(f[i] - mx);
    }
    for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
    __syncthreads();
    sum = 0.f;
#pragma unroll
    for (int w = 0; w < 8; w++) sum += red[w];
    const float inv = 1.0f / sum;
    for (int v = threadIdx.x; v < c8; v += blockDim.x) {
        float f[8];
        unpack8(reinterpret_cast<const h8*>(row)[v], f);
#pragma unroll
        for (int i = 0; i < 8; i++) f[i] = (v * 8 + i < cols) ? __expf(f[i] - mx) * inv : 0.f;
        reinterpret_cast<h8*>(row)[v] = pack8(f);
    }
}

// short rows (cols_pad <= 256): one warp per row, 8 rows per CTA
__global__ void __launch_bounds__(256)
softmax_warp_kernel(act_t* __restrict__ s, int64_t rows, int cols, int cols_pad) {
    pdl_wait();
    pdl_trigger();
    const int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (r >= rows) return;
    const int lane = threadIdx.x & 31;
    act_t* row = s + (size_t)r * cols_pad;
    const int c8 = cols_pad / 8;
    float f[8];
    const bool act = lane < c8;
    if (act) unpack8(reinterpret_cast<const h8*>(row)[lane], f);
    float mx = -INFINITY;
    if (act) {
#pragma unroll
        for (int i = 0; i < 8; i++) if (lane * 8 + i < cols) mx = fmaxf(mx, f[i]);
    }
    for (int off = 16; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
    float sum = 0.f;
    if (act) {
#pragma unroll
        for (int i = 0; i < 8; i++) { f[i] = (lane * 8 + i < cols) ? __expf(f[i] - mx) : 0.f; sum += f[i]; }
    }
    for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
    if (act) {
        const float inv = 1.0f / sum;
#pragma unroll
        for (int i = 0; i < 8; i++) f[i] *= inv;
        reinterpret_cast<h8*>(row)[lane] = pack8(f);
    }
}

// softmax backward in place on dP -> dS:  dS = P * (dP - sum(dP*P))      (VAE mid attention)
__global__ void __launch_bounds__(256)
softmax_bwd_kernel(const act_t* __restrict__ p, act_t* __restrict__ dp, int cols_pad) {
    pdl_wait();
    pdl_trigger();
    __shared__ float red[8];
    const act_t* pr = p + (size_t)blockIdx.x * cols_pad;
    act_t* dr = dp + (size_t)blockIdx.x * cols_pad;
    const int c8 = cols_pad / 8;
    float dot = 0.f;
    for (int v = threadIdx.x; v < c8; v += blockDim.x) {
        float a[8], b[8];
        unpack8(reinterpret_cast<const h8*>(pr)[v], a);
        unpack8(reinterpret_cast<const h8*>(dr)[v], b);
#pragma unroll
        for (int i = 0; i < 8; i++) dot += a[i] * b[i];
    }
    for (int off = 16; off > 0; off >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, off);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = dot;
    __syncthreads();
    dot = 0.f;
#pragma unroll
    for (int w = 0; w < 8; w++) dot += red[w];
    for (int v = threadIdx.x; v < c8; v += blockDim.x) {
        float a[8], b[8];
        unpack8(reinterpret_cast<const h8*>(pr)[v], a);
        unpack8(reinterpret_cast<const h8*>(dr)[v], b);
#pragma unroll
        for (int i = 0; i < 8; i++) b[i] = a[i] * (b[i] - dot);
        reinterpret_cast<h8*>(dr)[v] = pack8(b);
    }
}

// ---------------------------------------------------------------------------- GEGLU: y = a * gelu(b), x = [rows, 2*inner]
__global__ void __launch_bounds__(256)
geglu_kernel(const act_t* __restrict__ x, act_t* __restrict__ y, int64_t rows, int inner) {
    pdl_wait();
    pdl_trigger();
    const int i8 = inner / 8;
    const int64_t total = rows * i8;
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < total; v += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = v / i8;
        const int cv = (int)(v % i8);
        float a[8], b[8];
        unpack8(reinterpret_cast<const h8*>(x + r * 2 * inner)[cv], a);
        unpack8(reinterpret_cast<const h8*>(x + r * 2 * inner + inner)[cv], b);
#pragma unroll
        for (int i = 0; i < 8; i++) a[i] *= 0.5f * b[i] * (1.0f + erff(b[i] * 0.70710678118654752f));
        reinterpret_cast<h8*>(y + r * inner)[cv] = pack8(a);
    }
}

// ---------------------------------------------------------------------------- elementwise helpers
// mode 0: y = silu(x); mode 1: y = x + a; mode 2: dy * silu'(x)
__global__ void __launch_bounds__(256)
eltwise_kernel(const act_t* __restrict__ x, const act_t* __restrict__ a, act_t* __restrict__ y,
               int64_t n8, int mode) {
    pdl_wait();
    pdl_trigger();
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < n8; v += (int64_t)gridDim.x * blockDim.x) {
        float f[8], g[8];
        unpack8(reinterpret_cast<const h8*>(x)[v], f);
        if (mode != 0) unpack8(reinterpret_cast<const h8*>(a)[v], g);
#pragma unroll
        for (int i = 0; i < 8; i++) f[i] = mode == 0 ? silu(f[i]) : (mode == 1 ? f[i] + g[i] : g[i] * silu_grad(f[i]));
        reinterpret_cast<h8*>(y)[v] = pack8(f);
    }
}

// CFG + SDS gradient: grad = w * (eps_u + s (eps_c - eps_u) - noise)      (basic.py:595-603,642)
__global__ void __launch_bounds__(256)
sds_grad_kernel(const float* __restrict__ eps_uncond, const float* __restrict__ eps_cond, const float* __restrict__ noise,
                float* __restrict__ grad, float* __restrict__ noise_pred, float guidance_scale, float weight, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float np = eps_uncond[i] + guidance_scale * (eps_cond[i] - eps_uncond[i]);
        if (noise_pred) noise_pred[i] = np;
        grad[i] = weight * (np - noise[i]);
    }
}

static inline unsigned grid_for(int64_t work, int threads) {
    const int64_t g = (work + threads - 1) / threads;
    return (unsigned)(g < 8 * kNumSMs ? (g > 0 ? g : 1) : 8 * kNumSMs);
}

}  // namespace nn
}  // namespace dwg

using namespace dwg;
using namespace dwg::nn;
typedef act_t h16;

static inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
static int g_gn_last_launches = 2;
// The one-launch cluster GroupNorm is correct (tests) but measured SLOWER inside the step than stats + apply
// (15.9 vs 15.3 ms/step: 8-CTA clusters with up to 123 KB of shared memory start late behind the persistent GEMM
// CTAs, the small two-pass CTAs slip in early under programmatic dependent launch): off by default.
static int g_gn_fused = 0;
extern "C" int dwg_groupnorm_set_fused(int on) { g_gn_fused = on ? 1 : 0; return DWG_OK; }
/* kernels the last dwg_groupnorm_fwd call launched: 1 (one-launch cluster kernel) or 2 (stats + apply) */
extern "C" int dwg_groupnorm_last_launches(void) { return g_gn_last_launches; }

/* Shared-memory carve-out preference of the streaming kernels (percent of the L1/shared array, -1 = driver default).
 * The tcgen05 GEMM / attention kernels need the maximum carve-out; when the small kernels between them ask for the
 * default one, every GEMM <-> norm alternation re-partitions the SM's L1/shared memory (tools/carveout_probe.py). */
extern "C" int dwg_nn_set_carveout(int percent) {
    const int v = percent < 0 ? (int)cudaSharedmemCarveoutDefault : (percent > 100 ? 100 : percent);
    cudaFuncSetAttribute(gn_stats_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, v);
    cudaFuncSetAttribute(gn_apply_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, v);
    cudaFuncSetAttribute(gn_apply_cs_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, v);
    cudaFuncSetAttribute(gn_bwd_stats_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, v);
    cudaFuncSetAttribute(gn_bwd_apply_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, v);
    cudaFuncSetAttribute(layernorm_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, v);
    cudaFuncSetAttribute(softmax_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, v);
    cudaFuncSetAttribute(softmax_warp_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, v);
    cudaFuncSetAttribute(softmax_bwd_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, v);
    cudaFuncSetAttribute(geglu_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, v);
    cudaFuncSetAttribute(eltwise_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, v);
    cudaFuncSetAttribute(sds_grad_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, v);
    return check_launch("dwg_nn_set_carveout");
}

extern "C" int dwg_groupnorm_fwd(const void* x, const float* gamma, const float* beta, void* y, void* stats_,
                                 int N, int HW, int C, int G, float eps, int do_silu, void* stream) {
    fix_t* stats = reinterpret_cast<fix_t*>(stats_);
    DWG_REQUIRE(x && gamma && beta && y && stats, "null pointer");
    DWG_REQUIRE(C % 8 == 0 && C % G == 0 && al16(x) && al16(y) && (reinterpret_cast<uintptr_t>(stats) & 7) == 0, "C must be a multiple of 8 and of G; 16-byte aligned tensors");
    cudaStream_t st = (cudaStream_t)stream;
    // one-launch cluster kernel when the (image, 4-group slab) block fits the shared memory of 8 CTAs
    {
        const int cpg = C / G;
        const int rows_c = (HW + GN_CS - 1) / GN_CS;
        const size_t smem = (((size_t)rows_c * (size_t)(4 * cpg) * 2 + 15) & ~(size_t)15) + (size_t)(8 * cpg) * 4;
        static bool attr_done = false;
        if (g_gn_fused && !attr_done) {
            cudaFuncSetAttribute(gn_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
            attr_done = true;
        }
        if (g_gn_fused && (G % 4) == 0 && ((4 * cpg) % 8) == 0 && smem <= 200 * 1024 && (int64_t)N * (G / 4) * GN_CS <= 4 * kNumSMs) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(GN_CS, G / 4, N); cfg.blockDim = dim3(GN_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
            cudaLaunchAttribute attr[2];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = GN_CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
            attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attr[1].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 2 : 1;
            cudaLaunchKernelEx(&cfg, gn_fused_kernel, (const h16*)x, gamma, beta, (h16*)y, stats, HW, C, G, eps, do_silu & 1, rows_c);
            g_gn_last_launches = 1;
            return check_launch("dwg_groupnorm_fwd (fused)");
        }
    }
    g_gn_last_launches = 2;
    if (!(do_silu & 2)) cudaMemsetAsync(stats, 0, sizeof(fix_t) * 2 * N * G, st);      // bit 1: the caller pre-zeroed the statistics
    const int rp_ = (C / 8) <= 256 ? 256 / (C / 8) : 1;        // row lanes per CTA
    int rows_per_cta = (int)(((int64_t)N * HW + 4 * kNumSMs - 1) / (4 * kNumSMs));      // ~4 CTAs per SM
    if (rows_per_cta < 4 * rp_) rows_per_cta = 4 * rp_;
    if (rows_per_cta > 64 * rp_) rows_per_cta = 64 * rp_;
    dim3 grid((HW + rows_per_cta - 1) / rows_per_cta, N);
    launch_pdl(gn_stats_kernel, grid, dim3(256), sizeof(fix_t) * 2 * G, st, (const h16*)x, stats, HW, C, G, rows_per_cta);
    int rows_apply = rows_per_cta;
    dim3 grid2((HW + rows_apply - 1) / rows_apply, N);
    launch_pdl(gn_apply_kernel, grid2, dim3(256), 0, st, (const h16*)x, (const fix_t*)stats, gamma, beta, (h16*)y, HW, C, G, eps, do_silu & 1, rows_apply);
    return check_launch("dwg_groupnorm_fwd");
}

static int gn_apply_cs_impl(const void* x, const void* x2, int C1, const float* gamma, const float* beta, void* y, const void* colstats,
                            const void* colstats2, void* stats_out, int N, int HW, int C, int G, float eps, int do_silu, void* stream, const char* what) {
    DWG_REQUIRE(x && gamma && beta && y && colstats, "null pointer");
    DWG_REQUIRE(C % 8 == 0 && C % G == 0 && G <= 32 && al16(x) && al16(y), "C must be a multiple of 8 and of G, G <= 32; 16-byte aligned tensors");
    DWG_REQUIRE(!x2 || (colstats2 && C1 > 0 && C1 < C && C1 % 8 == 0 && al16(x2)), "two-source form: C1 in (0, C), a multiple of 8, second statistics given");
    cudaStream_t st = (cudaStream_t)stream;
    const int rp_ = (C / 8) <= 256 ? 256 / (C / 8) : 1;
    int rows_per_cta = (int)(((int64_t)N * HW + 4 * kNumSMs - 1) / (4 * kNumSMs));
    if (rows_per_cta < 4 * rp_) rows_per_cta = 4 * rp_;
    if (rows_per_cta > 64 * rp_) rows_per_cta = 64 * rp_;
    dim3 grid((HW + rows_per_cta - 1) / rows_per_cta, N);
    launch_pdl(gn_apply_cs_kernel, grid, dim3(256), 0, st, (const h16*)x, (const fix_t*)colstats, gamma, beta, (h16*)y, (fix_t*)stats_out, HW, C, G, eps,
               do_silu & 1, rows_per_cta, (const h16*)x2, (const fix_t*)colstats2, x2 ? C1 : C);
    return check_launch(what);
}
extern "C" int dwg_groupnorm_apply_cs(const void* x, const float* gamma, const float* beta, void* y, const void* colstats, void* stats_out,
                                      int N, int HW, int C, int G, float eps, int do_silu, void* stream) {
    return gn_apply_cs_impl(x, nullptr, C, gamma, beta, y, colstats, nullptr, stats_out, N, HW, C, G, eps, do_silu, stream, "dwg_groupnorm_apply_cs");
}
/* GroupNorm(+SiLU) of the channel concatenation [x1 (C1 channels) | x2 (C - C1 channels)] without materialising it: each source comes
 * with its own column statistics (see include/dwg.h). */
extern "C" int dwg_groupnorm_apply_cs2(const void* x1, const void* x2, int C1, const float* gamma, const float* beta, void* y, const void* colstats1,
                                       const void* colstats2, void* stats_out, int N, int HW, int C, int G, float eps, int do_silu, void* stream) {
    DWG_REQUIRE(x2 && colstats2, "null pointer");
    return gn_apply_cs_impl(x1, x2, C1, gamma, beta, y, colstats1, colstats2, stats_out, N, HW, C, G, eps, do_silu, stream, "dwg_groupnorm_apply_cs2");
}

extern "C" int dwg_groupnorm_bwd(const void* x, const void* dy, const void* stats_, const float* gamma, const float* beta,
                                 const void* dx_add, void* dx, void* bstats_, int N, int HW, int C, int G, float eps,
                                 int do_silu, void* stream) {
    const fix_t* stats = reinterpret_cast<const fix_t*>(stats_);
    fix_t* bstats = reinterpret_cast<fix_t*>(bstats_);
    DWG_REQUIRE(x && dy && stats && gamma && beta && dx && bstats, "null pointer");
    DWG_REQUIRE(C % 8 == 0 && C % G == 0, "C must be a multiple of 8 and of G");
    cudaStream_t st = (cudaStream_t)stream;
    if (!(do_silu & 2)) cudaMemsetAsync(bstats, 0, sizeof(fix_t) * 2 * N * G, st);
    const int rp_ = (C / 8) <= 256 ? 256 / (C / 8) : 1;        // row lanes per CTA
    int rows_per_cta = (int)(((int64_t)N * HW + 6 * kNumSMs - 1) / (6 * kNumSMs));      // ~6 CTAs per SM
    if (rows_per_cta < 4 * rp_) rows_per_cta = 4 * rp_;
    if (rows_per_cta > 64 * rp_) rows_per_cta = 64 * rp_;
    dim3 grid((HW + rows_per_cta - 1) / rows_per_cta, N);
    launch_pdl(gn_bwd_stats_kernel, grid, dim3(256), sizeof(fix_t) * 2 * G, st, (const h16*)x, (const h16*)dy, stats, gamma, beta, bstats,
               HW, C, G, eps, do_silu & 1, rows_per_cta);
    launch_pdl(gn_bwd_apply_kernel, grid, dim3(256), 0, st, (const h16*)x, (const h16*)dy, stats, (const fix_t*)bstats, gamma, beta,
               (const h16*)dx_add, (h16*)dx, HW, C, G, eps, do_silu & 1, rows_per_cta);
    return check_launch("dwg_groupnorm_bwd");
}

extern "C" int dwg_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y, int64_t rows, int C, float eps, void* stream) {
    DWG_REQUIRE(x && gamma && beta && y && C % 8 == 0, "bad arguments");
    launch_pdl(layernorm_kernel, dim3((unsigned)((rows + 7) / 8)), dim3(256), 0, (cudaStream_t)stream, (const h16*)x, gamma, beta, (h16*)y, rows, C, eps);
    return check_launch("dwg_layernorm_fwd");
}

extern "C" int dwg_softmax_rows(void* s, int64_t rows, int cols, int cols_pad, void* stream) {
    DWG_REQUIRE(s && cols > 0 && cols_pad >= cols && cols_pad % 8 == 0 && rows < (1ll << 31), "bad arguments");
    if (cols_pad <= 256)
        softmax_warp_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>((h16*)s, rows, cols, cols_pad);
    else
        softmax_kernel<<<(unsigned)rows, 256, 0, (cudaStream_t)stream>>>((h16*)s, cols, cols_pad);
    return check_launch("dwg_softmax_rows");
}

extern "C" int dwg_softmax_rows_bwd(const void* p, void* dp, int64_t rows, int cols_pad, void* stream) {
    DWG_REQUIRE(p && dp && cols_pad % 8 == 0 && rows < (1ll << 31), "bad arguments");
    softmax_bwd_kernel<<<(unsigned)rows, 256, 0, (cudaStream_t)stream>>>((const h16*)p, (h16*)dp, cols_pad);
    return check_launch("dwg_softmax_rows_bwd");
}

extern "C" int dwg_geglu(const void* x, void* y, int64_t rows, int inner, void* stream) {
    DWG_REQUIRE(x && y && inner % 8 == 0, "bad arguments");
    launch_pdl(geglu_kernel, dim3(grid_for(rows * (inner / 8), 256)), dim3(256), 0, (cudaStream_t)stream, (const h16*)x, (h16*)y, rows, inner);
    return check_launch("dwg_geglu");
}

// ---------------------------------------------------------------------------- layout boundary kernels
// The network boundaries of the diffusion path change layout AND type: planar fp32 (images, latents: the reference's NCHW tensors)
// <-> channels-last fp16 padded to 8 channels (the tcgen05 convolutions' operand).  One kernel each way instead of
// permute + pad + cast (+ scale) as three or four torch kernels.
__global__ void __launch_bounds__(256)
nchw_to_nhwc8_kernel(const float* __restrict__ src, act_t* __restrict__ dst, int C, int64_t HW, int Cp, float scale, float shift, int64_t total) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;          // one thread = one pixel's 8-channel vector
    if (i >= total) return;
    const int v8 = Cp / 8;
    const int64_t pix = i / v8;
    const int cv = (int)(i - pix * v8);
    const int64_t n = pix / HW, p = pix - n * HW;
    float f[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const int c = cv * 8 + k;
        f[k] = c < C ? fmaf(src[(n * C + c) * HW + p], scale, shift) : 0.f;
    }
    reinterpret_cast<h8*>(dst)[i] = pack8(f);
}
__global__ void __launch_bounds__(256)
nhwc_to_nchw_kernel(const act_t* __restrict__ src, float* __restrict__ dst, int C, int64_t HW, int Cp, float scale, int64_t total) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;          // one thread = one pixel (C <= 8 valid channels of the first vector)
    if (i >= total) return;
    const int64_t n = i / HW, p = i - n * HW;
    if ((Cp & 7) == 0) {
        float f[8];
        unpack8(reinterpret_cast<const h8*>(src + i * Cp)[0], f);
#pragma unroll
        for (int k = 0; k < 8; k++)
            if (k < C) dst[(n * C + k) * HW + p] = f[k] * scale;
    } else {                                                     // unpadded rows (a 3-channel dgrad output): scalar reads
        for (int k = 0; k < C; k++) dst[(n * C + k) * HW + p] = act_to_f(src[i * Cp + k]) * scale;
    }
}
/* dst [N,H,W,Cp] fp16 (Cp = C rounded up to 8, padding channels zero) = scale * src [N,C,H,W] fp32 + shift */
extern "C" int dwg_nchw_f32_to_nhwc_f16(const float* src, void* dst, int N, int C, int64_t HW, int Cp, float scale, float shift, void* stream) {
    DWG_REQUIRE(src && dst && N > 0 && C > 0 && HW > 0 && Cp % 8 == 0 && Cp >= C && al16(dst), "bad arguments");
    const int64_t total = (int64_t)N * HW * (Cp / 8);
    nchw_to_nhwc8_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(src, (h16*)dst, C, HW, Cp, scale, shift, total);
    return check_launch("dwg_nchw_f32_to_nhwc_f16");
}
/* dst [N,C,H,W] fp32 = scale * src [N,H,W,Cp][..., :C] (C <= 8) */
extern "C" int dwg_nhwc_f16_to_nchw_f32(const void* src, float* dst, int N, int C, int64_t HW, int Cp, float scale, void* stream) {
    DWG_REQUIRE(src && dst && N > 0 && C > 0 && C <= 8 && HW > 0 && Cp >= C && (Cp % 8 != 0 || al16(src)), "bad arguments");
    const int64_t total = (int64_t)N * HW;
    nhwc_to_nchw_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const h16*)src, dst, C, HW, Cp, scale, total);
    return check_launch("dwg_nhwc_f16_to_nchw_f32");
}

extern "C" int dwg_eltwise_f16(const void* x, const void* a, void* y, int64_t n, int mode, void* stream) {
    DWG_REQUIRE(x && y && n % 8 == 0 && mode >= 0 && mode <= 2 && (mode == 0 || a), "bad arguments");
    launch_pdl(eltwise_kernel, dim3(grid_for(n / 8, 256)), dim3(256), 0, (cudaStream_t)stream, (const h16*)x, (const h16*)a, (h16*)y, n / 8, mode);
    return check_launch("dwg_eltwise_f16");
}

extern "C" int dwg_sds_grad(const float* eps_uncond, const float* eps_cond, const float* noise, float* grad, float* noise_pred,
                            float guidance_scale, float weight, int64_t n, void* stream) {
    DWG_REQUIRE(eps_uncond && eps_cond && noise && grad, "null pointer");
    sds_grad_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(eps_uncond, eps_cond, noise, grad, noise_pred, guidance_scale, weight, n);
    return check_launch("dwg_sds_grad");
}
