// Fused multi-head attention forward on tcgen05 (R15: every self/cross attention of the UNet and
// ControlNet transformer blocks).  O = softmax(Q K^T / sqrt(d)) V without ever writing the
// [T, Tk] score matrix to HBM (the unfused path moved ~2 GB per 64x64-resolution layer).
//
// One CTA = 128 query rows of one (batch, head).  Per 128-key block:
//   warp 4 (one elected lane): TMA loads of K / V^T (double-buffered, 128B-swizzled), then
//       S = Q K^T        tcgen05.mma  M=128 N=128 K=64*chunks   -> TMEM columns [0,128)
//   warps 0-3 (thread == query row, TMEM lane == row): online softmax straight out of TMEM
//       (tcgen05.ld), P written as bf16 into a swizzled shared-memory A-operand tile
//   warp 4:  O_blk = P V  tcgen05.mma  M=128 N=round16(d) K=128 -> TMEM columns [128, 128+N)
//   warps 0-3: O = alpha * O + O_blk in registers (fp32); final O / l -> bf16 -> HBM.
// V is consumed as V^T ([d, Tk], key index contiguous = K-major B operand), which the projection
// GEMM produces for free by swapping its operands (see dwg/diffusion/model.py).
#include <cuda.h>
#include <cuda_bf16.h>

#include "common.cuh"

namespace dwg {
namespace attn {

constexpr int BQ = 128, BKV = 128;

struct Params {
    int T, Tk, heads, hd;
    float scale_log2;               // softmax scale * log2(e)
    __nv_bfloat16* out;             // [B, T, heads*hd]
    int64_t out_row_stride;         // = heads*hd
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\t"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 :: "r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ uint32_t make_idesc(int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(BQ >> 4) << 24);
}

template <int CH /* ceil(hd/64) */, int NPV /* round16(hd) */>
__global__ void __launch_bounds__(160, (CH == 1 && NPV <= 64) ? 2 : 1)
fa_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
              const __grid_constant__ CUtensorMap tmV, const Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    constexpr uint32_t TILE = 128 * 128;                     // [128 rows][64 bf16], 128B-swizzled
    constexpr uint32_t VT = NPV * 128;                       // [NPV rows][64 keys]
    uint8_t* sQ = smem;                                      // CH tiles
    uint8_t* sK = sQ + CH * TILE;                            // 2 stages x CH tiles
    uint8_t* sV = sK + 2 * CH * TILE;                        // 2 stages x 2 key-chunks x VT
    uint8_t* sP = sV + 2 * 2 * VT;                           // 2 key-chunks x TILE
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * TILE);
    uint64_t* q_full = bars; uint64_t* kv_full = bars + 1; uint64_t* kv_empty = bars + 3;
    uint64_t* s_full = bars + 5; uint64_t* p_full = bars + 6; uint64_t* o_full = bars + 7;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

    const int warp = threadIdx.x >> 5;
    const int q0 = blockIdx.x * BQ, head = blockIdx.y, b = blockIdx.z;
    const int nblk = (p.Tk + BKV - 1) / BKV;

    if (warp == 4) {
        if (elect_one()) {
            asm volatile("prefetch.tensormap [%0];" :: "l"(&tmQ) : "memory");
            asm volatile("prefetch.tensormap [%0];" :: "l"(&tmK) : "memory");
            asm volatile("prefetch.tensormap [%0];" :: "l"(&tmV) : "memory");
            mbar_init(q_full, 1);
            mbar_init(&kv_full[0], 1); mbar_init(&kv_full[1], 1);
            mbar_init(&kv_empty[0], 1); mbar_init(&kv_empty[1], 1);
            mbar_init(s_full, 1); mbar_init(p_full, 128); mbar_init(o_full, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "n"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_S = *tmem_slot, tmem_O = tmem_S + 128;
    pdl_wait();
    pdl_trigger();

    if (warp == 4) {
        if (elect_one()) {
            constexpr uint32_t KV_BYTES = CH * TILE + 2 * VT;
            auto load_kv = [&](int j, int s) {
                mbar_expect_tx(&kv_full[s], KV_BYTES);
#pragma unroll
                for (int c = 0; c < CH; c++) tma_load_4d(sK + (s * CH + c) * TILE, &tmK, &kv_full[s], c * 64, j * BKV, head, b);
#pragma unroll
                for (int kc = 0; kc < 2; kc++) tma_load_4d(sV + (s * 2 + kc) * VT, &tmV, &kv_full[s], j * BKV + kc * 64, 0, head, b);
            };
            mbar_expect_tx(q_full, CH * TILE);
#pragma unroll
            for (int c = 0; c < CH; c++) tma_load_4d(sQ + c * TILE, &tmQ, q_full, c * 64, q0, head, b);
            load_kv(0, 0);
            const uint32_t idS = make_idesc(128), idO = make_idesc(NPV);
            for (int j = 0; j < nblk; j++) {
                const int s = j & 1;
                if (j + 1 < nblk) {
                    mbar_wait(&kv_empty[s ^ 1], (uint32_t)(((j + 1) >> 1) & 1) ^ 1u);
                    load_kv(j + 1, s ^ 1);
                }
                if (j == 0) mbar_wait(q_full, 0);
                mbar_wait(&kv_full[s], (uint32_t)((j >> 1) & 1));
                tc_fence_after();
#pragma unroll
                for (int c = 0; c < CH; c++) {
                    const uint64_t dq = make_desc(smem_u32(sQ + c * TILE));
                    const uint64_t dk = make_desc(smem_u32(sK + (s * CH + c) * TILE));
#pragma unroll
                    for (int k = 0; k < 4; k++) tc_mma(tmem_S, dq + (uint64_t)(k * 2), dk + (uint64_t)(k * 2), idS, (c | k) != 0 ? 1u : 0u);
                }
                tc_commit(s_full);
                mbar_wait(p_full, (uint32_t)(j & 1));          // P is in shared memory, S has been consumed
                tc_fence_after();
#pragma unroll
                for (int kc = 0; kc < 2; kc++) {
                    const uint64_t dp = make_desc(smem_u32(sP + kc * TILE));
                    const uint64_t dv = make_desc(smem_u32(sV + (s * 2 + kc) * VT));
#pragma unroll
                    for (int k = 0; k < 4; k++) tc_mma(tmem_O, dp + (uint64_t)(k * 2), dv + (uint64_t)(k * 2), idO, (kc | k) != 0 ? 1u : 0u);
                }
                tc_commit(o_full);
                tc_commit(&kv_empty[s]);
            }
        }
        __syncwarp();
    } else {
        const int r = threadIdx.x;                            // query row of the tile == TMEM lane
        const uint32_t lane_base = ((uint32_t)(warp * 32)) << 16;
        float m = -INFINITY, l = 0.f;
        float O[NPV];
#pragma unroll
        for (int c = 0; c < NPV; c++) O[c] = 0.f;
        for (int j = 0; j < nblk; j++) {
            mbar_wait(s_full, (uint32_t)(j & 1));
            tc_fence_after();
            const int kv_valid = min(BKV, p.Tk - j * BKV);
            // pass 1: row maximum
            float mx = -INFINITY;
#pragma unroll 2
            for (int c0 = 0; c0 < BKV; c0 += 16) {
                uint32_t v[16];
                tc_ld16(tmem_S + lane_base + (uint32_t)c0, v);
#pragma unroll
                for (int i = 0; i < 16; i++) if (c0 + i < kv_valid) mx = fmaxf(mx, __uint_as_float(v[i]));
            }
            const float m_new = fmaxf(m, mx * p.scale_log2);
            const float alpha = exp2f(m - m_new);               // 0 on the first block (m = -inf)
            float lsum = 0.f;
            // pass 2: P = exp2(s - m_new) -> bf16 -> swizzled smem tile
#pragma unroll 2
            for (int c0 = 0; c0 < BKV; c0 += 16) {
                uint32_t v[16];
                tc_ld16(tmem_S + lane_base + (uint32_t)c0, v);
                uint32_t pk[8];
#pragma unroll
                for (int i = 0; i < 16; i += 2) {
                    const float p0 = (c0 + i < kv_valid) ? exp2f(__uint_as_float(v[i]) * p.scale_log2 - m_new) : 0.f;
                    const float p1 = (c0 + i + 1 < kv_valid) ? exp2f(__uint_as_float(v[i + 1]) * p.scale_log2 - m_new) : 0.f;
                    const __nv_bfloat162 h = __floats2bfloat162_rn(p0, p1);
                    // accumulate the ROUNDED probabilities so that l matches what the PV matmul sees
                    const float2 hf = __bfloat1622float2(h);
                    lsum += hf.x + hf.y;
                    pk[i >> 1] = *reinterpret_cast<const uint32_t*>(&h);
                }
                // 16 columns = two 16-byte chunks of the 128-byte row; chunk index XOR (row & 7)
                const int kc = c0 >> 6;                         // which 64-key tile
                const int ch = (c0 & 63) >> 3;                  // chunk (8 bf16) inside the row
                uint8_t* rowp = sP + kc * TILE + r * 128;
                *reinterpret_cast<uint4*>(rowp + (((ch) ^ (r & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                *reinterpret_cast<uint4*>(rowp + (((ch + 1) ^ (r & 7)) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
            }
            l = l * alpha + lsum;
            m = m_new;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> async proxy (MMA)
            tc_fence_before();
            mbar_arrive(p_full);
            // O = alpha * O + P V
            mbar_wait(o_full, (uint32_t)(j & 1));
            tc_fence_after();
#pragma unroll
            for (int c0 = 0; c0 < NPV; c0 += 16) {
                uint32_t v[16];
                tc_ld16(tmem_O + lane_base + (uint32_t)c0, v);
#pragma unroll
                for (int i = 0; i < 16; i++) O[c0 + i] = O[c0 + i] * alpha + __uint_as_float(v[i]);
            }
            tc_fence_before();
        }
        const int t = q0 + r;
        if (t < p.T) {
            const float inv = 1.0f / l;
            __nv_bfloat16* dst = p.out + ((int64_t)b * p.T + t) * p.out_row_stride + (int64_t)head * p.hd;
#pragma unroll
            for (int c0 = 0; c0 < NPV; c0 += 8) {
                if (c0 < p.hd) {
                    uint4 pk;
                    __nv_bfloat162 h0 = __floats2bfloat162_rn(O[c0] * inv, O[c0 + 1] * inv), h1 = __floats2bfloat162_rn(O[c0 + 2] * inv, O[c0 + 3] * inv);
                    __nv_bfloat162 h2 = __floats2bfloat162_rn(O[c0 + 4] * inv, O[c0 + 5] * inv), h3 = __floats2bfloat162_rn(O[c0 + 6] * inv, O[c0 + 7] * inv);
                    pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
                    pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
                    *reinterpret_cast<uint4*>(dst + c0) = pk;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_S), "n"(256));
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) != cudaSuccess || !sym) return nullptr;
        fn = reinterpret_cast<EncodeTiledFn>(sym);
    }
    return fn;
}
static int make_map(CUtensorMap* m, const void* base, const uint64_t d[4], const uint64_t s[3], const uint32_t box[4]) {
    EncodeTiledFn enc = get_encode();
    if (!enc) { set_error("cuTensorMapEncodeTiled unavailable"); return DWG_ERR_CUDA; }
    cuuint64_t dd[4] = {d[0], d[1], d[2], d[3]};
    cuuint64_t ss[3] = {s[0], s[1], s[2]};
    cuuint32_t bb[4] = {box[0], box[1], box[2], box[3]};
    cuuint32_t ee[4] = {1, 1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dd, ss, bb, ee, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("attention: cuTensorMapEncodeTiled failed (%d)", (int)r); return DWG_ERR_INVALID; }
    return DWG_OK;
}

template <int CH, int NPV>
static int launch(const CUtensorMap& q, const CUtensorMap& k, const CUtensorMap& v, const Params& p, dim3 grid, cudaStream_t st) {
    const size_t smem = 1024 + (size_t)CH * 16384 + 2 * CH * 16384 + 4 * NPV * 128 + 2 * 16384 + 128;
    cudaFuncSetAttribute(fa_fwd_kernel<CH, NPV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    launch_pdl(fa_fwd_kernel<CH, NPV>, grid, dim3(160), smem, st, q, k, v, p);
    return check_launch("dwg_attention_fwd");
}

}  // namespace attn
}  // namespace dwg

using namespace dwg;
using namespace dwg::attn;

// q [B,T,heads*hd] (row stride q_ld), k [B,Tk,heads*hd] (row stride k_ld), vt [B, heads*hd, Tkp] (V transposed,
// batch stride vt_batch_stride elements, columns >= Tk must be finite), out [B,T,heads*hd].  All bf16.  hd in {40, 64, 80, 128} (multiple of 8, <= 128).
extern "C" int dwg_attention_fwd(const void* q, int64_t q_ld, const void* k, int64_t k_ld, const void* vt, int64_t Tkp,
                                 int64_t vt_batch_stride, void* out, int B, int heads, int T, int Tk, int hd, float scale, void* stream) {
    DWG_REQUIRE(q && k && vt && out, "null pointer");
    DWG_REQUIRE(hd % 8 == 0 && hd >= 8 && hd <= 128, "head dim must be a multiple of 8 and <= 128");
    DWG_REQUIRE(q_ld % 8 == 0 && k_ld % 8 == 0 && Tkp % 8 == 0 && Tkp >= Tk && vt_batch_stride % 8 == 0, "strides must be multiples of 8 elements");
    const int C = heads * hd;
    CUtensorMap tq, tk, tv;
    {
        const uint64_t d[4] = {(uint64_t)hd, (uint64_t)T, (uint64_t)heads, (uint64_t)B};
        const uint64_t s[3] = {(uint64_t)q_ld * 2, (uint64_t)hd * 2, (uint64_t)T * q_ld * 2};
        const uint32_t box[4] = {64, 128, 1, 1};
        int rc = make_map(&tq, q, d, s, box); if (rc) return rc;
    }
    {
        const uint64_t d[4] = {(uint64_t)hd, (uint64_t)Tk, (uint64_t)heads, (uint64_t)B};
        const uint64_t s[3] = {(uint64_t)k_ld * 2, (uint64_t)hd * 2, (uint64_t)Tk * k_ld * 2};
        const uint32_t box[4] = {64, 128, 1, 1};
        int rc = make_map(&tk, k, d, s, box); if (rc) return rc;
    }
    const int npv = (hd + 15) / 16 * 16;
    {
        const uint64_t d[4] = {(uint64_t)Tkp, (uint64_t)hd, (uint64_t)heads, (uint64_t)B};
        const uint64_t s[3] = {(uint64_t)Tkp * 2, (uint64_t)hd * Tkp * 2, (uint64_t)vt_batch_stride * 2};
        const uint32_t box[4] = {64, (uint32_t)npv, 1, 1};
        int rc = make_map(&tv, vt, d, s, box); if (rc) return rc;
    }
    Params p;
    p.T = T; p.Tk = Tk; p.heads = heads; p.hd = hd; p.scale_log2 = scale * 1.4426950408889634f;
    p.out = reinterpret_cast<__nv_bfloat16*>(out); p.out_row_stride = C;
    dim3 grid((T + BQ - 1) / BQ, heads, B);
    cudaStream_t st = (cudaStream_t)stream;
    const int ch = (hd + 63) / 64;
    if (ch == 1 && npv == 16) return launch<1, 16>(tq, tk, tv, p, grid, st);
    if (ch == 1 && npv == 32) return launch<1, 32>(tq, tk, tv, p, grid, st);
    if (ch == 1 && npv == 48) return launch<1, 48>(tq, tk, tv, p, grid, st);
    if (ch == 1 && npv == 64) return launch<1, 64>(tq, tk, tv, p, grid, st);
    if (ch == 2 && npv == 80) return launch<2, 80>(tq, tk, tv, p, grid, st);
    if (ch == 2 && npv == 96) return launch<2, 96>(tq, tk, tv, p, grid, st);
    if (ch == 2 && npv == 112) return launch<2, 112>(tq, tk, tv, p, grid, st);
    if (ch == 2 && npv == 128) return launch<2, 128>(tq, tk, tv, p, grid, st);
    set_error("dwg_attention_fwd: unsupported head dim %d", hd);
    return DWG_ERR_INVALID;
}
