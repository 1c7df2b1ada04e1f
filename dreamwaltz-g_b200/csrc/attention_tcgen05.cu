// Fused multi-head attention forward on tcgen05 (R15: every self/cross attention of the UNet and
// ControlNet transformer blocks).  O = softmax(Q K^T / sqrt(d)) V without ever writing the
// [T, Tk] score matrix to HBM (the unfused path moved ~2 GB per 64x64-resolution layer).
//
// One CTA = 128 query rows of one (batch, head); two CTAs share an SM.  Per 128-key block j:
//   warp 4 (one elected lane): TMA loads of K / V^T (double-buffered, 128B-swizzled), then
//       S(j+1) = Q K^T   tcgen05.mma  M=128 N=128 K=64*chunks   -> TMEM columns [0,128)
//         issued as soon as the softmax warps have pulled S(j) into registers (s_free), so the
//         tensor pipe works on the next block while the SIMT pipes do this block's exponentials;
//       O += P(j) V(j)   tcgen05.mma  M=128 N=round16(d) K=128 -> TMEM columns [128, 128+N)
//   warps 0-3 (thread == query row, TMEM lane == row): ONE pass over S: tcgen05.ld of the whole
//       128-column row into registers, row max, p = exp2(s*scale - m_ref), P written as bf16 into a
//       swizzled shared-memory A-operand tile.  O stays in TMEM for the whole key loop; the running
//       maximum is LAZY (FA4-style): m_ref only moves when the block maximum exceeds it by more than
//       2^8, and only then is the O row rescaled in TMEM (tcgen05.ld / tcgen05.st).
// V is consumed as V^T ([d, Tk], key index contiguous = K-major B operand), which the projection
// GEMM produces for free by swapping its operands (see dwg/diffusion/model.py).
#include <cuda.h>

#include "common.cuh"

namespace dwg {
namespace attn {

constexpr int BQ = 128, BKV = 128;

struct Params {
    int T, Tk, heads, hd;
    float scale_log2;               // softmax scale * log2(e)
    act_t* out;             // [B, T, heads*hd]
    int64_t out_row_stride;         // = heads*hd
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// one MUFU.EX2 per element (ex2_fast() wraps the same instruction in denormal range handling: ~4 issue slots; results
// below 2^-126 flush to zero here, which is what softmax wants).  The softmax warps are issue / MUFU bound.
__device__ __forceinline__ float ex2_fast(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
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
// 32 consecutive fp32 columns of this thread's TMEM lane, WITHOUT the wait (batch several, then tc_wait_ld)
__device__ __forceinline__ void tc_ld32_nowait(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
                 "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr));
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
                 "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 :: "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
                    "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
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
    return (1u << 4) | kIdescFmtAB | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(BQ >> 4) << 24);
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
    uint64_t* s_full = bars + 5; uint64_t* p_full = bars + 6; uint64_t* o_full = bars + 7; uint64_t* s_free = bars + 8;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

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
            mbar_init(s_full, 1); mbar_init(p_full, 128); mbar_init(o_full, 1); mbar_init(s_free, 128);
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
            auto issue_S = [&](int s) {
#pragma unroll
                for (int c = 0; c < CH; c++) {
                    const uint64_t dq = make_desc(smem_u32(sQ + c * TILE));
                    const uint64_t dk = make_desc(smem_u32(sK + (s * CH + c) * TILE));
#pragma unroll
                    for (int k = 0; k < 4; k++) tc_mma(tmem_S, dq + (uint64_t)(k * 2), dk + (uint64_t)(k * 2), idS, (c | k) != 0 ? 1u : 0u);
                }
                tc_commit(s_full);
            };
            mbar_wait(q_full, 0);
            mbar_wait(&kv_full[0], 0);
            tc_fence_after();
            issue_S(0);
            for (int j = 0; j < nblk; j++) {
                const int s = j & 1;
                if (j + 1 < nblk) {
                    // stage s^1 was last read by PV(j-1)
                    mbar_wait(&kv_empty[s ^ 1], (uint32_t)(((j + 1) >> 1) & 1) ^ 1u);
                    load_kv(j + 1, s ^ 1);
                    mbar_wait(&kv_full[s ^ 1], (uint32_t)(((j + 1) >> 1) & 1));
                    mbar_wait(s_free, (uint32_t)(j & 1));           // S(j) is in the softmax warps' registers
                    tc_fence_after();
                    issue_S(s ^ 1);                                  // S(j+1) overlaps the exponentials of block j
                }
                mbar_wait(p_full, (uint32_t)(j & 1));               // P(j) is in shared memory, O has been rescaled if needed
                tc_fence_after();
#pragma unroll
                for (int kc = 0; kc < 2; kc++) {
                    const uint64_t dp = make_desc(smem_u32(sP + kc * TILE));
                    const uint64_t dv = make_desc(smem_u32(sV + (s * 2 + kc) * VT));
#pragma unroll
                    for (int k = 0; k < 4; k++) tc_mma(tmem_O, dp + (uint64_t)(k * 2), dv + (uint64_t)(k * 2), idO, (j | kc | k) != 0 ? 1u : 0u);
                }
                tc_commit(o_full);
                tc_commit(&kv_empty[s]);
            }
        }
        __syncwarp();
    } else {
        const int r = threadIdx.x;                            // query row of the tile == TMEM lane
        const uint32_t lane_base = ((uint32_t)(warp * 32)) << 16;
        float m_ref = -INFINITY, l = 0.f;
        for (int j = 0; j < nblk; j++) {
            mbar_wait(s_full, (uint32_t)(j & 1));
            tc_fence_after();
            float sv[BKV];
#pragma unroll
            for (int c0 = 0; c0 < BKV; c0 += 32) tc_ld32_nowait(tmem_S + lane_base + (uint32_t)c0, sv + c0);
            tc_wait_ld();
            tc_fence_before();
            mbar_arrive(s_free);                                // the tensor pipe may overwrite S with block j+1
            if (j == nblk - 1) {
                const int kv_valid = p.Tk - j * BKV;
#pragma unroll
                for (int i = 0; i < BKV; i++) if (i >= kv_valid) sv[i] = -INFINITY;
            }
            float mx0 = sv[0], mx1 = sv[1], mx2 = sv[2], mx3 = sv[3];
#pragma unroll
            for (int i = 4; i < BKV; i += 4) {
                mx0 = fmaxf(mx0, sv[i]); mx1 = fmaxf(mx1, sv[i + 1]); mx2 = fmaxf(mx2, sv[i + 2]); mx3 = fmaxf(mx3, sv[i + 3]);
            }
            const float m_blk = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * p.scale_log2;
            // lazy running maximum: move the reference only when the block exceeds it by > 2^8
            const bool move = m_blk > m_ref + 8.0f;
            const float m_new = move ? m_blk : m_ref;
            const float alpha = move ? ex2_fast(m_ref - m_new) : 1.0f;      // 0 on the first block (m_ref = -inf)
            if (j > 0) {
                mbar_wait(o_full, (uint32_t)((j - 1) & 1));     // PV(j-1) done: O is stable and sP may be rewritten
                tc_fence_after();
                if (__any_sync(0xffffffffu, move)) {
#pragma unroll
                    for (int c0 = 0; c0 < NPV; c0 += 16) {
                        uint32_t v[16];
                        tc_ld16(tmem_O + lane_base + (uint32_t)c0, v);
#pragma unroll
                        for (int i = 0; i < 16; i++) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
                        tc_st16(tmem_O + lane_base + (uint32_t)c0, v);
                    }
                    tc_wait_st();
                }
            }
            l *= alpha;
            m_ref = m_new;
            float ls0 = 0.f, ls1 = 0.f;
#pragma unroll
            for (int c0 = 0; c0 < BKV; c0 += 16) {
                uint32_t pk[8];
#pragma unroll
                for (int i = 0; i < 16; i += 2) {
                    const float p0 = ex2_fast(fmaf(sv[c0 + i], p.scale_log2, -m_new));
                    const float p1 = ex2_fast(fmaf(sv[c0 + i + 1], p.scale_log2, -m_new));
                    ls0 += p0; ls1 += p1;
                    const act2_t h = f2_to_act2(p0, p1);
                    pk[i >> 1] = *reinterpret_cast<const uint32_t*>(&h);
                }
                // 16 columns = two 16-byte chunks of the 128-byte row; chunk index XOR (row & 7)
                const int kc = c0 >> 6;                         // which 64-key tile
                const int ch = (c0 & 63) >> 3;                  // chunk (8 bf16) inside the row
                uint8_t* rowp = sP + kc * TILE + r * 128;
                *reinterpret_cast<uint4*>(rowp + (((ch) ^ (r & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                *reinterpret_cast<uint4*>(rowp + (((ch + 1) ^ (r & 7)) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
            }
            l += ls0 + ls1;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> async proxy (MMA)
            tc_fence_before();
            mbar_arrive(p_full);
        }
        mbar_wait(o_full, (uint32_t)((nblk - 1) & 1));
        tc_fence_after();
        const int t = q0 + r;
        const float inv = 1.0f / l;
        act_t* dst = p.out + ((int64_t)b * p.T + t) * p.out_row_stride + (int64_t)head * p.hd;
#pragma unroll
        for (int c0 = 0; c0 < NPV; c0 += 16) {
            uint32_t v[16];
            tc_ld16(tmem_O + lane_base + (uint32_t)c0, v);
#pragma unroll
            for (int h8 = 0; h8 < 16; h8 += 8) {
                if (t < p.T && c0 + h8 < p.hd) {
                    uint4 pk;
                    act2_t h0 = f2_to_act2(__uint_as_float(v[h8]) * inv, __uint_as_float(v[h8 + 1]) * inv);
                    act2_t h1 = f2_to_act2(__uint_as_float(v[h8 + 2]) * inv, __uint_as_float(v[h8 + 3]) * inv);
                    act2_t h2 = f2_to_act2(__uint_as_float(v[h8 + 4]) * inv, __uint_as_float(v[h8 + 5]) * inv);
                    act2_t h3 = f2_to_act2(__uint_as_float(v[h8 + 6]) * inv, __uint_as_float(v[h8 + 7]) * inv);
                    pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
                    pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
                    *reinterpret_cast<uint4*>(dst + c0 + h8) = pk;
                }
            }
        }
        tc_fence_before();
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
    CUresult r = enc(m, DWG_TMAP_ACT, 4, const_cast<void*>(base), dd, ss, bb, ee, CU_TENSOR_MAP_INTERLEAVE_NONE,
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
    p.out = reinterpret_cast<act_t*>(out); p.out_row_stride = C;
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
