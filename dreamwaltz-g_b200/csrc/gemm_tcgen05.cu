// tcgen05 tensor-core GEMM / implicit-GEMM convolution for the diffusion blocks (R14/R15).
//
//   D[M,N] = epilogue( sum_k A[M,k] * B[N,k] )        bf16 operands, fp32 accumulation in TMEM
//
// One kernel serves every dense layer of the UNet / ControlNet / VAE encoder:
//   * plain (strided-batched) GEMM : linear layers, attention QK^T and PV, 1x1 convolutions;
//   * implicit-GEMM convolution    : 3x3 (or kxk) NHWC convolution WITHOUT an im2col buffer --
//     for every filter tap the A tile is fetched by a 4-D TMA box at the shifted pixel
//     coordinates; TMA zero-fills out-of-bounds pixels (= padding) and channels (= K tail), and
//     its element strides implement stride-2 convolutions.
// Structure (Blackwell-native, see blackwell_cuda_programming.md "Anatomy"):
//   warp 0   : TMA producer   (cp.async.bulk.tensor -> 128B-swizzled smem ring, mbarrier tx)
//   warp 1   : MMA issuer     (one elected lane: tcgen05.mma.cta_group::1.kind::f16, M=128,
//                              N=BN, K=16 per instruction; tcgen05.commit frees smem stages)
//   warp 2   : TMEM allocator (BN fp32 accumulator columns)
//   warps 4-7: epilogue       (tcgen05.ld 32x32b -> registers -> bias / time-embedding /
//                              residual / activation / scale -> bf16 or fp32 global stores)
// fp32 accumulators never touch registers during the main loop.
#include <cuda.h>
#include <cuda_bf16.h>

#include "common.cuh"

namespace dwg {
namespace gemm {

constexpr int BM = 128;
constexpr int BK = 64;              // 64 bf16 = 128 bytes = one swizzle-128B row
constexpr int UMMA_K = 16;
constexpr int kThreads = 256;

enum Act { ACT_NONE = 0, ACT_SILU = 1, ACT_GELU = 2, ACT_GEGLU = 3 };   // GEGLU: columns are (value, gate) pairs -> N/2 outputs

struct Params {
    // problem
    int M, N, K;                    // per batch entry (conv: M = N_img*Ho*Wo, K = Cin per tap)
    int taps_w, taps_h;             // 1,1 for plain GEMM
    int k_chunks;                   // ceil(K / 64)
    uint32_t a_bytes;               // bytes one A-tile TMA delivers (the box may hold < 128 rows for tiny images)
    // batching (plain GEMM): z = blockIdx.z -> (z % nb1, z / nb1)
    int nb1;
    // conv geometry (taps > 1 or conv_mode)
    int conv_mode;
    int Ho, Wo, BH, BW, BNI;        // output size and the pixel box of one 128-row tile
    int tiles_w, tiles_h;           // tiles along w / h per image group
    int stride, pad_h, pad_w;
    // output
    void* C;
    int out_bf16;
    int64_t ldc, c_b1, c_b2;        // element strides of C (row, batch dims)
    // epilogue
    const float* bias;              // [N] or null
    const float* bias2;             // [rows_b2][N] per-image bias (time embedding) or null
    int bias2_rows_per;             // rows (pixels) per bias2 row
    const __nv_bfloat16* residual;  // same layout as C (bf16) or null
    int64_t ldr, r_b1, r_b2;
    float alpha;                    // scale applied to the accumulator before bias
    int act;
    // persistent tile scheduler
    int m_tiles, n_tiles, nz, ksplit, total_tiles;
    float* workspace;               // [ksplit][rows_total][N] fp32 partial sums when ksplit > 1
    int64_t ws_slice;               // rows_total * N
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
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
                 "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                   "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                   "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, 128B-swizzled operand tile [rows][64 bf16]: 8-row groups are 1024 B apart (SBO),
// LBO is unused for swizzled K-major layouts (canonical value 1), descriptor version 1 (sm_100).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);                 // start address  [0,14)
    d |= (uint64_t)1 << 16;                                  // LBO            [16,30)
    d |= (uint64_t)(1024 >> 4) << 32;                        // SBO            [32,46)
    d |= (uint64_t)1 << 46;                                  // version        [46,48)
    d |= (uint64_t)2 << 61;                                  // SWIZZLE_128B   [61,64)
    return d;
}

__device__ __forceinline__ float apply_act(float x, int act) {
    if (act == ACT_SILU) return x / (1.0f + __expf(-x));
    if (act == ACT_GELU) return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f));
    return x;
}

struct TileCoord { int m_tile, n_tile, z, ks; };
__device__ __forceinline__ TileCoord decode_tile(const Params& p, int t) {
    TileCoord c;
    c.m_tile = t % p.m_tiles; t /= p.m_tiles;          // m fastest: CTAs running together share the B (weight) tile in L2
    c.n_tile = t % p.n_tiles; t /= p.n_tiles;
    c.ks = t % p.ksplit; t /= p.ksplit;
    c.z = t;
    return c;
}

// Persistent, warp-specialised: the CTA walks tiles blockIdx.x, +gridDim.x, ...; the smem ring and
// the two TMEM accumulator stages let the TMA / MMA of tile i+1 overlap the epilogue of tile i.
template <int BN, int STAGES>
__global__ void __launch_bounds__(kThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    constexpr uint32_t A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2;
    constexpr int TMEM_COLS = 2 * BN <= 32 ? 32 : (2 * BN <= 64 ? 64 : (2 * BN <= 128 ? 128 : (2 * BN <= 256 ? 256 : 512)));
    uint8_t* sA = smem;
    uint8_t* sB = smem + STAGES * A_BYTES;
    uint64_t* full = reinterpret_cast<uint64_t*>(sB + STAGES * B_BYTES);
    uint64_t* empty = full + STAGES;
    uint64_t* tmem_full = empty + STAGES;          // [2]
    uint64_t* tmem_empty = tmem_full + 2;          // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    float* s_bias = reinterpret_cast<float*>(tmem_slot + 4);          // [BN]

    const int warp = threadIdx.x >> 5;
    const int iters_total = p.taps_h * p.taps_w * p.k_chunks;

    if (warp == 0 && elect_one()) {
        asm volatile("prefetch.tensormap [%0];" :: "l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" :: "l"(&tmB) : "memory");
    }
    if (warp == 1 && elect_one()) {
        for (int s = 0; s < STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int a = 0; a < 2; a++) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "n"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();                 // everything above overlapped the predecessor's tail; its outputs are visible from here on
    pdl_trigger();

    // k-iteration range of a split
    auto it_range = [&](int ks, int& it0, int& it1) {
        it0 = (int)(((int64_t)iters_total * ks) / p.ksplit);
        it1 = (int)(((int64_t)iters_total * (ks + 1)) / p.ksplit);
    };

    if (warp == 0) {
        if (elect_one()) {
            uint32_t ring = 0;
            for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
                const TileCoord tc = decode_tile(p, t);
                int a_c1, a_c2, a_c3, b1 = 0, b2 = 0;
                if (p.conv_mode) {
                    const int tw = tc.m_tile % p.tiles_w;
                    const int th = (tc.m_tile / p.tiles_w) % p.tiles_h;
                    const int tn = tc.m_tile / (p.tiles_w * p.tiles_h);
                    a_c1 = tw * p.BW * p.stride - p.pad_w;
                    a_c2 = th * p.BH * p.stride - p.pad_h;
                    a_c3 = tn * p.BNI;
                } else {
                    b1 = tc.z % p.nb1; b2 = tc.z / p.nb1;
                    a_c1 = tc.m_tile * BM; a_c2 = b1; a_c3 = b2;
                }
                int it0, it1;
                it_range(tc.ks, it0, it1);
                for (int it = it0; it < it1; it++, ring++) {
                    const int s = ring % STAGES;
                    const uint32_t ph = (ring / STAGES) & 1u;
                    mbar_wait(&empty[s], ph ^ 1u);
                    mbar_expect_tx(&full[s], p.a_bytes + B_BYTES);
                    const int kc = it % p.k_chunks;
                    const int tap = it / p.k_chunks;
                    if (p.conv_mode) {
                        const int kw = tap % p.taps_w, kh = tap / p.taps_w;
                        tma_load_4d(sA + s * A_BYTES, &tmA, &full[s], kc * BK, a_c1 + kw, a_c2 + kh, a_c3);
                        tma_load_4d(sB + s * B_BYTES, &tmB, &full[s], kc * BK, tap, tc.n_tile * BN, 0);
                    } else {
                        tma_load_4d(sA + s * A_BYTES, &tmA, &full[s], kc * BK, a_c1, a_c2, a_c3);
                        tma_load_4d(sB + s * B_BYTES, &tmB, &full[s], kc * BK, tc.n_tile * BN, b1, b2);
                    }
                }
            }
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
        uint32_t ring = 0, tl = 0;
        for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, tl++) {
            const TileCoord tc = decode_tile(p, t);
            int it0, it1;
            it_range(tc.ks, it0, it1);
            const uint32_t as = tl & 1u;
            mbar_wait(&tmem_empty[as], ((tl >> 1) & 1u) ^ 1u);       // epilogue has drained this accumulator stage
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + as * BN;
            for (int it = it0; it < it1; it++, ring++) {
                const int s = ring % STAGES;
                const uint32_t ph = (ring / STAGES) & 1u;
                mbar_wait(&full[s], ph);
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t da = make_desc(smem_u32(sA + s * A_BYTES));
                    const uint64_t db = make_desc(smem_u32(sB + s * B_BYTES));
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; k++)
                        tc_mma(tmem_d, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (it > it0 || k > 0) ? 1u : 0u);
                    tc_commit(&empty[s]);
                    if (it == it1 - 1) tc_commit(&tmem_full[as]);
                }
                __syncwarp();
            }
        }
    } else if (warp >= 4) {
        const int q = warp & 3;
        const int lane = threadIdx.x & 31;
        const int r = q * 32 + lane;
        const int et = threadIdx.x - 128;                   // 0..127 among the epilogue threads
        uint32_t tl = 0;
        // vector fast path: 16-byte aligned rows for the residual / output
        const bool res_vec = p.residual && ((reinterpret_cast<uintptr_t>(p.residual) & 15) == 0) && (p.ldr % 8 == 0) &&
                             (p.r_b1 % 8 == 0) && (p.r_b2 % 8 == 0);
        for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, tl++) {
            const TileCoord tc = decode_tile(p, t);
            const uint32_t as = tl & 1u;
            int64_t c_off, r_off = 0, b2row = 0, ws_off = 0;
            bool row_ok;
            if (p.conv_mode) {
                const int tw = tc.m_tile % p.tiles_w;
                const int th = (tc.m_tile / p.tiles_w) % p.tiles_h;
                const int tn = tc.m_tile / (p.tiles_w * p.tiles_h);
                const int w = tw * p.BW + (r % p.BW);
                const int h = th * p.BH + (r / p.BW) % p.BH;
                const int n = tn * p.BNI + r / (p.BW * p.BH);
                row_ok = (w < p.Wo) && (h < p.Ho) && ((int64_t)n * p.Ho * p.Wo < (int64_t)p.M) && (r < p.BW * p.BH * p.BNI);
                const int64_t pix = ((int64_t)n * p.Ho + h) * p.Wo + w;
                c_off = pix * p.ldc;
                r_off = pix * p.ldr;
                ws_off = pix * p.N;
                b2row = p.bias2_rows_per > 0 ? pix / p.bias2_rows_per : 0;
            } else {
                const int b1 = tc.z % p.nb1, b2 = tc.z / p.nb1;
                const int64_t m = (int64_t)tc.m_tile * BM + r;
                row_ok = m < p.M;
                c_off = (int64_t)b1 * p.c_b1 + (int64_t)b2 * p.c_b2 + m * p.ldc;
                r_off = (int64_t)b1 * p.r_b1 + (int64_t)b2 * p.r_b2 + m * p.ldr;
                ws_off = ((int64_t)tc.z * p.M + m) * p.N;
                b2row = p.bias2_rows_per > 0 ? m / p.bias2_rows_per : 0;
            }
            const int ntile0 = tc.n_tile * BN;
            // ---- stage the bias slice of this tile in shared memory; prefetch the whole residual
            // row into registers BEFORE waiting for the accumulator (overlaps the MMA main loop)
            asm volatile("bar.sync 1, 128;" ::: "memory");           // previous tile's readers are done
            for (int c = et; c < BN; c += 128) {
                const int n = ntile0 + c;
                s_bias[c] = (p.bias && n < p.N) ? p.bias[n] : 0.f;
            }
            uint4 rr[BN / 8];
            const bool use_res_vec = res_vec && row_ok && p.ksplit == 1;
            if (use_res_vec) {
#pragma unroll
                for (int i = 0; i < BN / 8; i++) {
                    const int n = ntile0 + i * 8;
                    rr[i] = (n + 8 <= p.N) ? *reinterpret_cast<const uint4*>(p.residual + r_off + n) : make_uint4(0, 0, 0, 0);
                }
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            mbar_wait(&tmem_full[as], (tl >> 1) & 1u);
            tc_fence_after();
            const uint32_t taddr_row = tmem_base + as * BN + ((uint32_t)(q * 32) << 16);
#pragma unroll
            for (int c0 = 0; c0 < BN; c0 += 32) {
                uint32_t v[32];
                tc_ld32(taddr_row + (uint32_t)c0, v);
                if (c0 + 32 >= BN) {                       // last TMEM read of this tile: release the accumulator stage
                    tc_fence_before();
                    mbar_arrive(&tmem_empty[as]);
                }
                const int ncol0 = ntile0 + c0;
                if (!row_ok || ncol0 >= p.N) continue;
                const int nvalid = min(32, p.N - ncol0);
                if (p.ksplit > 1) {
                    // partial sums go to their own slice (no atomics, no memset); splitk_finalize adds the slices
                    float* ws = p.workspace + (int64_t)tc.ks * p.ws_slice + ws_off + ncol0;
                    if (nvalid == 32 && ((reinterpret_cast<uintptr_t>(ws) & 15) == 0)) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4)
                            *reinterpret_cast<float4*>(ws + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
                    } else {
                        for (int j = 0; j < nvalid; j++) ws[j] = __uint_as_float(v[j]);
                    }
                    continue;
                }
                float f[32];
#pragma unroll
                for (int j = 0; j < 32; j++) f[j] = __uint_as_float(v[j]) * p.alpha + s_bias[c0 + j];
                if (p.bias2) {
                    const float* b2p = p.bias2 + b2row * p.N + ncol0;
                    if (nvalid == 32 && ((reinterpret_cast<uintptr_t>(b2p) & 15) == 0)) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 bb = __ldg(reinterpret_cast<const float4*>(b2p + j));
                            f[j] += bb.x; f[j + 1] += bb.y; f[j + 2] += bb.z; f[j + 3] += bb.w;
                        }
                    } else {
                        for (int j = 0; j < nvalid; j++) f[j] += b2p[j];
                    }
                }
                if (p.act == ACT_GEGLU) {
                    // fused GEGLU (diffusers GEGLU: hidden * gelu(gate)); weight rows were interleaved at load time
                    float gl[16];
#pragma unroll
                    for (int j = 0; j < 16; j++) gl[j] = f[2 * j] * apply_act(f[2 * j + 1], ACT_GELU);
                    __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.C) + c_off + (ncol0 >> 1);
                    if (nvalid == 32 && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
                        for (int j = 0; j < 16; j += 8) {
                            uint4 pk;
                            __nv_bfloat162 h0 = __floats2bfloat162_rn(gl[j], gl[j + 1]), h1 = __floats2bfloat162_rn(gl[j + 2], gl[j + 3]);
                            __nv_bfloat162 h2 = __floats2bfloat162_rn(gl[j + 4], gl[j + 5]), h3 = __floats2bfloat162_rn(gl[j + 6], gl[j + 7]);
                            pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
                            pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
                            *reinterpret_cast<uint4*>(dst + j) = pk;
                        }
                    } else {
                        for (int j = 0; j < nvalid / 2; j++) dst[j] = __float2bfloat16(gl[j]);
                    }
                    continue;
                }
                if (p.act != ACT_NONE) {
#pragma unroll
                    for (int j = 0; j < 32; j++) f[j] = apply_act(f[j], p.act);
                }
                if (p.residual) {
                    if (use_res_vec && nvalid == 32) {
#pragma unroll
                        for (int i = 0; i < 4; i++) {
                            const uint4 u = rr[c0 / 8 + i];
                            const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                            for (int k = 0; k < 4; k++) {
                                const float2 t2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w4[k]));
                                f[i * 8 + 2 * k] += t2.x; f[i * 8 + 2 * k + 1] += t2.y;
                            }
                        }
                    } else {
                        for (int j = 0; j < nvalid; j++) f[j] += __bfloat162float(p.residual[r_off + ncol0 + j]);
                    }
                }
                if (p.out_bf16) {
                    __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.C) + c_off + ncol0;
                    if (nvalid == 32 && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
                        for (int j = 0; j < 32; j += 8) {
                            uint4 pk;
                            __nv_bfloat162 h0 = __floats2bfloat162_rn(f[j], f[j + 1]), h1 = __floats2bfloat162_rn(f[j + 2], f[j + 3]);
                            __nv_bfloat162 h2 = __floats2bfloat162_rn(f[j + 4], f[j + 5]), h3 = __floats2bfloat162_rn(f[j + 6], f[j + 7]);
                            pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
                            pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
                            *reinterpret_cast<uint4*>(dst + j) = pk;
                        }
                    } else {
                        for (int j = 0; j < nvalid; j++) dst[j] = __float2bfloat16(f[j]);
                    }
                } else {
                    float* dst = reinterpret_cast<float*>(p.C) + c_off + ncol0;
                    if (nvalid == 32 && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
                    } else {
                        for (int j = 0; j < nvalid; j++) dst[j] = f[j];
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "n"(TMEM_COLS));
    }
}

// split-K finalize: out = act(alpha * ws + bias + bias2) + residual
__global__ void __launch_bounds__(256)
splitk_finalize_kernel(const Params p, int64_t rows_total) {
    const int64_t total = rows_total * p.N;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = i / p.N;
        const int n = (int)(i - row * p.N);
        int64_t c_off, r_off, b2row;
        if (p.conv_mode) {
            c_off = row * p.ldc; r_off = row * p.ldr;
            b2row = p.bias2_rows_per > 0 ? row / p.bias2_rows_per : 0;
        } else {
            const int64_t z = row / p.M, m = row - z * p.M;
            const int b1 = (int)(z % p.nb1), b2 = (int)(z / p.nb1);
            c_off = (int64_t)b1 * p.c_b1 + (int64_t)b2 * p.c_b2 + m * p.ldc;
            r_off = (int64_t)b1 * p.r_b1 + (int64_t)b2 * p.r_b2 + m * p.ldr;
            b2row = p.bias2_rows_per > 0 ? m / p.bias2_rows_per : 0;
        }
        float acc = 0.f;
        for (int k = 0; k < p.ksplit; k++) acc += p.workspace[(int64_t)k * p.ws_slice + i];
        float x = acc * p.alpha;
        if (p.bias) x += p.bias[n];
        if (p.bias2) x += p.bias2[b2row * p.N + n];
        x = apply_act(x, p.act);
        if (p.residual) x += __bfloat162float(p.residual[r_off + n]);
        if (p.out_bf16) reinterpret_cast<__nv_bfloat16*>(p.C)[c_off + n] = __float2bfloat16(x);
        else reinterpret_cast<float*>(p.C)[c_off + n] = x;
    }
}

// ---------------------------------------------------------------------------- host side
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

// 4-D bf16 tensor map: dims (innermost first), byte strides for dims 1..3, box, element strides
static int make_map(CUtensorMap* m, const void* base, const uint64_t dims[4], const uint64_t strides_bytes[3],
                    const uint32_t box[4], const uint32_t estr[4]) {
    EncodeTiledFn enc = get_encode();
    if (!enc) { set_error("cuTensorMapEncodeTiled unavailable"); return DWG_ERR_CUDA; }
    cuuint64_t d[4] = {dims[0], dims[1], dims[2], dims[3]};
    cuuint64_t s[3] = {strides_bytes[0], strides_bytes[1], strides_bytes[2]};
    cuuint32_t b[4] = {box[0], box[1], box[2], box[3]};
    cuuint32_t e[4] = {estr[0], estr[1], estr[2], estr[3]};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), d, s, b, e, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d): dims=(%llu,%llu,%llu,%llu) strides=(%llu,%llu,%llu) box=(%u,%u,%u,%u)", (int)r,
                  (unsigned long long)d[0], (unsigned long long)d[1], (unsigned long long)d[2], (unsigned long long)d[3],
                  (unsigned long long)s[0], (unsigned long long)s[1], (unsigned long long)s[2], b[0], b[1], b[2], b[3]);
        return DWG_ERR_INVALID;
    }
    return DWG_OK;
}

static int g_num_sms = 0;
static float* g_ws = nullptr;            // split-K workspace (grown on demand, never shrunk)
static size_t g_ws_bytes = 0;

template <int BN, int STAGES>
static int launch_t(const CUtensorMap& tmA, const CUtensorMap& tmB, Params& p, int64_t rows_total, cudaStream_t st) {
    const size_t smem = 1024 + (size_t)STAGES * (BM * BK * 2 + BN * BK * 2) + (2 * STAGES + 4) * 8 + 16 + BN * 4 + 16;
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(gemm_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr_set = true;
    }
    if (!g_num_sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (g_num_sms <= 0) g_num_sms = kNumSMs;
    }
    p.n_tiles = (p.N + BN - 1) / BN;
    const int iters = p.taps_h * p.taps_w * p.k_chunks;
    const int64_t base_tiles = (int64_t)p.m_tiles * p.n_tiles * p.nz;
    // split-K when the tile count cannot fill the machine and the K loop is long
    int ks = 1;
    if (base_tiles * 2 <= g_num_sms && iters >= 8 && p.act != ACT_GEGLU) {
        ks = (int)(g_num_sms / base_tiles);
        if (ks > iters / 2) ks = iters / 2;
        if (ks > 16) ks = 16;
        if (ks < 1) ks = 1;
    }
    p.ksplit = ks;
    p.total_tiles = (int)(base_tiles * ks);
    if (ks > 1) {
        const size_t need = sizeof(float) * (size_t)rows_total * p.N * ks;
        p.ws_slice = rows_total * (int64_t)p.N;
        if (need > g_ws_bytes) {
            // NOTE: grows outside of stream capture only (warm-up pass sizes it); see DESIGN.md
            if (g_ws) cudaFree(g_ws);
            if (cudaMalloc(&g_ws, need) != cudaSuccess) { g_ws = nullptr; g_ws_bytes = 0; set_error("split-K workspace allocation failed"); return DWG_ERR_CUDA; }
            g_ws_bytes = need;
        }
        p.workspace = g_ws;
    }
    const int grid = p.total_tiles < g_num_sms ? p.total_tiles : g_num_sms;
    launch_pdl(gemm_kernel<BN, STAGES>, dim3(grid), dim3(kThreads), smem, st, tmA, tmB, p);
    if (ks > 1) {
        const int64_t total = rows_total * p.N;
        int64_t g = (total + 255) / 256;
        if (g > 4 * g_num_sms) g = 4 * g_num_sms;
        splitk_finalize_kernel<<<(unsigned)g, 256, 0, st>>>(p, rows_total);
    }
    return check_launch("tcgen05 gemm");
}

static int pick_bn(int N) {
    if (N <= 64) return 64;
    if (N % 256 == 0 || N >= 1024) return 256;
    return 128;
}

}  // namespace gemm
}  // namespace dwg

using namespace dwg;
using namespace dwg::gemm;

// D[b2][b1][M,N] = act(alpha * A[b2][b1][M,K] @ B[b2][b1][N,K]^T + bias[N] + bias2) + residual
// All strides in ELEMENTS.  A/B bf16, K contiguous.  out_bf16: 1 -> bf16 C, 0 -> fp32 C.
extern "C" int dwg_gemm_bf16(const void* A, int64_t lda, int64_t a_b1, int64_t a_b2,
                             const void* B, int64_t ldb, int64_t b_b1, int64_t b_b2,
                             void* C, int64_t ldc, int64_t c_b1, int64_t c_b2, int out_bf16,
                             int M, int N, int K, int nb1, int nb2,
                             const float* bias, const float* bias2, int bias2_rows_per,
                             const void* residual, int64_t ldr, int64_t r_b1, int64_t r_b2,
                             float alpha, int act, void* stream) {
    DWG_REQUIRE(A && B && C, "null pointer");
    DWG_REQUIRE(M > 0 && N > 0 && K > 0 && nb1 > 0 && nb2 > 0, "bad sizes");
    DWG_REQUIRE((lda % 8) == 0 && (ldb % 8) == 0 && (a_b1 % 8) == 0 && (a_b2 % 8) == 0 && (b_b1 % 8) == 0 && (b_b2 % 8) == 0,
                "A/B strides must be multiples of 8 elements (16 bytes) for TMA");
    DWG_REQUIRE(((uintptr_t)A & 15) == 0 && ((uintptr_t)B & 15) == 0, "A/B must be 16-byte aligned");
    DWG_REQUIRE(act != ACT_GEGLU || (N % 2 == 0 && out_bf16 && !residual), "GEGLU epilogue: even N, bf16 output, no residual");
    CUtensorMap tmA, tmB;
    const uint32_t ones[4] = {1, 1, 1, 1};
    const int bn = pick_bn(N);
    {
        const uint64_t dims[4] = {(uint64_t)K, (uint64_t)M, (uint64_t)nb1, (uint64_t)nb2};
        const uint64_t str[3] = {(uint64_t)lda * 2, (uint64_t)(nb1 > 1 ? a_b1 : lda * (int64_t)M) * 2, (uint64_t)(nb2 > 1 ? a_b2 : lda * (int64_t)M) * 2};
        const uint32_t box[4] = {BK, BM, 1, 1};
        int rc = make_map(&tmA, A, dims, str, box, ones);
        if (rc) return rc;
    }
    {
        const uint64_t dims[4] = {(uint64_t)K, (uint64_t)N, (uint64_t)nb1, (uint64_t)nb2};
        const uint64_t str[3] = {(uint64_t)ldb * 2, (uint64_t)(nb1 > 1 ? b_b1 : ldb * (int64_t)N) * 2, (uint64_t)(nb2 > 1 ? b_b2 : ldb * (int64_t)N) * 2};
        const uint32_t box[4] = {BK, (uint32_t)bn, 1, 1};
        int rc = make_map(&tmB, B, dims, str, box, ones);
        if (rc) return rc;
    }
    Params p = {};
    p.M = M; p.N = N; p.K = K; p.taps_w = 1; p.taps_h = 1; p.k_chunks = (K + BK - 1) / BK; p.nb1 = nb1; p.conv_mode = 0;
    p.a_bytes = BM * BK * 2;
    p.C = C; p.out_bf16 = out_bf16; p.ldc = ldc; p.c_b1 = c_b1; p.c_b2 = c_b2;
    p.bias = bias; p.bias2 = bias2; p.bias2_rows_per = bias2_rows_per;
    p.residual = reinterpret_cast<const __nv_bfloat16*>(residual); p.ldr = ldr; p.r_b1 = r_b1; p.r_b2 = r_b2;
    p.alpha = alpha; p.act = act;
    p.m_tiles = (M + BM - 1) / BM; p.nz = nb1 * nb2;
    const int64_t rows_total = (int64_t)M * nb1 * nb2;
    cudaStream_t st = (cudaStream_t)stream;
    if (bn == 64) return launch_t<64, 6>(tmA, tmB, p, rows_total, st);
    if (bn == 256) return launch_t<256, 4>(tmA, tmB, p, rows_total, st);
    return launch_t<128, 5>(tmA, tmB, p, rows_total, st);
}

// NHWC convolution as implicit GEMM.  x [Nimg,H,W,Cin] bf16 (Cin % 8 == 0), w [Cout,kh,kw,Cin] bf16,
// y [Nimg,Ho,Wo,Cout] (bf16 or fp32).  Padding is zero-fill (pad_h/pad_w applied on the top/left;
// the bottom/right extent follows from Ho/Wo, which covers SD's asymmetric (0,1,0,1) padding).
extern "C" int dwg_conv2d_nhwc_bf16(const void* x, const void* w, void* y, int out_bf16,
                                    int Nimg, int H, int W, int Cin, int Cout, int ksize, int stride,
                                    int pad_h, int pad_w, int Ho, int Wo,
                                    const float* bias, const float* bias2_per_image,
                                    const void* residual, int act, void* stream) {
    DWG_REQUIRE(x && w && y, "null pointer");
    DWG_REQUIRE(Cin % 8 == 0, "Cin must be a multiple of 8 (pad the channels)");
    DWG_REQUIRE(stride == 1 || stride == 2, "stride must be 1 or 2");
    DWG_REQUIRE(ksize == 1 || ksize == 3, "ksize must be 1 or 3");
    // pixel box of one 128-row tile
    int BW = 1;
    while (BW * 2 <= Wo && BW * 2 <= BM) BW *= 2;
    DWG_REQUIRE(Wo % BW == 0 || Wo < BW * 2, "unsupported output width");
    int BH = 1;
    while (BH * 2 <= Ho && BW * BH * 2 <= BM) BH *= 2;
    int BNI = BM / (BW * BH);
    if (BNI > Nimg) BNI = 1 << (31 - __builtin_clz(Nimg));         // largest power of two <= Nimg
    const int tiles_w = (Wo + BW - 1) / BW, tiles_h = (Ho + BH - 1) / BH, tiles_n = (Nimg + BNI - 1) / BNI;
    CUtensorMap tmA, tmB;
    const int bn = pick_bn(Cout);
    {
        const uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)Nimg};
        const uint64_t str[3] = {(uint64_t)Cin * 2, (uint64_t)W * Cin * 2, (uint64_t)H * W * Cin * 2};
        const uint32_t box[4] = {BK, (uint32_t)((BW - 1) * stride + 1), (uint32_t)((BH - 1) * stride + 1), (uint32_t)BNI};
        const uint32_t es[4] = {1, (uint32_t)stride, (uint32_t)stride, 1};
        int rc = make_map(&tmA, x, dims, str, box, es);
        if (rc) return rc;
    }
    {
        const int taps = ksize * ksize;
        const uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)taps, (uint64_t)Cout, 1};
        const uint64_t str[3] = {(uint64_t)Cin * 2, (uint64_t)taps * Cin * 2, (uint64_t)Cout * taps * Cin * 2};
        const uint32_t box[4] = {BK, 1, (uint32_t)bn, 1};
        const uint32_t ones[4] = {1, 1, 1, 1};
        int rc = make_map(&tmB, w, dims, str, box, ones);
        if (rc) return rc;
    }
    Params p = {};
    p.M = Nimg * Ho * Wo; p.N = Cout; p.K = Cin; p.taps_w = ksize; p.taps_h = ksize; p.k_chunks = (Cin + BK - 1) / BK;
    p.a_bytes = (uint32_t)(BW * BH * BNI * BK * 2);
    p.nb1 = 1; p.conv_mode = 1; p.Ho = Ho; p.Wo = Wo; p.BH = BH; p.BW = BW; p.BNI = BNI; p.tiles_w = tiles_w; p.tiles_h = tiles_h;
    p.stride = stride; p.pad_h = pad_h; p.pad_w = pad_w;
    p.C = y; p.out_bf16 = out_bf16; p.ldc = Cout;
    p.bias = bias; p.bias2 = bias2_per_image; p.bias2_rows_per = bias2_per_image ? Ho * Wo : 0;
    p.residual = reinterpret_cast<const __nv_bfloat16*>(residual); p.ldr = Cout;
    p.alpha = 1.0f; p.act = act;
    // rows of a tile with n >= Nimg are masked in the epilogue through M
    p.m_tiles = tiles_w * tiles_h * tiles_n; p.nz = 1;
    const int64_t rows_total = (int64_t)Nimg * Ho * Wo;
    cudaStream_t st = (cudaStream_t)stream;
    if (bn == 64) return launch_t<64, 6>(tmA, tmB, p, rows_total, st);
    if (bn == 256) return launch_t<256, 4>(tmA, tmB, p, rows_total, st);
    return launch_t<128, 5>(tmA, tmB, p, rows_total, st);
}
