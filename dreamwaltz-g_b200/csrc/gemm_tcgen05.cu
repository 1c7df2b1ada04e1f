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
// Structure (v3):
//   warp 0    : TMA producer   (cp.async.bulk.tensor -> 128B-swizzled smem ring, mbarrier tx)
//   warp 1    : MMA issuer     (one elected lane: tcgen05.mma.cta_group::1.kind::f16, M=128,
//                               N = BN (run-time, any multiple of 32 up to 256), K=16)
//   warp 2    : TMEM allocator (2 accumulator stages x 256 fp32 columns)
//   warps 4-11: epilogue       two warps per TMEM lane quadrant, alternating 32-column chunks:
//                               tcgen05.ld -> bias / time-embedding / activation / GEGLU ->
//                               (+ residual chunk fetched by TMA into smem) -> swizzled smem
//                               staging -> TMA store (cp.async.bulk.tensor ... global.shared).
//                               TMA clips the M / N tails, so the fast path has no masks.
//   split-K   : the CTAs of one output tile store fp32 partial tiles (coalesced, one private slice per
//               (tile, split)) and bump a per-tile counter; the LAST arriver adds the slices in split
//               order (fixed summation order = run-to-run deterministic) and runs the normal
//               epilogue -- no finalize kernel, no atomics on the data.
//   CTA pairs : (PAIR = true) two CTAs of one cluster (same TPC) run tcgen05.mma.cta_group::2 on a 256 x BN
//               tile: each CTA stages its own 128 A rows and HALF of the B tile (BN/2 rows); the leader's
//               single MMA lane drives both tensor cores, reading the B halves from both shared memories.
//               Per k-iteration a CTA then pulls 16 KB + BN*64 B instead of 16 KB + BN*128 B through the
//               ~6.3 KB/cycle chip-wide L2 port (B300_MICROARCH.md "LTS throughput cap"), which is what
//               bounds every large GEMM / convolution of the step.  Barriers: operand "full" lives in the
//               leader (both CTAs' TMA loads complete_tx on it), "empty" / "accumulator full" are reached
//               by multicast tcgen05.commit, "accumulator empty" collects local + remote epilogue arrivals.
//   halo mode : (3x3, stride 1, "same") the nine taps of one 64-channel slice read shifted windows of ONE
//               (16+2) x (8+2) pixel halo that is fetched once (18 line loads of 10 pixels, 16-row pitch) into
//               its own ring; the A descriptor of tap (kh,kw) simply starts kh lines + kw rows into the halo
//               (8-row groups = image lines, SBO = line pitch).  A traffic through the L2 port drops from
//               9x to 1.4x of the activation; the B (weight) taps stream through the ordinary ring.
// The epilogue body is deliberately compact (no unrolling over chunks, one instantiation per
// output kind): v2's 265 KB of SASS made short kernels instruction-fetch bound (ncu: stall_no_inst).
#include <cuda.h>

#include "common.cuh"

namespace dwg {
namespace gemm {

constexpr int BM = 128;
constexpr int BK = 64;              // 64 bf16 = 128 bytes = one swizzle-128B row
constexpr int UMMA_K = 16;
constexpr int kEpiWarps = 8;
constexpr int kThreads = 128 + 32 * kEpiWarps;
constexpr int kMaxStages = 8;
constexpr uint32_t A_BYTES = BM * BK * 2;
constexpr int TMEM_STAGE_COLS = 256;
// halo mode (3x3 stride-1 convolutions): tile = 8 (w) x 16 (h) output pixels of one image
constexpr int HALO_BW = 8, HALO_BH = 16;
constexpr int HALO_LINES = HALO_BH + 2;                       // 18 input lines
constexpr int HALO_LINE_PIX = HALO_BW + 2;                    // 10 pixels fetched per line
constexpr uint32_t HALO_LINE_BYTES = 16 * 128;                // 16-row pitch: every line starts on a swizzle-atom boundary
constexpr uint32_t HALO_BYTES = HALO_LINES * HALO_LINE_BYTES; // 36 KB per stage
constexpr uint32_t HALO_TX = HALO_LINES * HALO_LINE_PIX * 128;
constexpr int kMaxAStages = 4;

enum Act { ACT_NONE = 0, ACT_SILU = 1, ACT_GELU = 2, ACT_GEGLU = 3 };   // GEGLU: columns are (value, gate) pairs -> N/2 outputs
enum Epi { EPI_F16 = 0, EPI_F32 = 1, EPI_GEGLU = 2 };

struct Params {
    // problem
    int M, N, K;                    // per batch entry (conv: M = N_img*Ho*Wo, K = Cin per tap)
    int taps_w, taps_h;             // 1,1 for plain GEMM
    int k_chunks;                   // ceil(K / 64)
    uint32_t a_bytes;               // bytes one A-tile TMA delivers (the box may hold < 128 rows for tiny images)
    int nb1;                        // batching (plain GEMM): z -> (z % nb1, z / nb1)
    // conv geometry
    int conv_mode;
    int Ho, Wo, BH, BW, BNI, Nimg;  // output size and the pixel box of one 128-row tile
    int tiles_w, tiles_h;
    int stride, pad_h, pad_w;
    int ebw, ebh;                   // pixel box (w, h extent) of ONE epilogue warp's 32 rows
    // epilogue
    const float* bias;              // [N] or null
    const float* bias2;             // [rows_b2][N] per-image bias (time embedding) or null
    int bias2_rows_per;
    int has_res;                    // residual fetched through tmR (bf16, same geometry as C)
    // LayerNorm folded into this GEMM (plain GEMM only).  The operand holds the UN-normalised activations and gamma is
    // pre-multiplied into the weight (W' = W * gamma); with c1 = row sums of W' and c2 = W beta (+ bias, passed as the bias):
    //   ln_mode 1 (activations = A rows):    out[m][n] = rstd_m * (acc - mean_m * c1[n]) + c2[n]
    //   ln_mode 2 (activations = B rows):    out[m][n] = rstd_n * (acc - mean_n * c1[m]) + ln_rowbias[m]
    // ln_stats = fixed-point (sum, sum of squares) per activation row, written by the producing GEMM (rowstats_out).
    int ln_mode; float ln_inv_dim, ln_eps;
    const unsigned long long* ln_stats; const float* ln_c1; const float* ln_rowbias;
    unsigned long long* rowstats_out;   // [M][2]: per output row, sum / sum of squares of the stored fp16 values (2^-20 fixed point)
    int a_bcast1;                   // A has no batch-1 dimension (stride 0: one weight matrix for every batch entry)
    float alpha;
    int act;
    // slow path (unaligned output / residual strides): direct per-thread stores
    int direct;
    void* C; int64_t ldc, c_b1, c_b2;
    const act_t* residual; int64_t ldr, r_b1, r_b2;
    // tiling
    int BN, stages;
    int m_tiles, n_tiles, nz, ksplit, total_tiles;
    int halo, a_stages, halo_bo;    // halo mode, depth of the halo ring, descriptor base-offset convention (see make_desc_halo)
    int m_sched;                    // scheduler units along M: m_tiles (single CTA) or m_tiles / 2 (CTA pair)
    float* workspace;               // split-K fp32 partial-tile slices [tile][split]
    int* counters;                  // one per output tile, self-resetting
    unsigned long long* trace;      // optional [8] per-launch timeline of CTA 0 (globaltimer ns), null = off
    unsigned long long* colstats;   // optional [groups][N][2] fixed-point (2^-20) per-column sum / sum of squares of the fp16 OUTPUT (GroupNorm statistics)
    int cs_rows;                    // plain GEMM: rows per statistics group (image); conv: unused (group = image)
};

__device__ __forceinline__ unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define TRACE(i) do { if (p.trace && blockIdx.x == 0) p.trace[i] = gtimer(); } while (0)
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
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the same-offset mbarrier of CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t rank) {
    asm volatile("{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, %1;\n\t"
                 "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}" :: "r"(smem_u32(bar)), "r"(rank) : "memory");
}
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;      // shared::cluster address of the same offset in the pair's even (leader) CTA
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 :: "r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
// CTA-pair variant: executed by both CTAs, the transaction bytes are credited to the LEADER's barrier
__device__ __forceinline__ void tma_load_4d_pair(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 :: "r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, const void* src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 :: "l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
// CTA pair: the commit arrives on the same-offset barrier of both CTAs
__device__ __forceinline__ void tc_commit_pair(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 :: "r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tc_mma_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
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

// asynchronous variant: the registers are valid only after tc_wait_ld() + reg_fence() on the same array
__device__ __forceinline__ void tc_ld32_nowait(uint32_t taddr, float (&v)[32]) {
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
// empty asm that "modifies" the 32 registers: pins every later use behind the preceding tc_wait_ld()
__device__ __forceinline__ void reg_fence(float (&v)[32]) {
    asm volatile("" : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]), "+f"(v[4]), "+f"(v[5]), "+f"(v[6]), "+f"(v[7]),
                      "+f"(v[8]), "+f"(v[9]), "+f"(v[10]), "+f"(v[11]), "+f"(v[12]), "+f"(v[13]), "+f"(v[14]), "+f"(v[15]),
                      "+f"(v[16]), "+f"(v[17]), "+f"(v[18]), "+f"(v[19]), "+f"(v[20]), "+f"(v[21]), "+f"(v[22]), "+f"(v[23]),
                      "+f"(v[24]), "+f"(v[25]), "+f"(v[26]), "+f"(v[27]), "+f"(v[28]), "+f"(v[29]), "+f"(v[30]), "+f"(v[31]));
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

// Halo operand of tap (kh, kw): 16 groups of 8 rows, one group per image line (SBO = line pitch), starting kw rows
// into the line.  The start is then not aligned to the 1024-byte swizzle repeat.  Measured on sm_100a
// (tools/halo_probe.py): the 128B-swizzle XOR is applied to the ABSOLUTE shared-memory address (exactly as TMA
// wrote the lines), so the shifted start needs NO correction -- the "base offset" field [49,52) must stay 0
// (setting it to (start >> 7) & 7 gives wrong results).
__device__ __forceinline__ uint64_t make_desc_halo(uint32_t saddr, uint32_t base_off) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(HALO_LINE_BYTES >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(base_off & 7u) << 49;
    d |= (uint64_t)2 << 61;
    return d;
}

// x * sigmoid(x) with sigmoid(x) = 0.5 + 0.5 tanh(x/2): one MUFU op (tanh.approx.f32, rel err 2^-11 < bf16 rounding)
__device__ __forceinline__ float silu_f(float x) {
    const float h = 0.5f * x;
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
    return fmaf(h, t, h);
}
// exact-erf GELU (diffusers GEGLU / F.gelu default) with erf from Abramowitz-Stegun 7.1.26
// (|abs err| <= 1.5e-7, far below the bf16 output rounding): one rcp + one ex2 instead of erff's
// ~40-instruction branchy expansion.
__device__ __forceinline__ float gelu_f(float x) {
    const float z = fabsf(x) * 0.70710678118654752f;
    const float t = __fdividef(1.0f, fmaf(0.3275911f, z, 1.0f));
    float poly = fmaf(1.061405429f, t, -1.453152027f);
    poly = fmaf(poly, t, 1.421413741f);
    poly = fmaf(poly, t, -0.284496736f);
    poly = fmaf(poly, t, 0.254829592f);
    const float erf_abs = 1.0f - poly * t * __expf(-z * z);
    return 0.5f * x * (1.0f + copysignf(erf_abs, x));
}

struct TileCoord { int m_tile, n_tile, z, ks; };
__device__ __forceinline__ TileCoord decode_tile(const Params& p, int t, int m_mul, int m_add) {
    TileCoord c;
    c.m_tile = (t % p.m_sched) * m_mul + m_add; t /= p.m_sched;          // m fastest: CTAs running together share the B (weight) tile in L2
    c.n_tile = t % p.n_tiles; t /= p.n_tiles;
    c.ks = t % p.ksplit; t /= p.ksplit;
    c.z = t;
    return c;
}

__device__ __forceinline__ uint32_t pack_act(float a, float b) { return pack_act2(a, b); }

// Column statistics of one finished tile: thread `col` folds the (up to four) 32-row quadrants that belong to the same
// statistics group and adds the totals as 2^-20 fixed-point integers (order-free => deterministic) to
// colstats[slot][group][column][2]; slot = m_tile & 3 spreads the same-address atomics of concurrently finishing tiles.
__device__ __forceinline__ void cs_flush(const Params& p, const float* s_cs, const int* s_csmeta, uint32_t parity, int col) {
    const int* meta = s_csmeta + parity * 8;
    if (col < 0 || col >= p.BN || !meta[0]) return;
    const int n = meta[2] + col;
    if (n >= p.N) return;
    const int slot = meta[1];
    const int64_t groups = p.conv_mode ? p.Nimg : ((int64_t)p.nz * p.M + p.cs_rows - 1) / p.cs_rows;
    float S = 0.f, Q = 0.f;
    int cur = -1;
    for (int q = 0; q < 4; q++) {
        const int g = meta[4 + q];
        if (g != cur) {
            if (cur >= 0) {
                unsigned long long* dst = p.colstats + (((size_t)slot * groups + cur) * p.N + n) * 2;
                atomicAdd(dst, (unsigned long long)__float2ll_rn(S * 1048576.0f));
                atomicAdd(dst + 1, (unsigned long long)__float2ll_rn(Q * 1048576.0f));
            }
            cur = g; S = 0.f; Q = 0.f;
        }
        if (g >= 0) {
            const float2 v = *reinterpret_cast<const float2*>(s_cs + ((parity * 4 + q) * 256 + col) * 2);
            S += v.x; Q += v.y;
        }
    }
    if (cur >= 0) {
        unsigned long long* dst = p.colstats + (((size_t)slot * groups + cur) * p.N + n) * 2;
        atomicAdd(dst, (unsigned long long)__float2ll_rn(S * 1048576.0f));
        atomicAdd(dst + 1, (unsigned long long)__float2ll_rn(Q * 1048576.0f));
    }
}

// Persistent, warp-specialised: the CTA walks tiles blockIdx.x, +gridDim.x, ...; the smem ring and
// the two TMEM accumulator stages let the TMA / MMA of tile i+1 overlap the epilogue of tile i.
// LNF: LayerNorm folding / row statistics compiled in (a separate instantiation: the ~350 launches of a step that use neither
// run exactly the kernel they ran before)
template <int EPI, bool PAIR, bool LNF>
__global__ void __launch_bounds__(kThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
            const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR,
            const __grid_constant__ Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    constexpr int ACC_PER_CHUNK = (EPI == EPI_GEGLU) ? 64 : 32;        // accumulator columns per epilogue chunk (-> 32 outputs)
    constexpr uint32_t STG_BYTES = (EPI == EPI_F32) ? 4096 : 2048;     // 32 rows x 32 outputs
    const uint32_t B_BYTES = (uint32_t)(PAIR ? p.BN / 2 : p.BN) * (BK * 2);      // this CTA's share of the B tile
    const uint32_t rank = PAIR ? cluster_ctarank() : 0u;                      // 0 = leader (issues the MMAs of the pair)
    const int m_mul = PAIR ? 2 : 1, m_add = (int)rank;
    const int tile0 = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int tile_step = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    uint8_t* sA = smem;
    uint8_t* sB = sA + (p.halo ? p.a_stages * HALO_BYTES : p.stages * A_BYTES);
    uint8_t* sStage = sB + p.stages * B_BYTES;                         // [kEpiWarps][2][STG_BYTES]
    uint8_t* sRes = sStage + kEpiWarps * 2 * STG_BYTES;                // [kEpiWarps][2][2048] (only when has_res)
    float* s_bias = reinterpret_cast<float*>(sRes + (p.has_res ? kEpiWarps * 2 * 2048 : 0));      // [tile parity][image 0 / image 1 or LN c1 | mean / LN rstd][256]
    float* s_cs = s_bias + 2 * 3 * 256;                                // [tile parity][quadrant][256 columns][2]  (only when p.colstats)
    int* s_csmeta = reinterpret_cast<int*>(s_cs + (p.colstats ? 2 * 4 * 256 * 2 : 0));      // [tile parity][8]: valid, slot, ntile0, -, grp[4]
    uint64_t* full = reinterpret_cast<uint64_t*>(s_csmeta + (p.colstats ? 16 : 0));
    uint64_t* empty = full + kMaxStages;
    uint64_t* tmem_full = empty + kMaxStages;          // [2]
    uint64_t* tmem_empty = tmem_full + 2;              // [2]
    uint64_t* res_bar = tmem_empty + 2;                // [kEpiWarps][2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_bar + 2 * kEpiWarps);
    volatile int* s_flag = reinterpret_cast<volatile int*>(tmem_slot + 1);
    uint64_t* fullA = reinterpret_cast<uint64_t*>(tmem_slot + 2);      // [kMaxAStages] halo ring
    uint64_t* emptyA = fullA + kMaxAStages;

    const int warp = threadIdx.x >> 5;
    const int iters_total = p.taps_h * p.taps_w * p.k_chunks;
    if (threadIdx.x == 0) TRACE(0);

    if (warp == 0 && elect_one()) {
        asm volatile("prefetch.tensormap [%0];" :: "l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" :: "l"(&tmB) : "memory");
        asm volatile("prefetch.tensormap [%0];" :: "l"(&tmC) : "memory");
        if (p.has_res) asm volatile("prefetch.tensormap [%0];" :: "l"(&tmR) : "memory");
    }
    if (warp == 1 && elect_one()) {
        for (int s = 0; s < p.stages; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int a = 0; a < 2; a++) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], PAIR ? 2 * kEpiWarps : kEpiWarps); }
        for (int i = 0; i < 2 * kEpiWarps; i++) mbar_init(&res_bar[i], 1);
        for (int a = 0; a < kMaxAStages; a++) { mbar_init(&fullA[a], 1); mbar_init(&emptyA[a], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        if (PAIR) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "n"(2 * TMEM_STAGE_COLS));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "n"(2 * TMEM_STAGE_COLS));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
        }
    }
    tc_fence_before();
    if (PAIR) cluster_sync_all();            // the peer's barriers are initialised before anything arrives on them
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) TRACE(1);
    pdl_wait();
    if (threadIdx.x == 0) TRACE(2);                 // everything above overlapped the predecessor's tail; its outputs are visible from here on
    pdl_trigger();

    // k-iteration range of a split
    auto it_range = [&](int ks, int& it0, int& it1) {
        it0 = (int)(((int64_t)iters_total * ks) / p.ksplit);
        it1 = (int)(((int64_t)iters_total * (ks + 1)) / p.ksplit);
    };

    if (warp == 0) {
        if (elect_one()) {
            uint32_t s = 0, ph = 0;
            const int n_off = PAIR ? (int)rank * (p.BN / 2) : 0;          // this CTA's half of the B tile
            for (int t = tile0; t < p.total_tiles; t += tile_step) {
                const TileCoord tc = decode_tile(p, t, m_mul, m_add);
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
                    a_c1 = tc.m_tile * BM; a_c2 = p.a_bcast1 ? 0 : b1; a_c3 = b2;
                }
                int it0, it1;
                it_range(tc.ks, it0, it1);
                if (p.halo) {
                    // halo mode: this lane streams the weight taps only (channel-slice major, tap minor)
                    for (int kc = 0; kc < p.k_chunks; kc++) {
                        for (int tap = 0; tap < 9; tap++) {
                            mbar_wait(&empty[s], ph ^ 1u);
                            if (PAIR) {
                                if (rank == 0) mbar_expect_tx(&full[s], 2u * B_BYTES);
                                tma_load_4d_pair(sB + s * B_BYTES, &tmB, &full[s], kc * BK, tap, tc.n_tile * p.BN + n_off, 0);
                            } else {
                                mbar_expect_tx(&full[s], B_BYTES);
                                tma_load_4d(sB + s * B_BYTES, &tmB, &full[s], kc * BK, tap, tc.n_tile * p.BN, 0);
                            }
                            if (++s == (uint32_t)p.stages) { s = 0; ph ^= 1u; }
                        }
                    }
                    continue;
                }
                int kc = it0 % p.k_chunks, tap = it0 / p.k_chunks;
                for (int it = it0; it < it1; it++) {
                    mbar_wait(&empty[s], ph ^ 1u);
                    if (PAIR) {
                        // the leader's barrier counts the bytes of BOTH CTAs (the peer only issues its loads)
                        if (rank == 0) mbar_expect_tx(&full[s], 2u * (p.a_bytes + B_BYTES));
                        if (p.conv_mode) {
                            const int kw = tap % p.taps_w, kh = tap / p.taps_w;
                            tma_load_4d_pair(sA + s * A_BYTES, &tmA, &full[s], kc * BK, a_c1 + kw, a_c2 + kh, a_c3);
                            tma_load_4d_pair(sB + s * B_BYTES, &tmB, &full[s], kc * BK, tap, tc.n_tile * p.BN + n_off, 0);
                        } else {
                            tma_load_4d_pair(sA + s * A_BYTES, &tmA, &full[s], kc * BK, a_c1, a_c2, a_c3);
                            tma_load_4d_pair(sB + s * B_BYTES, &tmB, &full[s], kc * BK, tc.n_tile * p.BN + n_off, b1, b2);
                        }
                    } else {
                        mbar_expect_tx(&full[s], p.a_bytes + B_BYTES);
                        if (p.conv_mode) {
                            const int kw = tap % p.taps_w, kh = tap / p.taps_w;
                            tma_load_4d(sA + s * A_BYTES, &tmA, &full[s], kc * BK, a_c1 + kw, a_c2 + kh, a_c3);
                            tma_load_4d(sB + s * B_BYTES, &tmB, &full[s], kc * BK, tap, tc.n_tile * p.BN, 0);
                        } else {
                            tma_load_4d(sA + s * A_BYTES, &tmA, &full[s], kc * BK, a_c1, a_c2, a_c3);
                            tma_load_4d(sB + s * B_BYTES, &tmB, &full[s], kc * BK, tc.n_tile * p.BN, b1, b2);
                        }
                    }
                    if (++kc == p.k_chunks) { kc = 0; tap++; }
                    if (++s == (uint32_t)p.stages) { s = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == 3) {
        // halo ring producer: 18 line loads (10 pixels x 64 channels each) per 64-channel slice of a tile
        if (p.halo && elect_one()) {
            uint32_t sa = 0, pha = 0;
            for (int t = tile0; t < p.total_tiles; t += tile_step) {
                const TileCoord tc = decode_tile(p, t, m_mul, m_add);
                const int tw = tc.m_tile % p.tiles_w;
                const int th = (tc.m_tile / p.tiles_w) % p.tiles_h;
                const int tn = tc.m_tile / (p.tiles_w * p.tiles_h);
                const int w0 = tw * HALO_BW - 1, h0 = th * HALO_BH - 1;
                for (int kc = 0; kc < p.k_chunks; kc++) {
                    mbar_wait(&emptyA[sa], pha ^ 1u);
                    uint8_t* dst = sA + sa * HALO_BYTES;
                    if (PAIR) {
                        if (rank == 0) mbar_expect_tx(&fullA[sa], 2u * HALO_TX);       // the leader's barrier counts both CTAs' bytes
                        for (int ln = 0; ln < HALO_LINES; ln++)
                            tma_load_4d_pair(dst + ln * HALO_LINE_BYTES, &tmA, &fullA[sa], kc * BK, w0, h0 + ln, tn);
                    } else {
                        mbar_expect_tx(&fullA[sa], HALO_TX);
                        for (int ln = 0; ln < HALO_LINES; ln++)
                            tma_load_4d(dst + ln * HALO_LINE_BYTES, &tmA, &fullA[sa], kc * BK, w0, h0 + ln, tn);
                    }
                    if (++sa == (uint32_t)p.a_stages) { sa = 0; pha ^= 1u; }
                }
            }
        }
    } else if (warp == 1 && rank == 0) {
        const uint32_t idesc = (1u << 4) | kIdescFmtAB | ((uint32_t)(p.BN >> 3) << 17) | ((uint32_t)((PAIR ? 2 * BM : BM) >> 4) << 24);
        uint32_t s = 0, ph = 0, tl = 0, sa = 0, pha = 0;
        for (int t = tile0; t < p.total_tiles; t += tile_step, tl++) {
            const TileCoord tc = decode_tile(p, t, m_mul, m_add);
            int it0, it1;
            it_range(tc.ks, it0, it1);
            const uint32_t as = tl & 1u;
            mbar_wait(&tmem_empty[as], ((tl >> 1) & 1u) ^ 1u);       // epilogue has drained this accumulator stage
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + as * TMEM_STAGE_COLS;
            if (p.halo) {
                for (int kc = 0; kc < p.k_chunks; kc++) {
                    mbar_wait(&fullA[sa], pha);
                    for (int tap = 0; tap < 9; tap++) {
                        mbar_wait(&full[s], ph);
                        tc_fence_after();
                        if (elect_one()) {
                            if (tl == 0 && kc == 0 && tap == 0) TRACE(3);
                            const uint32_t kh = (uint32_t)tap / 3u, kw = (uint32_t)tap % 3u;
                            const uint64_t da = make_desc_halo(smem_u32(sA + sa * HALO_BYTES) + kh * HALO_LINE_BYTES + kw * 128u, p.halo_bo ? kw : 0u);
                            const uint64_t db = make_desc(smem_u32(sB + s * B_BYTES));
                            const int nk = min(BK / UMMA_K, (p.K - kc * BK + UMMA_K - 1) / UMMA_K);
#pragma unroll
                            for (int k = 0; k < BK / UMMA_K; k++) {
                                const uint32_t acc = (kc > 0 || tap > 0 || k > 0) ? 1u : 0u;
                                if (k < nk) {
                                    if (PAIR) tc_mma_pair(tmem_d, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, acc);
                                    else tc_mma(tmem_d, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, acc);
                                }
                            }
                            if (PAIR) tc_commit_pair(&empty[s]); else tc_commit(&empty[s]);
                            if (tap == 8) { if (PAIR) tc_commit_pair(&emptyA[sa]); else tc_commit(&emptyA[sa]); }
                            if (tap == 8 && kc == p.k_chunks - 1) {
                                if (PAIR) tc_commit_pair(&tmem_full[as]); else tc_commit(&tmem_full[as]);
                                if (tl == 0) TRACE(4);
                            }
                        }
                        __syncwarp();
                        if (++s == (uint32_t)p.stages) { s = 0; ph ^= 1u; }
                    }
                    if (++sa == (uint32_t)p.a_stages) { sa = 0; pha ^= 1u; }
                }
                continue;
            }
            for (int it = it0; it < it1; it++) {
                mbar_wait(&full[s], ph);
                tc_fence_after();
                if (elect_one()) {
                    if (tl == 0 && it == it0) TRACE(3);
                    const uint64_t da = make_desc(smem_u32(sA + s * A_BYTES));
                    const uint64_t db = make_desc(smem_u32(sB + s * B_BYTES));
                    // K = 16 steps of this 64-wide chunk that hold real data (the rest is TMA zero-fill: K tails, 8/16-channel convs)
                    const int nk = min(BK / UMMA_K, (p.K - (it % p.k_chunks) * BK + UMMA_K - 1) / UMMA_K);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; k++) {
                        if (k < nk) {
                            if (PAIR) tc_mma_pair(tmem_d, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (it > it0 || k > 0) ? 1u : 0u);
                            else tc_mma(tmem_d, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (it > it0 || k > 0) ? 1u : 0u);
                        }
                    }
                    if (PAIR) tc_commit_pair(&empty[s]); else tc_commit(&empty[s]);
                    if (it == it1 - 1) {
                        if (PAIR) tc_commit_pair(&tmem_full[as]); else tc_commit(&tmem_full[as]);
                        if (tl == 0) TRACE(4);
                    }
                }
                __syncwarp();
                if (++s == (uint32_t)p.stages) { s = 0; ph ^= 1u; }
            }
        }
    } else if (warp >= 4) {
        const int e = warp - 4;                             // 0..7
        const int q = warp & 3;                             // TMEM lane quadrant this warp may read
        const int half = e >> 2;                            // which chunks of the tile (even / odd)
        const int lane = threadIdx.x & 31;
        uint8_t* stg = sStage + e * 2 * STG_BYTES;
        uint8_t* rsm = sRes + e * 2 * 2048;
        uint64_t* rbar = res_bar + e * 2;
        const int nchunks = p.BN / ACC_PER_CHUNK;
        uint32_t tl = 0, slot = 0, rph = 0;                 // slot: running chunk counter (buffer = slot & 1); rph: phase bits of rbar[0..1]
        // swizzled 16-byte-chunk offsets of this thread's staging row
        const uint32_t row_off = (EPI == EPI_F32) ? (uint32_t)lane * 128u : (uint32_t)lane * 64u;
        const uint32_t sw = (EPI == EPI_F32) ? (uint32_t)(lane & 7) : (uint32_t)((lane >> 1) & 3);
        const uint32_t rsw = (uint32_t)((lane >> 1) & 3);   // residual rows are always bf16 (64 B, SWIZZLE_64B)
        for (int t = tile0; t < p.total_tiles; t += tile_step, tl++) {
            const TileCoord tc = decode_tile(p, t, m_mul, m_add);
            const uint32_t as = tl & 1u;
            const int r0 = q * 32;
            int c1, c2, c3;
            bool box_ok;
            int64_t b2row = 0, d_row = 0, d_res = 0;       // bias2 row; direct-path row offsets
            bool row_ok = true;
            if (p.conv_mode) {
                const int tw = tc.m_tile % p.tiles_w;
                const int th = (tc.m_tile / p.tiles_w) % p.tiles_h;
                const int tn = tc.m_tile / (p.tiles_w * p.tiles_h);
                c1 = tw * p.BW + (r0 % p.BW);
                c2 = th * p.BH + (r0 / p.BW) % p.BH;
                c3 = tn * p.BNI + r0 / (p.BW * p.BH);
                box_ok = (c1 < p.Wo) && (c2 < p.Ho) && (c3 < p.Nimg) && (r0 < p.BW * p.BH * p.BNI);
                const int w = c1 + lane % p.ebw, h = c2 + (lane / p.ebw) % p.ebh, n = c3 + lane / (p.ebw * p.ebh);
                row_ok = box_ok && (w < p.Wo) && (h < p.Ho) && (n < p.Nimg);
                const int64_t pix = ((int64_t)n * p.Ho + h) * p.Wo + w;
                b2row = n;
                d_row = pix * p.ldc; d_res = pix * p.ldr;
            } else {
                c1 = tc.m_tile * BM + r0; c2 = tc.z % p.nb1; c3 = tc.z / p.nb1;
                box_ok = c1 < p.M;
                const int64_t m = c1 + lane;
                row_ok = m < p.M;
                b2row = p.bias2_rows_per > 0 ? m / p.bias2_rows_per : 0;
                d_row = (int64_t)c2 * p.c_b1 + (int64_t)c3 * p.c_b2 + m * p.ldc;
                d_res = (int64_t)c2 * p.r_b1 + (int64_t)c3 * p.r_b2 + m * p.ldr;
            }
            const int ntile0 = tc.n_tile * p.BN;
            const int otile0 = (EPI == EPI_GEGLU) ? (ntile0 >> 1) : ntile0;      // first OUTPUT column of the tile
            const bool use_res = p.has_res && box_ok && !p.direct;
            // ---- bias (+ per-image time-embedding bias) of the tile's columns -> shared memory, while the MMAs run.
            // Double-buffered by tile parity: the one barrier per tile also orders the reuse two tiles later.
            const bool stage_b2 = p.bias2 && p.conv_mode && p.BNI <= 2;
            float* sb = s_bias + (tl & 1u) * 768;
            {
                const int img0 = p.conv_mode ? (tc.m_tile / (p.tiles_w * p.tiles_h)) * p.BNI : 0;
                for (int cc = threadIdx.x - 128; cc < p.BN; cc += 32 * kEpiWarps) {
                    const int n = ntile0 + cc;
                    float v0 = 0.f, v1 = 0.f;
                    if (n < p.N) {
                        if (p.bias) v0 = v1 = p.bias[n];
                        if (stage_b2) {
                            v0 += p.bias2[(int64_t)img0 * p.N + n];
                            if (img0 + 1 < p.Nimg) v1 += p.bias2[(int64_t)(img0 + 1) * p.N + n];
                        }
                    }
                    if (LNF && p.ln_mode == 1) v1 = n < p.N ? p.ln_c1[n] : 0.f;
                    if (LNF && p.ln_mode == 2) {
                        float mean = 0.f, rstd = 0.f;
                        if (n < p.N) {
                            const double inv = (double)p.ln_inv_dim * (1.0 / 1048576.0);
                            const int64_t sn = (int64_t)tc.z * p.N + n;
                            const double mu = (double)(long long)p.ln_stats[2 * sn] * inv, qq = (double)(long long)p.ln_stats[2 * sn + 1] * inv;
                            mean = (float)mu;
                            rstd = rsqrtf((float)fmax(qq - mu * mu, 0.0) + p.ln_eps);
                        }
                        v1 = mean;
                        sb[512 + cc] = rstd;
                    }
                    sb[cc] = v0; sb[256 + cc] = v1;
                }
                asm volatile("bar.sync 2, %0;" :: "n"(32 * kEpiWarps) : "memory");
            }
            if (EPI == EPI_F16 && p.colstats) {
                // flush the column statistics of the PREVIOUS tile (every epilogue warp has passed the barrier above, so they are
                // complete): thread = column folds the four lane quadrants and issues ONE integer RED per (group, column, moment)
                if (tl > 0) cs_flush(p, s_cs, s_csmeta, (tl - 1) & 1u, (int)threadIdx.x - 128);
                // this tile's record (read by the flush behind the NEXT barrier): statistics group of each lane quadrant, slot, first column
                int* meta = s_csmeta + (tl & 1u) * 8;
                if (half == 0 && lane == 0) {
                    const int grp = p.conv_mode ? c3 : ((int)(((int64_t)(c3 * p.nb1 + c2) * p.M + c1) / p.cs_rows));
                    meta[4 + q] = box_ok ? grp : -1;
                }
                if (e == 0 && lane == 0) { meta[0] = (p.ksplit == 1) ? 1 : 0; meta[1] = tc.m_tile & 3; meta[2] = ntile0; }
            }
            float ln_a = 0.f, ln_b = 0.f, ln_c = 0.f;                 // row mode: (mean, rstd) of this row; column mode: (c1, row bias)
            if (LNF && p.ln_mode == 1 && row_ok) {
                const int64_t m = (int64_t)tc.z * p.M + c1 + lane;
                const double inv = (double)p.ln_inv_dim * (1.0 / 1048576.0);
                const double mu = (double)(long long)p.ln_stats[2 * m] * inv, qq = (double)(long long)p.ln_stats[2 * m + 1] * inv;
                ln_a = (float)mu;
                ln_b = rsqrtf((float)fmax(qq - mu * mu, 0.0) + p.ln_eps);
            } else if (LNF && p.ln_mode == 2 && row_ok) {
                ln_a = p.ln_c1[c1 + lane];
                ln_c = p.ln_rowbias ? p.ln_rowbias[c1 + lane] : 0.f;
            }
            const float* sbr = sb + ((stage_b2 && (r0 + lane) / (p.BW * p.BH) > 0) ? 256 : 0);
            const bool b2_global = p.bias2 && !stage_b2;                           // generic (plain GEMM / many images per tile) path
            // residual chunk of this warp's first chunk: in flight while the MMA main loop runs
            if (use_res && half < nchunks && lane == 0 && (p.ksplit == 1)) {
                const uint32_t b = slot & 1u;
                mbar_expect_tx(&rbar[b], 2048);
                tma_load_4d(rsm + b * 2048, &tmR, &rbar[b], otile0 + half * 32, c1, c2, c3);
            }
            mbar_wait(&tmem_full[as], (tl >> 1) & 1u);
            tc_fence_after();
            if (tl == 0 && e == 0 && lane == 0) TRACE(5);
            const uint32_t taddr_row = tmem_base + as * TMEM_STAGE_COLS + ((uint32_t)(q * 32) << 16);

            bool released = false;
            const int out_tile = (tc.z * p.n_tiles + tc.n_tile) * p.m_tiles + tc.m_tile;
            // split-K partial tiles in global memory (L2-resident), one private slice per (tile, split), laid out
            // [32-col chunk][float4 j][row]: coalesced.  Plain stores, no atomics: the last arriver adds the slices in the
            // fixed order 0 .. ksplit-1, so the result does not depend on which CTA finishes last (run-to-run deterministic).
            const int64_t tile_f4 = (int64_t)(p.BN / 32) * (8 * 128);
            float4* wsp = reinterpret_cast<float4*>(p.workspace) + (int64_t)out_tile * p.ksplit * tile_f4 + r0 + lane;
            if (p.ksplit > 1) {
                // ---- split-K: every CTA of the tile stores its partial accumulator, then bumps the tile counter;
                // the LAST arriver owns the reduction and the epilogue.
                const int acc_chunks = p.BN / 32;
                float4* mine = wsp + (int64_t)tc.ks * tile_f4;
#pragma unroll 1
                for (int c = half; c < acc_chunks; c += 2) {
                    float v[32];
                    tc_ld32(taddr_row + (uint32_t)(c * 32), v);
                    float4* dst = mine + (int64_t)c * (8 * 128);
#pragma unroll
                    for (int j = 0; j < 8; j++) __stcg(dst + j * 128, make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]));
                }
                released = true;                                     // the accumulator stage is free again
                tc_fence_before();
                __syncwarp();
                if (lane == 0) { if (PAIR && rank != 0) mbar_arrive_remote(&tmem_empty[as], 0); else mbar_arrive(&tmem_empty[as]); }
                __threadfence();
                asm volatile("bar.sync 1, %0;" :: "n"(32 * kEpiWarps) : "memory");
                if (e == 0 && lane == 0) {
                    const int old = atomicAdd(&p.counters[out_tile], 1);
                    const int last = (old == p.ksplit - 1);
                    if (last) p.counters[out_tile] = 0;             // self-reset for the next launch
                    __threadfence();
                    *s_flag = last;
                }
                asm volatile("bar.sync 1, %0;" :: "n"(32 * kEpiWarps) : "memory");
                const int last = *s_flag;
                asm volatile("bar.sync 1, %0;" :: "n"(32 * kEpiWarps) : "memory");      // everyone has read the flag before the next tile rewrites it
                if (!last) continue;
                if (EPI == EPI_F16 && p.colstats && e == 0 && lane == 0) s_csmeta[(tl & 1u) * 8] = 1;      // this CTA owns the epilogue of the tile
                if (use_res && half < nchunks && lane == 0) {
                    const uint32_t b = slot & 1u;
                    mbar_expect_tx(&rbar[b], 2048);
                    tma_load_4d(rsm + b * 2048, &tmR, &rbar[b], otile0 + half * 32, c1, c2, c3);
                }
            }

            // software-pipelined TMEM reads: chunk c+2 is in flight while chunk c is processed
            const bool pipelined = (EPI != EPI_GEGLU) && (p.ksplit == 1);
            float nx[32];
            if (pipelined && half < nchunks) tc_ld32_nowait(taddr_row + (uint32_t)(half * ACC_PER_CHUNK), nx);
#pragma unroll 1
            for (int c = half; c < nchunks; c += 2, slot++) {
                const uint32_t b = slot & 1u;
                // prefetch the residual of this warp's NEXT chunk of the tile (its buffer was consumed one chunk ago)
                if (use_res && c + 2 < nchunks && lane == 0) {
                    const uint32_t nb = b ^ 1u;
                    mbar_expect_tx(&rbar[nb], 2048);
                    tma_load_4d(rsm + nb * 2048, &tmR, &rbar[nb], otile0 + (c + 2) * 32, c1, c2, c3);
                }
                float f[32];
                const int acol0 = c * ACC_PER_CHUNK;                    // accumulator column of the chunk inside the tile
                const int ocol0 = otile0 + c * 32;                      // output column
                const float4* bp = reinterpret_cast<const float4*>(sbr + acol0);
                if (EPI == EPI_GEGLU) {
                    float v[32], g[32];
                    tc_ld32(taddr_row + (uint32_t)acol0, v);
                    tc_ld32(taddr_row + (uint32_t)(acol0 + 32), g);
                    if (LNF && p.ln_mode == 1) {
                        const float* c1p = sb + 256 + acol0;
#pragma unroll
                        for (int j = 0; j < 32; j++) { v[j] = ln_b * (v[j] - ln_a * c1p[j]); g[j] = ln_b * (g[j] - ln_a * c1p[32 + j]); }
                    }
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        const float4 b0 = bp[j], b1 = bp[8 + j];
                        // (value, gate) pairs are interleaved along N
                        f[2 * j] = fmaf(v[4 * j], p.alpha, b0.x) * gelu_f(fmaf(v[4 * j + 1], p.alpha, b0.y));
                        f[2 * j + 1] = fmaf(v[4 * j + 2], p.alpha, b0.z) * gelu_f(fmaf(v[4 * j + 3], p.alpha, b0.w));
                        f[16 + 2 * j] = fmaf(g[4 * j], p.alpha, b1.x) * gelu_f(fmaf(g[4 * j + 1], p.alpha, b1.y));
                        f[16 + 2 * j + 1] = fmaf(g[4 * j + 2], p.alpha, b1.z) * gelu_f(fmaf(g[4 * j + 3], p.alpha, b1.w));
                    }
                } else {
                    if (p.ksplit > 1) {
                        // last arriver: add the ksplit partial slices of this chunk in split order (fixed summation order).
                        // The loads of THREE slices are in flight together (the loop is L2-latency bound: one slice per
                        // round trip made a ksplit = 6 reduction cost 6 us); the additions stay in slice order.
                        const float4* src = wsp + (int64_t)c * (8 * 128);
#pragma unroll
                        for (int j = 0; j < 8; j++) {
                            const float4 u = __ldcg(src + j * 128);
                            f[4 * j] = u.x; f[4 * j + 1] = u.y; f[4 * j + 2] = u.z; f[4 * j + 3] = u.w;
                        }
                        int sp = 1;
#pragma unroll 1
                        for (; sp + 2 < p.ksplit; sp += 3) {
                            const float4* s2 = src + (int64_t)sp * tile_f4;
                            float4 u0[8], u1[8], u2[8];
#pragma unroll
                            for (int j = 0; j < 8; j++) {
                                u0[j] = __ldcg(s2 + j * 128); u1[j] = __ldcg(s2 + tile_f4 + j * 128); u2[j] = __ldcg(s2 + 2 * tile_f4 + j * 128);
                            }
#pragma unroll
                            for (int j = 0; j < 8; j++) {
                                f[4 * j] = ((f[4 * j] + u0[j].x) + u1[j].x) + u2[j].x; f[4 * j + 1] = ((f[4 * j + 1] + u0[j].y) + u1[j].y) + u2[j].y;
                                f[4 * j + 2] = ((f[4 * j + 2] + u0[j].z) + u1[j].z) + u2[j].z; f[4 * j + 3] = ((f[4 * j + 3] + u0[j].w) + u1[j].w) + u2[j].w;
                            }
                        }
#pragma unroll 1
                        for (; sp < p.ksplit; sp++) {
                            const float4* s2 = src + (int64_t)sp * tile_f4;
#pragma unroll
                            for (int j = 0; j < 8; j++) {
                                const float4 u = __ldcg(s2 + j * 128);
                                f[4 * j] += u.x; f[4 * j + 1] += u.y; f[4 * j + 2] += u.z; f[4 * j + 3] += u.w;
                            }
                        }
                    } else {
                        tc_wait_ld();
                        reg_fence(nx);
                        if (tl == 0 && e == 0 && lane == 0 && c == 0) TRACE(8);
#pragma unroll
                        for (int j = 0; j < 32; j++) f[j] = nx[j];
                        reg_fence(f);                                   // the copy is complete before nx is handed to the next load
                        if (c + 2 < nchunks) tc_ld32_nowait(taddr_row + (uint32_t)((c + 2) * ACC_PER_CHUNK), nx);
                    }
                    if (LNF && p.ln_mode == 1) {
                        const float* c1p = sb + 256 + acol0;
#pragma unroll
                        for (int j = 0; j < 32; j++) f[j] = ln_b * (f[j] - ln_a * c1p[j]);
                    } else if (LNF && p.ln_mode == 2) {
                        const float* mp = sb + 256 + acol0;
                        const float* rp = sb + 512 + acol0;
#pragma unroll
                        for (int j = 0; j < 32; j++) f[j] = fmaf(rp[j], f[j] - mp[j] * ln_a, ln_c);
                    }
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        const float4 bb = bp[j];
                        f[4 * j] = fmaf(f[4 * j], p.alpha, bb.x); f[4 * j + 1] = fmaf(f[4 * j + 1], p.alpha, bb.y);
                        f[4 * j + 2] = fmaf(f[4 * j + 2], p.alpha, bb.z); f[4 * j + 3] = fmaf(f[4 * j + 3], p.alpha, bb.w);
                    }
                    if (b2_global) {
#pragma unroll
                        for (int j = 0; j < 32; j++) {
                            const int n = ocol0 + j;
                            if (n < p.N) f[j] += p.bias2[b2row * p.N + n];
                        }
                    }
                    if (p.act == ACT_SILU) {
#pragma unroll
                        for (int j = 0; j < 32; j++) f[j] = silu_f(f[j]);
                    } else if (p.act == ACT_GELU) {
#pragma unroll
                        for (int j = 0; j < 32; j++) f[j] = gelu_f(f[j]);
                    }
                }
                if (c + 2 >= nchunks && !released) {       // last TMEM read of this warp for the tile: release the accumulator stage
                    released = true;
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) { if (PAIR && rank != 0) mbar_arrive_remote(&tmem_empty[as], 0); else mbar_arrive(&tmem_empty[as]); }
                }
                if (p.direct) {
                    // ---- slow path: unaligned strides; masked per-thread stores
                    if (row_ok) {
                        const int nvalid = min(32, ((EPI == EPI_GEGLU) ? (p.N >> 1) : p.N) - ocol0);
#pragma unroll
                        for (int j = 0; j < 32; j++) {                  // fully unrolled: a dynamic index would push f[] into local memory
                            if (j >= nvalid) break;
                            float x = f[j];
                            if (p.residual) x += act_to_f(p.residual[d_res + ocol0 + j]);
                            if (EPI == EPI_F32) reinterpret_cast<float*>(p.C)[d_row + ocol0 + j] = x;
                            else reinterpret_cast<act_t*>(p.C)[d_row + ocol0 + j] = f_to_act(x);
                        }
                    }
                    continue;
                }
                if (!box_ok) continue;
                if (use_res) {
                    mbar_wait(&rbar[b], (rph >> b) & 1u);
                    rph ^= (1u << b);
                    const uint8_t* rrow = rsm + b * 2048 + lane * 64;
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const uint4 u = *reinterpret_cast<const uint4*>(rrow + ((j ^ rsw) << 4));
                        const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                        for (int k = 0; k < 4; k++) {
                            const float2 t2 = act2_to_f2(*reinterpret_cast<const act2_t*>(&w4[k]));
                            f[8 * j + 2 * k] += t2.x; f[8 * j + 2 * k + 1] += t2.y;
                        }
                    }
                }
                if (tl == 0 && e == 0 && lane == 0 && c == 0) TRACE(9);
                // staging buffer b was last read by the TMA store issued two chunks ago
                if (lane == 0) bulk_wait_read<1>();
                __syncwarp();
                uint8_t* srow = stg + b * STG_BYTES + row_off;
                if (EPI == EPI_F32) {
#pragma unroll
                    for (int j = 0; j < 8; j++)
                        *reinterpret_cast<float4*>(srow + ((j ^ sw) << 4)) = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
                } else {
                    float rs = 0.f, rq = 0.f;
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        uint4 pk;
                        pk.x = pack_act(f[8 * j], f[8 * j + 1]); pk.y = pack_act(f[8 * j + 2], f[8 * j + 3]);
                        pk.z = pack_act(f[8 * j + 4], f[8 * j + 5]); pk.w = pack_act(f[8 * j + 6], f[8 * j + 7]);
                        *reinterpret_cast<uint4*>(srow + ((j ^ sw) << 4)) = pk;
                        if (LNF && p.rowstats_out) {
                            const uint32_t w4[4] = {pk.x, pk.y, pk.z, pk.w};
#pragma unroll
                            for (int k = 0; k < 4; k++) {
                                const float2 t2 = act2_to_f2(*reinterpret_cast<const act2_t*>(&w4[k]));
                                const bool k0 = ocol0 + 8 * j + 2 * k < p.N, k1 = ocol0 + 8 * j + 2 * k + 1 < p.N;
                                if (k0) { rs += t2.x; rq = fmaf(t2.x, t2.x, rq); }
                                if (k1) { rs += t2.y; rq = fmaf(t2.y, t2.y, rq); }
                            }
                        }
                    }
                    if (LNF && p.rowstats_out && row_ok) {
                        // LayerNorm statistics of the consumer: one pair of fixed-point integer REDs per (row, 32-column chunk)
                        unsigned long long* dst = p.rowstats_out + 2 * ((int64_t)tc.z * p.M + c1 + lane);
                        atomicAdd(dst, (unsigned long long)__float2ll_rn(rs * 1048576.0f));
                        atomicAdd(dst + 1, (unsigned long long)__float2ll_rn(rq * 1048576.0f));
                    }
                }
                if (tl == 0 && e == 0 && lane == 0 && c == 0) TRACE(10);
                fence_async_smem();
                __syncwarp();
                if (tl == 0 && e == 0 && lane == 0 && c == 0) TRACE(11);
                if (lane == 0) {
                    tma_store_4d(&tmC, stg + b * STG_BYTES, ocol0, c1, c2, c3);
                    bulk_commit();
                }
                if (tl == 0 && e == 0 && lane == 0 && c == 0) TRACE(12);
                if (EPI == EPI_F16 && p.colstats) {
                    // ---- GroupNorm statistics of the consumer, taken from the staged fp16 chunk (exactly the values that reach
                    // memory): lane = (column pair, row parity) walks 16 rows of the swizzled 32 x 32 staging tile (conflict-free:
                    // a half-warp reads one 64-byte row), the two row parities meet by one shuffle, and the 32 column totals leave
                    // as fixed-point integer REDs (order-independent => deterministic) into colstats[image][column][2].
                    const uint8_t* sbuf = stg + b * STG_BYTES;
                    const int cp = lane & 15, par = lane >> 4;
                    const int nrows = p.conv_mode ? 32 : min(32, p.M - c1);
                    float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
#pragma unroll 4
                    for (int i = 0; i < 16; i++) {
                        const int r = 2 * i + par;
                        if (r < nrows) {
                            const uint32_t w = *reinterpret_cast<const uint32_t*>(sbuf + r * 64 + ((((cp >> 2) ^ ((r >> 1) & 3))) << 4) + ((cp & 3) << 2));
                            const float2 v = act2_to_f2(*reinterpret_cast<const act2_t*>(&w));
                            s0 += v.x; s1 += v.y; q0 = fmaf(v.x, v.x, q0); q1 = fmaf(v.y, v.y, q1);
                        }
                    }
                    s0 += __shfl_xor_sync(0xffffffffu, s0, 16); s1 += __shfl_xor_sync(0xffffffffu, s1, 16);
                    q0 += __shfl_xor_sync(0xffffffffu, q0, 16); q1 += __shfl_xor_sync(0xffffffffu, q1, 16);
                    if (par == 0) {
                        // stage (sum, sumsq) of columns 2cp, 2cp+1 of this 32-row quadrant: single writer per (quadrant, column)
                        float* dst = s_cs + ((((tl & 1u) * 4 + q) * 256) + (c * 32 + 2 * cp)) * 2;
                        *reinterpret_cast<float4*>(dst) = make_float4(s0, q0, s1, q1);
                    }
                }
            }
            if (!released) {                                 // warps with no chunk in this tile (BN == 32, half == 1)
                tc_fence_before();
                __syncwarp();
                if (lane == 0) { if (PAIR && rank != 0) mbar_arrive_remote(&tmem_empty[as], 0); else mbar_arrive(&tmem_empty[as]); }
            }
        }
        if (EPI == EPI_F16 && p.colstats && tl > 0) {
            asm volatile("bar.sync 2, %0;" :: "n"(32 * kEpiWarps) : "memory");
            cs_flush(p, s_cs, s_csmeta, (tl - 1) & 1u, (int)threadIdx.x - 128);
        }
        if (e == 0 && lane == 0) TRACE(6);
        if (lane == 0) bulk_wait_read<0>();                  // smem must stay valid until the last TMA store has read it
        if (e == 0 && lane == 0) TRACE(7);
    }
    tc_fence_before();
    if (PAIR) cluster_sync_all();            // neither CTA may retire (or free TMEM) while the pair's MMAs / remote arrivals are in flight
    else __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "n"(2 * TMEM_STAGE_COLS));
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "n"(2 * TMEM_STAGE_COLS));
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

// 4-D tensor map: dims (innermost first), byte strides for dims 1..3, box, element strides
static int make_map(CUtensorMap* m, CUtensorMapDataType dt, CUtensorMapSwizzle swz, const void* base, const uint64_t dims[4],
                    const uint64_t strides_bytes[3], const uint32_t box[4], const uint32_t estr[4]) {
    EncodeTiledFn enc = get_encode();
    if (!enc) { set_error("cuTensorMapEncodeTiled unavailable"); return DWG_ERR_CUDA; }
    cuuint64_t d[4] = {dims[0], dims[1], dims[2], dims[3]};
    cuuint64_t s[3] = {strides_bytes[0], strides_bytes[1], strides_bytes[2]};
    cuuint32_t b[4] = {box[0], box[1], box[2], box[3]};
    cuuint32_t e[4] = {estr[0], estr[1], estr[2], estr[3]};
    CUresult r = enc(m, dt, 4, const_cast<void*>(base), d, s, b, e, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d): dims=(%llu,%llu,%llu,%llu) strides=(%llu,%llu,%llu) box=(%u,%u,%u,%u)", (int)r,
                  (unsigned long long)d[0], (unsigned long long)d[1], (unsigned long long)d[2], (unsigned long long)d[3],
                  (unsigned long long)s[0], (unsigned long long)s[1], (unsigned long long)s[2], b[0], b[1], b[2], b[3]);
        return DWG_ERR_INVALID;
    }
    return DWG_OK;
}
static int make_map_act(CUtensorMap* m, const void* base, const uint64_t dims[4], const uint64_t strides_bytes[3],
                         const uint32_t box[4], const uint32_t estr[4]) {
    return make_map(m, DWG_TMAP_ACT, CU_TENSOR_MAP_SWIZZLE_128B, base, dims, strides_bytes, box, estr);
}

static int g_num_sms = 0;
static float* g_ws = nullptr;            // split-K fp32 partial-tile slices
static size_t g_ws_bytes = 0;
static int* g_counters = nullptr;        // split-K arrival counters (zero when idle)
constexpr int kMaxCounters = 1 << 16;
constexpr size_t kSmemBudget = 232448 - 1024;       // opt-in maximum minus the 1024-byte alignment slack

static int g_plan_cs = 0;                  // the launch being planned accumulates column statistics (16 KB + 64 B more shared memory)
static size_t smem_fixed(int epi, int has_res) {
    const size_t stg = (epi == EPI_F32) ? 4096 : 2048;
    return (size_t)kEpiWarps * 2 * stg + (has_res ? (size_t)kEpiWarps * 2 * 2048 : 0) + 2 * 3 * 256 * 4 + (g_plan_cs ? 2 * 4 * 256 * 2 * 4 + 64 : 0) +
           (2 * kMaxStages + 4 + 2 * kEpiWarps + 2 * kMaxAStages) * 8 + 64;
}

// Tile / split selection: a small analytic model of one CTA's critical path, in SM cycles.
//   k-iteration = max(tensor pipe 2*BN, smem operand read 128 + BN, L2->SM feed of the CTAs running together)
//   tile        = k-iterations + epilogue (TMEM drain + math + staging, two warps per lane quadrant)
struct Plan { int BN, ks, stages, pair, halo; };      // halo: wanted for 3x3 stride-1 convolutions (the entry point checks legality)
static unsigned long long* g_trace = nullptr;
// Split-K scratch is per "lane": GEMMs enqueued on two streams that may run concurrently (ControlNet
// beside the UNet encoder) must not share tile accumulators / counters.  dwg_gemm_set_lane() selects the
// half the following launches use (host-side state, baked into the launch parameters).
constexpr int kLanes = 2;
static int g_lane = 0;
static int g_force_bn = 0, g_force_ks = 0;        // tuning override (dwg_gemm_tune), 0 = automatic
static int g_force_halo = -1;                     // dwg_gemm_tune_halo: -1 = automatic, 0 = never, 1 = whenever legal
static int g_halo_bo = 0;                         // descriptor base-offset field of the shifted halo windows (debug knob; 0 is what sm_100a wants)
static int g_last_halo = 0;
static int g_force_pair = -1;                     // dwg_gemm_tune_pair: -1 = automatic, 0 = never, 1 = whenever legal
static Plan g_last_plan = {0, 0, 0, 0, 0};
static int g_last_key[6] = {0, 0, 0, 0, 0, 0};
// Measured plans for the shapes of the SDS step on a 148-SM B200 (tools/gemm_autotune.py writes the
// table: cold weights, warm activations, CUDA-graph replays); anything else falls back to the model.
struct TunedPlan { int m_tiles, nz, N, iters, epi, has_res, BN, ks, pair, halo; };
static const TunedPlan kTuned[] = {
#include "gemm_plan_table.inc"
    {0, 0, 0, 0, 0, 0, 0, 0, 0, 0}};
// CTA pairs need an even number of 128-row tiles (no phantom half) and a tile width whose halves are whole swizzle groups
static bool pair_legal(int m_tiles, int BN) { return (m_tiles % 2) == 0 && (BN % 32) == 0 && BN >= 32; }
static int stages_for(int BN, int pair, int epi, int has_res) {
    int stages = (int)((kSmemBudget - smem_fixed(epi, has_res)) / (A_BYTES + (size_t)(pair ? BN / 2 : BN) * 128));
    return stages > kMaxStages ? kMaxStages : stages;
}
static int g_counter_cap = 0;              // counters available to the launch being planned (set by the entry point from its workspace)
static Plan plan_tiles_auto(int m_tiles, int nz, int N, int iters, int epi, int has_res, int64_t ws_cap_floats);
static Plan plan_tiles(int m_tiles, int nz, int N, int iters, int epi, int has_res, int64_t ws_cap_floats) {
    Plan pl = plan_tiles_auto(m_tiles, nz, N, iters, epi, has_res, ws_cap_floats);
    g_last_key[0] = m_tiles; g_last_key[1] = nz; g_last_key[2] = N; g_last_key[3] = iters; g_last_key[4] = epi; g_last_key[5] = has_res;
    if (g_num_sms == kNumSMs && g_force_bn == 0) {
        for (const TunedPlan* t = kTuned; t->BN; t++) {
            if (t->m_tiles == m_tiles && t->nz == nz && t->N == N && t->iters == iters && t->epi == epi && t->has_res == has_res) {
                const int64_t tiles = (int64_t)m_tiles * ((N + t->BN - 1) / t->BN) * nz;
                const int pair = (t->pair && pair_legal(m_tiles, t->BN)) ? 1 : 0;
                const int stages = stages_for(t->BN, pair, epi, has_res);
                if (stages >= 2 && (t->ks == 1 || (tiles <= g_counter_cap && tiles * t->ks * 128 * (int64_t)t->BN <= ws_cap_floats && t->ks <= iters)))
                    pl = {t->BN, t->ks, stages, pair, t->halo && t->ks == 1};
                break;
            }
        }
    }
    if (g_force_bn > 0) {
        const int gran = (epi == EPI_GEGLU) ? 64 : 32;
        int BN = (g_force_bn + gran - 1) / gran * gran;
        if (BN > 256) BN = 256;
        int ks = g_force_ks > 0 ? g_force_ks : 1;
        if (epi == EPI_GEGLU) ks = 1;
        if (ks > iters) ks = iters;
        const int64_t tiles = (int64_t)m_tiles * ((N + BN - 1) / BN) * nz;
        if (ks > 1 && (tiles > g_counter_cap || tiles * ks * 128 * (int64_t)BN > ws_cap_floats)) ks = 1;
        pl = {BN, ks, stages_for(BN, 0, epi, has_res), 0, 0};
    }
    if (g_force_pair >= 0) {
        pl.pair = (g_force_pair == 1 && pair_legal(m_tiles, pl.BN)) ? 1 : 0;
        pl.stages = stages_for(pl.BN, pl.pair, epi, has_res);
    }
    g_last_plan = pl;
    return pl;
}
static Plan plan_tiles_auto(int m_tiles, int nz, int N, int iters, int epi, int has_res, int64_t ws_cap_floats) {
    const int gran = (epi == EPI_GEGLU) ? 64 : 32;
    Plan best = {gran, 1, 2, 0, 0};
    double best_t = 1e30;
    const size_t fixed = smem_fixed(epi, has_res);
    for (int BN = gran; BN <= 256; BN += gran) {
        const int n_tiles = (N + BN - 1) / BN;
        const size_t stage_bytes = A_BYTES + (size_t)BN * 128;
        int stages = (int)((kSmemBudget - fixed) / stage_bytes);
        if (stages > kMaxStages) stages = kMaxStages;
        if (stages < 2) continue;
        const int64_t tiles = (int64_t)m_tiles * n_tiles * nz;
        for (int ks = 1; ks <= 32; ks++) {
            if (ks > 1 && (epi == EPI_GEGLU || iters / ks < 2 || tiles * ks > 2 * g_num_sms)) break;
            if (ks > 1 && (tiles > g_counter_cap || tiles * ks * 128 * (int64_t)BN > ws_cap_floats)) break;
            const int64_t ctas = tiles * ks;
            const double active = (double)(ctas < g_num_sms ? ctas : g_num_sms);
            const double feed = (double)stage_bytes * active / 6000.0;          // ~6.3 KB/cycle chip-wide TMA throughput
            double t_iter = 2.0 * BN;
            if (128.0 + BN > t_iter) t_iter = 128.0 + BN;
            if (feed > t_iter) t_iter = feed;
            const double k_iters = (double)((iters + ks - 1) / ks);
            const double t_main = k_iters * t_iter;
            const double epi_cols = (epi == EPI_GEGLU) ? 14.0 : 7.0;            // cycles per accumulator column per warp pair
            const double t_epi = 300.0 + BN * epi_cols + (has_res ? 200.0 : 0.0);
            const double waves = (double)((ctas + g_num_sms - 1) / g_num_sms);
            double t_tile = t_main > t_epi ? t_main : t_epi;                    // steady state of a persistent CTA
            double t = 2500.0 + 1500.0 /* first TMA */ + (waves - 1.0) * t_tile + t_main + t_epi;
            if (ks > 1) t += 1500.0 + BN * 8.0;                                 // vector-atomic partial adds + counter + tile read-back
            if (t < best_t) { best_t = t; best = {BN, ks, stages, 0, 0}; }
        }
    }
    // shapes outside the measured table: pairs / halos pay once the launch is bound by the L2 -> SM operand feed,
    // i.e. every SM has a tile and the K loop is long enough to amortise the cluster hand-shakes
    const int64_t tiles = (int64_t)m_tiles * ((N + best.BN - 1) / best.BN) * nz;
    if (best.ks == 1 && tiles >= g_num_sms) {
        best.halo = 1;
        if (pair_legal(m_tiles, best.BN) && iters >= 8) { best.pair = 1; best.stages = stages_for(best.BN, 1, epi, has_res); }
    }
    return best;
}

template <int EPI, bool PAIR, bool LNF>
static cudaError_t launch_one(dim3 grid, size_t smem, cudaStream_t st, const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC,
                              const CUtensorMap& tmR, const Params& p) {
    static bool attr_done = false;
    if (!attr_done) {
        cudaFuncSetAttribute(gemm_kernel<EPI, PAIR, LNF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kSmemBudget + 1024));
        attr_done = true;
    }
    if (!PAIR) return launch_pdl(gemm_kernel<EPI, PAIR, LNF>, grid, dim3(kThreads), smem, st, tmA, tmB, tmC, tmR, p);
    // CTA pair = cluster of 2 (same TPC) + programmatic dependent launch
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 2 : 1;
    return cudaLaunchKernelEx(&cfg, gemm_kernel<EPI, PAIR, LNF>, tmA, tmB, tmC, tmR, p);
}

static int ensure_sms() {
    if (!g_num_sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (g_num_sms <= 0) g_num_sms = kNumSMs;
    }
    return DWG_OK;
}
static int ensure_globals() {
    ensure_sms();
    if (!g_counters) {
        // NOTE: allocated outside of stream capture (the first, un-captured warm-up call creates it)
        if (cudaMalloc(&g_counters, sizeof(int) * kMaxCounters) != cudaSuccess) { g_counters = nullptr; set_error("split-K counter allocation failed"); return DWG_ERR_CUDA; }
        cudaMemset(g_counters, 0, sizeof(int) * kMaxCounters);
        g_ws_bytes = (size_t)64 << 20;
        if (cudaMalloc(&g_ws, g_ws_bytes) != cudaSuccess) { g_ws = nullptr; g_ws_bytes = 0; set_error("split-K workspace allocation failed"); return DWG_ERR_CUDA; }
        cudaMemset(g_ws, 0, g_ws_bytes);                 // tile accumulators are zero when idle (the last arriver re-zeroes what it reads)
        cudaDeviceSynchronize();
    }
    return DWG_OK;
}

static int launch(const CUtensorMap& tmA, CUtensorMap& tmB_out, const CUtensorMap& tmC, const CUtensorMap& tmR, Params& p, int epi, cudaStream_t st) {
    const bool pair = p.m_sched != p.m_tiles;
    const size_t b_bytes = (size_t)(pair ? p.BN / 2 : p.BN) * 128;
    const size_t smem = 1024 + (p.halo ? (size_t)p.a_stages * HALO_BYTES + (size_t)p.stages * b_bytes : (size_t)p.stages * (A_BYTES + b_bytes)) +
                        smem_fixed(epi, p.has_res);
    const bool lnf = p.ln_mode != 0 || p.rowstats_out != nullptr;
    const dim3 grid = pair ? dim3(2 * (p.total_tiles < g_num_sms / 2 ? p.total_tiles : g_num_sms / 2)) : dim3(p.total_tiles < g_num_sms ? p.total_tiles : g_num_sms);
#define DWG_LAUNCH(E, PR, LN) launch_one<E, PR, LN>(grid, smem, st, tmA, tmB_out, tmC, tmR, p)
    if (lnf) {                                   // LayerNorm folding / row statistics: fp16-output epilogues only
        if (epi == EPI_F16) { if (pair) DWG_LAUNCH(EPI_F16, true, true); else DWG_LAUNCH(EPI_F16, false, true); }
        else if (epi == EPI_GEGLU) { if (pair) DWG_LAUNCH(EPI_GEGLU, true, true); else DWG_LAUNCH(EPI_GEGLU, false, true); }
        else { set_error("LayerNorm folding needs an fp16 output"); return DWG_ERR_INVALID; }
    } else if (pair) {
        if (epi == EPI_F16) DWG_LAUNCH(EPI_F16, true, false);
        else if (epi == EPI_F32) DWG_LAUNCH(EPI_F32, true, false);
        else DWG_LAUNCH(EPI_GEGLU, true, false);
    } else {
        if (epi == EPI_F16) DWG_LAUNCH(EPI_F16, false, false);
        else if (epi == EPI_F32) DWG_LAUNCH(EPI_F32, false, false);
        else DWG_LAUNCH(EPI_GEGLU, false, false);
    }
#undef DWG_LAUNCH
    return check_launch("tcgen05 gemm");
}

static bool aligned16(const void* ptr, int64_t esz, int64_t s0, int64_t s1, int64_t s2) {
    return ((uintptr_t)ptr & 15) == 0 && (s0 * esz) % 16 == 0 && (s1 * esz) % 16 == 0 && (s2 * esz) % 16 == 0;
}

}  // namespace gemm
}  // namespace dwg

using namespace dwg;
using namespace dwg::gemm;

// D[b2][b1][M,N] = act(alpha * A[b2][b1][M,K] @ B[b2][b1][N,K]^T + bias[N] + bias2) + residual
// All strides in ELEMENTS.  A/B bf16, K contiguous.  out_f16: 1 -> bf16 C, 0 -> fp32 C.
// Split-K scratch of one launch: counters (zero when idle, self-resetting) + fp32 partial-tile slices.
struct Scratch { float* data; int64_t floats; int* counters; int n_counters; };
constexpr int64_t kWsCounterBytes = 128 << 10;            // 32768 counters at the head of a caller-provided workspace
static int scratch_from_caller(void* ws, int64_t bytes, Scratch& s) {
    if (!ws || bytes < kWsCounterBytes + (1 << 20) || (reinterpret_cast<uintptr_t>(ws) & 255)) {
        set_error("workspace must be 256-byte aligned and hold at least dwg_gemm_workspace_bytes() bytes");
        return DWG_ERR_CAPACITY;
    }
    s.counters = reinterpret_cast<int*>(ws); s.n_counters = (int)(kWsCounterBytes / 4);
    s.data = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + kWsCounterBytes); s.floats = (bytes - kWsCounterBytes) / 4;
    return DWG_OK;
}
static int scratch_from_library(Scratch& s) {
    int rc = ensure_globals();
    if (rc) return rc;
    s.data = g_ws + (size_t)g_lane * (g_ws_bytes / 4 / kLanes); s.floats = (int64_t)(g_ws_bytes / 4 / kLanes);
    s.counters = g_counters + g_lane * (kMaxCounters / kLanes); s.n_counters = kMaxCounters / kLanes;
    return DWG_OK;
}

struct LnFold {                      // optional LayerNorm folding / row statistics of one plain GEMM (see Params)
    int mode = 0; const void* stats = nullptr; const float* c1 = nullptr; const float* rowbias = nullptr; int dim = 0; float eps = 0.f;
    void* rowstats_out = nullptr;
};
static int gemm_impl(const void* A, int64_t lda, int64_t a_b1, int64_t a_b2,
                             const void* B, int64_t ldb, int64_t b_b1, int64_t b_b2,
                             void* C, int64_t ldc, int64_t c_b1, int64_t c_b2, int out_f16,
                             int M, int N, int K, int nb1, int nb2,
                             const float* bias, const float* bias2, int bias2_rows_per,
                             const void* residual, int64_t ldr, int64_t r_b1, int64_t r_b2,
                             float alpha, int act, const Scratch& ws, void* colstats, int cs_rows, void* stream, const LnFold& ln = LnFold()) {
    DWG_REQUIRE(A && B && C, "null pointer");
    DWG_REQUIRE(M > 0 && N > 0 && K > 0 && nb1 > 0 && nb2 > 0, "bad sizes");
    DWG_REQUIRE((lda % 8) == 0 && (ldb % 8) == 0 && (a_b1 % 8) == 0 && (a_b2 % 8) == 0 && (b_b1 % 8) == 0 && (b_b2 % 8) == 0,
                "A/B strides must be multiples of 8 elements (16 bytes) for TMA");
    DWG_REQUIRE(((uintptr_t)A & 15) == 0 && ((uintptr_t)B & 15) == 0, "A/B must be 16-byte aligned");
    DWG_REQUIRE(act != ACT_GEGLU || (N % 2 == 0 && out_f16 && !residual), "GEGLU epilogue: even N, 16-bit output, no residual");
    DWG_REQUIRE(act != ACT_GEGLU || !bias || ((uintptr_t)bias & 15) == 0, "GEGLU bias must be 16-byte aligned");
    int rc = ensure_sms();
    if (rc) return rc;
    g_counter_cap = ws.n_counters;
    g_plan_cs = colstats ? 1 : 0;
    const int epi = act == ACT_GEGLU ? EPI_GEGLU : (out_f16 ? EPI_F16 : EPI_F32);
    const int64_t esz = out_f16 ? 2 : 4;
    if (nb1 == 1) c_b1 = ldc * (int64_t)M;
    if (nb2 == 1) c_b2 = c_b1 * nb1;
    if (residual && nb1 == 1) r_b1 = ldr * (int64_t)M;
    if (residual && nb2 == 1) r_b2 = r_b1 * nb1;
    Params p = {};
    p.direct = !aligned16(C, esz, ldc, c_b1, c_b2) || (residual && !aligned16(residual, 2, ldr, r_b1, r_b2)) ||
               (bias && ((uintptr_t)bias & 15)) || (bias2 && (((uintptr_t)bias2 & 15) || (N % 4))) || (act == ACT_GEGLU && (N % 64));
    DWG_REQUIRE(!(p.direct && act == ACT_GEGLU), "GEGLU epilogue needs 16-byte aligned output rows and N % 64 == 0");
    p.has_res = (residual && !p.direct) ? 1 : 0;
    p.M = M; p.N = N; p.K = K; p.taps_w = 1; p.taps_h = 1; p.k_chunks = (K + BK - 1) / BK; p.nb1 = nb1; p.conv_mode = 0;
    p.a_bcast1 = (nb1 > 1 && a_b1 == 0) ? 1 : 0;
    p.a_bytes = A_BYTES;
    p.C = C; p.ldc = ldc; p.c_b1 = c_b1; p.c_b2 = c_b2;
    p.bias = bias; p.bias2 = bias2; p.bias2_rows_per = bias2_rows_per;
    p.residual = reinterpret_cast<const act_t*>(residual); p.ldr = ldr; p.r_b1 = r_b1; p.r_b2 = r_b2;
    p.alpha = alpha; p.act = act;
    p.m_tiles = (M + BM - 1) / BM; p.nz = nb1 * nb2;
    const Plan pl = plan_tiles(p.m_tiles, p.nz, N, p.k_chunks, epi, p.has_res, ws.floats);
    p.BN = pl.BN; p.ksplit = pl.ks; p.stages = pl.stages;
    p.n_tiles = (N + p.BN - 1) / p.BN;
    p.m_sched = pl.pair ? p.m_tiles / 2 : p.m_tiles;
    p.total_tiles = p.m_sched * p.n_tiles * p.nz * p.ksplit;
    p.workspace = ws.data; p.counters = ws.counters; p.trace = g_trace;
    p.colstats = reinterpret_cast<unsigned long long*>(colstats); p.cs_rows = cs_rows;
    p.ln_mode = ln.mode; p.ln_stats = reinterpret_cast<const unsigned long long*>(ln.stats); p.ln_c1 = ln.c1; p.ln_rowbias = ln.rowbias;
    p.ln_inv_dim = ln.dim > 0 ? 1.0f / (float)ln.dim : 0.f; p.ln_eps = ln.eps;
    p.rowstats_out = reinterpret_cast<unsigned long long*>(ln.rowstats_out);
    DWG_REQUIRE(ln.mode >= 0 && ln.mode <= 2, "LayerNorm folding: mode must be 0, 1 or 2");
    DWG_REQUIRE(ln.mode == 0 || (ln.stats && ln.c1 && ln.dim > 0), "LayerNorm folding: statistics, c1 and the feature count are required");
    DWG_REQUIRE(ln.mode == 0 || nb2 == 1, "LayerNorm folding: at most one batch dimension");
    DWG_REQUIRE(ln.mode == 0 || !p.direct, "LayerNorm folding needs the TMA-store epilogue (16-byte aligned output / residual rows and bias)");
    DWG_REQUIRE(!ln.rowstats_out || (epi == EPI_F16 && nb2 == 1 && !p.direct), "row statistics need the fp16 TMA-store epilogue");
    DWG_REQUIRE(!colstats || (epi == EPI_F16 && !p.direct && cs_rows > 0 && (cs_rows % 32) == 0 && ((int64_t)M % cs_rows == 0 || nb1 * nb2 == 1)),
                "column statistics need the fp16 TMA-store epilogue and statistics groups that are whole multiples of 32 rows");

    CUtensorMap tmA, tmB, tmC, tmR;
    const uint32_t ones[4] = {1, 1, 1, 1};
    {
        // a_b1 == 0 with nb1 > 1: A is broadcast over the first batch dimension (the producer then always loads batch 0)
        const bool bc = nb1 > 1 && a_b1 == 0;
        const uint64_t dims[4] = {(uint64_t)K, (uint64_t)M, (uint64_t)(bc ? 1 : nb1), (uint64_t)nb2};
        const uint64_t str[3] = {(uint64_t)lda * 2, (uint64_t)(nb1 > 1 && !bc ? a_b1 : lda * (int64_t)M) * 2, (uint64_t)(nb2 > 1 ? a_b2 : lda * (int64_t)M) * 2};
        const uint32_t box[4] = {BK, BM, 1, 1};
        rc = make_map_act(&tmA, A, dims, str, box, ones);
        if (rc) return rc;
    }
    {
        const uint64_t dims[4] = {(uint64_t)K, (uint64_t)N, (uint64_t)nb1, (uint64_t)nb2};
        const uint64_t str[3] = {(uint64_t)ldb * 2, (uint64_t)(nb1 > 1 ? b_b1 : ldb * (int64_t)N) * 2, (uint64_t)(nb2 > 1 ? b_b2 : ldb * (int64_t)N) * 2};
        const uint32_t box[4] = {BK, (uint32_t)(pl.pair ? p.BN / 2 : p.BN), 1, 1};
        rc = make_map_act(&tmB, B, dims, str, box, ones);
        if (rc) return rc;
    }
    tmC = tmA; tmR = tmA;                                   // placeholders on the direct path (never dereferenced)
    if (!p.direct) {
        const uint64_t No = (uint64_t)(act == ACT_GEGLU ? N / 2 : N);
        const uint64_t dims[4] = {No, (uint64_t)M, (uint64_t)nb1, (uint64_t)nb2};
        const uint64_t str[3] = {(uint64_t)(ldc * esz), (uint64_t)(c_b1 * esz), (uint64_t)(c_b2 * esz)};
        const uint32_t box[4] = {32, 32, 1, 1};
        rc = make_map(&tmC, out_f16 ? DWG_TMAP_ACT : CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                      out_f16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, C, dims, str, box, ones);
        if (rc) return rc;
        if (p.has_res) {
            const uint64_t rstr[3] = {(uint64_t)ldr * 2, (uint64_t)r_b1 * 2, (uint64_t)r_b2 * 2};
            rc = make_map(&tmR, DWG_TMAP_ACT, CU_TENSOR_MAP_SWIZZLE_64B, residual, dims, rstr, box, ones);
            if (rc) return rc;
        }
    }
    return launch(tmA, tmB, tmC, tmR, p, epi, (cudaStream_t)stream);
}

// NHWC convolution as implicit GEMM.  x [Nimg,H,W,Cin] bf16 (Cin % 8 == 0), w [Cout,kh,kw,Cin] bf16,
// y [Nimg,Ho,Wo,Cout] (bf16 or fp32).  Padding is zero-fill (pad_h/pad_w applied on the top/left;
// the bottom/right extent follows from Ho/Wo, which covers SD's asymmetric (0,1,0,1) padding).
static int conv_impl(const void* x, const void* w, void* y, int out_f16,
                                    int Nimg, int H, int W, int Cin, int Cout, int ksize, int stride,
                                    int pad_h, int pad_w, int Ho, int Wo,
                                    const float* bias, const float* bias2_per_image,
                                    const void* residual, int act, const Scratch& ws, void* colstats, void* stream) {
    DWG_REQUIRE(x && w && y, "null pointer");
    DWG_REQUIRE(Cin % 8 == 0, "Cin must be a multiple of 8 (pad the channels)");
    DWG_REQUIRE(stride == 1 || stride == 2, "stride must be 1 or 2");
    DWG_REQUIRE(ksize == 1 || ksize == 3, "ksize must be 1 or 3");
    DWG_REQUIRE(act != ACT_GEGLU, "GEGLU is a linear-layer epilogue");
    int rc = ensure_sms();
    if (rc) return rc;
    g_counter_cap = ws.n_counters;
    g_plan_cs = colstats ? 1 : 0;
    // pixel box of one 128-row tile
    int BW = 1;
    while (BW * 2 <= Wo && BW * 2 <= BM) BW *= 2;
    DWG_REQUIRE(Wo % BW == 0 || Wo < BW * 2, "unsupported output width");
    int BH = 1;
    while (BH * 2 <= Ho && BW * BH * 2 <= BM) BH *= 2;
    int BNI = BM / (BW * BH);
    if (BNI > Nimg) BNI = 1 << (31 - __builtin_clz(Nimg));         // largest power of two <= Nimg
    const int tiles_w = (Wo + BW - 1) / BW, tiles_h = (Ho + BH - 1) / BH, tiles_n = (Nimg + BNI - 1) / BNI;
    const int epi = out_f16 ? EPI_F16 : EPI_F32;
    const int64_t esz = out_f16 ? 2 : 4;
    Params p = {};
    p.direct = (((uintptr_t)y & 15) || (Cout * esz) % 16 || (residual && (((uintptr_t)residual & 15) || (Cout * 2) % 16)) ||
                (bias && ((uintptr_t)bias & 15)) || (bias2_per_image && (((uintptr_t)bias2_per_image & 15) || (Cout % 4)))) ? 1 : 0;
    if (BW * BH * BNI < 32 && tiles_n > 1) p.direct = 1;       // a warp's 32-row store box would reach into the next image group
    p.has_res = (residual && !p.direct) ? 1 : 0;
    p.M = Nimg * Ho * Wo; p.N = Cout; p.K = Cin; p.taps_w = ksize; p.taps_h = ksize; p.k_chunks = (Cin + BK - 1) / BK;
    p.a_bytes = (uint32_t)(BW * BH * BNI * BK * 2);
    p.a_bcast1 = 0;
    p.nb1 = 1; p.conv_mode = 1; p.Ho = Ho; p.Wo = Wo; p.BH = BH; p.BW = BW; p.BNI = BNI; p.Nimg = Nimg; p.tiles_w = tiles_w; p.tiles_h = tiles_h;
    p.stride = stride; p.pad_h = pad_h; p.pad_w = pad_w;
    p.ebw = BW < 32 ? BW : 32;
    p.ebh = (32 / p.ebw) < BH ? (32 / p.ebw) : BH;
    int ebn = 32 / (p.ebw * p.ebh);
    p.C = y; p.ldc = Cout;
    p.bias = bias; p.bias2 = bias2_per_image; p.bias2_rows_per = bias2_per_image ? Ho * Wo : 0;
    p.residual = reinterpret_cast<const act_t*>(residual); p.ldr = Cout;
    p.alpha = 1.0f; p.act = act;
    p.m_tiles = tiles_w * tiles_h * tiles_n; p.nz = 1;
    const int iters = ksize * ksize * p.k_chunks;
    const Plan pl = plan_tiles(p.m_tiles, 1, Cout, iters, epi, p.has_res, ws.floats);
    p.BN = pl.BN; p.ksplit = pl.ks; p.stages = pl.stages;
    // ---- halo mode: 3x3 / stride 1 / "same" on images that tile into 8 x 16 pixel blocks, no split-K
    p.halo = 0;
    const bool halo_legal = ksize == 3 && stride == 1 && pad_h == 1 && pad_w == 1 && Ho == H && Wo == W && (Ho % HALO_BH) == 0 &&
                            (Wo % HALO_BW) == 0 && pl.ks == 1 && !p.direct;
    if (halo_legal && g_force_halo != 0 && (g_force_halo == 1 || pl.halo)) {
        const size_t b_bytes = (size_t)(pl.pair ? pl.BN / 2 : pl.BN) * 128;
        const int a_stages = 2;
        int64_t bst = ((int64_t)kSmemBudget - (int64_t)smem_fixed(epi, p.has_res) - (int64_t)a_stages * HALO_BYTES) / (int64_t)b_bytes;
        if (bst > kMaxStages) bst = kMaxStages;
        const int halo_m_tiles = (Wo / HALO_BW) * (Ho / HALO_BH) * Nimg;
        if (bst >= 3 && (!pl.pair || (halo_m_tiles % 2) == 0)) {         // the pair plan was made for an even tile count
            p.halo = 1; p.a_stages = a_stages; p.halo_bo = g_halo_bo; p.stages = (int)bst;
            BW = HALO_BW; BH = HALO_BH; BNI = 1;
            p.BH = BH; p.BW = BW; p.BNI = BNI; p.tiles_w = Wo / BW; p.tiles_h = Ho / BH;
            p.ebw = BW; p.ebh = 32 / BW; ebn = 1;
            p.m_tiles = p.tiles_w * p.tiles_h * Nimg;
        }
    }
    g_last_halo = p.halo;
    p.n_tiles = (Cout + p.BN - 1) / p.BN;
    p.m_sched = pl.pair ? p.m_tiles / 2 : p.m_tiles;
    p.total_tiles = p.m_sched * p.n_tiles * p.ksplit;
    p.workspace = ws.data; p.counters = ws.counters; p.trace = g_trace;
    p.colstats = reinterpret_cast<unsigned long long*>(colstats); p.cs_rows = 0;
    p.ln_mode = 0; p.ln_stats = nullptr; p.ln_c1 = nullptr; p.ln_rowbias = nullptr; p.ln_inv_dim = 0.f; p.ln_eps = 0.f; p.rowstats_out = nullptr;
    DWG_REQUIRE(!colstats || (epi == EPI_F16 && !p.direct && (Wo % BW) == 0 && (Ho % BH) == 0 && (Nimg % BNI) == 0 && ((BW * BH) % 32) == 0),
                "column statistics need the fp16 TMA-store epilogue and exactly tiled images with >= 32 pixels per tile image");

    CUtensorMap tmA, tmB, tmC, tmR;
    {
        const uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)Nimg};
        const uint64_t str[3] = {(uint64_t)Cin * 2, (uint64_t)W * Cin * 2, (uint64_t)H * W * Cin * 2};
        const uint32_t box_t[4] = {BK, (uint32_t)((BW - 1) * stride + 1), (uint32_t)((BH - 1) * stride + 1), (uint32_t)BNI};
        const uint32_t box_h[4] = {BK, (uint32_t)HALO_LINE_PIX, 1, 1};          // one halo line
        const uint32_t* box = p.halo ? box_h : box_t;
        const uint32_t es[4] = {1, (uint32_t)stride, (uint32_t)stride, 1};
        rc = make_map_act(&tmA, x, dims, str, box, es);
        if (rc) return rc;
    }
    const uint32_t ones[4] = {1, 1, 1, 1};
    {
        const int taps = ksize * ksize;
        const uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)taps, (uint64_t)Cout, 1};
        const uint64_t str[3] = {(uint64_t)Cin * 2, (uint64_t)taps * Cin * 2, (uint64_t)Cout * taps * Cin * 2};
        const uint32_t box[4] = {BK, 1, (uint32_t)(pl.pair ? p.BN / 2 : p.BN), 1};
        rc = make_map_act(&tmB, w, dims, str, box, ones);
        if (rc) return rc;
    }
    tmC = tmA; tmR = tmA;
    if (!p.direct) {
        const uint64_t dims[4] = {(uint64_t)Cout, (uint64_t)Wo, (uint64_t)Ho, (uint64_t)Nimg};
        const uint64_t str[3] = {(uint64_t)(Cout * esz), (uint64_t)((int64_t)Wo * Cout * esz), (uint64_t)((int64_t)Ho * Wo * Cout * esz)};
        const uint32_t box[4] = {32, (uint32_t)p.ebw, (uint32_t)p.ebh, (uint32_t)ebn};
        rc = make_map(&tmC, out_f16 ? DWG_TMAP_ACT : CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                      out_f16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, y, dims, str, box, ones);
        if (rc) return rc;
        if (p.has_res) {
            const uint64_t rstr[3] = {(uint64_t)Cout * 2, (uint64_t)Wo * Cout * 2, (uint64_t)Ho * Wo * Cout * 2};
            rc = make_map(&tmR, DWG_TMAP_ACT, CU_TENSOR_MAP_SWIZZLE_64B, residual, dims, rstr, box, ones);
            if (rc) return rc;
        }
    }
    return launch(tmA, tmB, tmC, tmR, p, epi, (cudaStream_t)stream);
}

// ---- public entry points.  The *_ws forms take the split-K scratch from the caller (no library-owned memory, safe for
// concurrent streams / host threads as long as each uses its own workspace); the plain forms are conveniences over a
// per-process scratch with two lanes (dwg_gemm_set_lane).
extern "C" int64_t dwg_gemm_workspace_bytes(void) { return kWsCounterBytes + ((int64_t)32 << 20); }

extern "C" int dwg_gemm_f16_ws(const void* A, int64_t lda, int64_t a_b1, int64_t a_b2, const void* B, int64_t ldb, int64_t b_b1, int64_t b_b2,
                               void* C, int64_t ldc, int64_t c_b1, int64_t c_b2, int out_f16, int M, int N, int K, int nb1, int nb2,
                               const float* bias, const float* bias2, int bias2_rows_per, const void* residual, int64_t ldr, int64_t r_b1,
                               int64_t r_b2, float alpha, int act, void* workspace, int64_t workspace_bytes, void* colstats, int colstats_rows, void* stream) {
    Scratch s;
    int rc = scratch_from_caller(workspace, workspace_bytes, s);
    if (rc) return rc;
    return gemm_impl(A, lda, a_b1, a_b2, B, ldb, b_b1, b_b2, C, ldc, c_b1, c_b2, out_f16, M, N, K, nb1, nb2, bias, bias2, bias2_rows_per, residual, ldr,
                     r_b1, r_b2, alpha, act, s, colstats, colstats_rows, stream);
}
extern "C" int dwg_gemm_f16(const void* A, int64_t lda, int64_t a_b1, int64_t a_b2, const void* B, int64_t ldb, int64_t b_b1, int64_t b_b2,
                            void* C, int64_t ldc, int64_t c_b1, int64_t c_b2, int out_f16, int M, int N, int K, int nb1, int nb2,
                            const float* bias, const float* bias2, int bias2_rows_per, const void* residual, int64_t ldr, int64_t r_b1,
                            int64_t r_b2, float alpha, int act, void* stream) {
    Scratch s;
    int rc = scratch_from_library(s);
    if (rc) return rc;
    return gemm_impl(A, lda, a_b1, a_b2, B, ldb, b_b1, b_b2, C, ldc, c_b1, c_b2, out_f16, M, N, K, nb1, nb2, bias, bias2, bias2_rows_per, residual, ldr,
                     r_b1, r_b2, alpha, act, s, nullptr, 0, stream);
}
/* dwg_gemm_f16_ws plus (a) LayerNorm of the activation operand folded into the epilogue and (b) LayerNorm statistics of the
 * OUTPUT rows for the next consumer -- see include/dwg.h. */
extern "C" int dwg_gemm_f16_ln(const void* A, int64_t lda, int64_t a_b, const void* B, int64_t ldb, int64_t b_b, void* C, int64_t ldc, int64_t c_b,
                               int M, int N, int K, int nb, const float* bias, const void* residual, int64_t ldr, int64_t r_b, int act,
                               int ln_mode, const void* ln_stats, const float* ln_c1, const float* ln_rowbias, int ln_dim, float ln_eps,
                               void* rowstats_out, void* workspace, int64_t workspace_bytes, void* colstats, int colstats_rows, void* stream) {
    Scratch s;
    int rc = scratch_from_caller(workspace, workspace_bytes, s);
    if (rc) return rc;
    LnFold ln;
    ln.mode = ln_mode; ln.stats = ln_stats; ln.c1 = ln_c1; ln.rowbias = ln_rowbias; ln.dim = ln_dim; ln.eps = ln_eps; ln.rowstats_out = rowstats_out;
    return gemm_impl(A, lda, a_b, 0, B, ldb, b_b, 0, C, ldc, c_b, 0, 1, M, N, K, nb, 1, bias, nullptr, 0, residual, ldr, r_b, 0, 1.0f, act, s,
                     colstats, colstats_rows, stream, ln);
}
extern "C" int dwg_conv2d_nhwc_f16_ws(const void* x, const void* w, void* y, int out_f16, int Nimg, int H, int W, int Cin, int Cout, int ksize,
                                      int stride, int pad_h, int pad_w, int Ho, int Wo, const float* bias, const float* bias2_per_image,
                                      const void* residual, int act, void* workspace, int64_t workspace_bytes, void* colstats, void* stream) {
    Scratch s;
    int rc = scratch_from_caller(workspace, workspace_bytes, s);
    if (rc) return rc;
    return conv_impl(x, w, y, out_f16, Nimg, H, W, Cin, Cout, ksize, stride, pad_h, pad_w, Ho, Wo, bias, bias2_per_image, residual, act, s, colstats, stream);
}
extern "C" int dwg_conv2d_nhwc_f16(const void* x, const void* w, void* y, int out_f16, int Nimg, int H, int W, int Cin, int Cout, int ksize,
                                   int stride, int pad_h, int pad_w, int Ho, int Wo, const float* bias, const float* bias2_per_image,
                                   const void* residual, int act, void* stream) {
    Scratch s;
    int rc = scratch_from_library(s);
    if (rc) return rc;
    return conv_impl(x, w, y, out_f16, Nimg, H, W, Cin, Cout, ksize, stride, pad_h, pad_w, Ho, Wo, bias, bias2_per_image, residual, act, s, nullptr, stream);
}

/* Tuning / introspection (tools/gemm_sweep.py): force the tile width and split count of the next
 * launches (0, 0 = automatic), and read back the plan of the last launch as (BN, ksplit, stages). */
extern "C" int dwg_gemm_tune(int force_bn, int force_ks) { g_force_bn = force_bn; g_force_ks = force_ks; return DWG_OK; }
extern "C" int dwg_gemm_last_plan(int* out3) {
    DWG_REQUIRE(out3, "null pointer");
    out3[0] = g_last_plan.BN; out3[1] = g_last_plan.ks; out3[2] = g_last_plan.stages;
    return DWG_OK;
}
/* CTA-pair (tcgen05 cta_group::2) mode of the following launches: -1 automatic, 0 never, 1 whenever legal;
 * dwg_gemm_last_pair() reads back what the last launch used. */
/* halo mode of 3x3 stride-1 convolutions: -1 automatic, 0 never, 1 whenever legal; base_offset_mode = 1 puts
 * (start >> 7) & 7 into the descriptors' base-offset field (wrong on sm_100a, kept for tools/halo_probe.py), 0 = default. */
extern "C" int dwg_gemm_tune_halo(int mode, int base_offset_mode) {
    g_force_halo = mode < 0 ? -1 : (mode ? 1 : 0);
    g_halo_bo = base_offset_mode ? 1 : 0;
    return DWG_OK;
}
extern "C" int dwg_gemm_last_halo(void) { return g_last_halo; }
extern "C" int dwg_gemm_tune_pair(int mode) { g_force_pair = mode < 0 ? -1 : (mode ? 1 : 0); return DWG_OK; }
extern "C" int dwg_gemm_last_pair(void) { return g_last_plan.pair; }
/* the planner key of the last launch: (m_tiles, nz, N, k_iterations, epilogue kind, has_residual) */
extern "C" int dwg_gemm_last_key(int* out6) {
    DWG_REQUIRE(out6, "null pointer");
    for (int i = 0; i < 6; i++) out6[i] = g_last_key[i];
    return DWG_OK;
}
/* Debug: device pointer to 8 x u64 that CTA 0 of every following launch fills with globaltimer stamps
 * (0 kernel entry, 1 setup done, 2 after griddepcontrol.wait, 3 first operand stage landed, 4 last MMA
 * committed, 5 accumulator visible to the epilogue, 6 last store issued, 7 stores drained); null = off. */
extern "C" int dwg_gemm_set_lane(int lane) {
    DWG_REQUIRE(lane >= 0 && lane < kLanes, "lane must be 0 or 1");
    g_lane = lane;
    return DWG_OK;
}
extern "C" int dwg_gemm_trace(void* dev_u64x8) { g_trace = reinterpret_cast<unsigned long long*>(dev_u64x8); return DWG_OK; }
