// Shared definitions for the tile rasteriser (R11-R13).
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"

namespace dwg {
namespace raster {

constexpr int TILE = 16;                 // 16x16 pixel tiles, 256-thread CTAs (upstream BLOCK_X/Y)
constexpr int TILE_PIX = TILE * TILE;
constexpr int QUAD = 8;                  // a 16x16 tile is blended by FOUR 64-thread CTAs, one per 8x8 quadrant: the heavy
constexpr int QPIX = QUAD * QUAD;        // silhouette tiles (thousands of instances) bound the kernel, and their eight
                                         // warps now run on four different SMs instead of sharing one SM's issue slots
constexpr int CHUNK = 128;               // instances staged per TMA bulk copy in the blend loops
constexpr int SORT_CHUNK = 2048;         // keys sorted in shared memory at a time

// One (tile, Gaussian) instance in depth order, gathered once by the pack kernel and then
// streamed (contiguously, 48 B, 16-byte aligned) by the forward and backward blend kernels.
// The first 16 bytes carry everything a warp needs to decide that none of its pixels can be
// touched (conservative screen-space extent of the alpha >= 1/255 ellipse, as two halves).
struct __align__(16) Rec {
    float x, y;              // pixel-space mean
    uint32_t ext;            // half2: (ext_x, ext_y), rounded up, conservative
    uint32_t idx;            // Gaussian index
    float cx, cy, cz;        // conic
    float op;                // opacity
    float r, g, b;           // colour
    float depth;             // view-space z
};
static_assert(sizeof(Rec) == 48, "Rec must be 48 bytes");

__host__ __device__ inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

struct GeomView {
    float2* xy; float* depth; float* cov3D; float4* conic_opacity; int4* rect; uint32_t* tiles_touched;
    __host__ __device__ static size_t bytes(int64_t N) {
        return align256(sizeof(float2) * N) + align256(sizeof(float) * N) + align256(sizeof(float) * 6 * N) +
               align256(sizeof(float4) * N) + align256(sizeof(int4) * N) + align256(sizeof(uint32_t) * N);
    }
    __host__ __device__ GeomView(void* base, int64_t N) {
        char* p = (char*)base;
        xy = (float2*)p; p += align256(sizeof(float2) * N);
        depth = (float*)p; p += align256(sizeof(float) * N);
        cov3D = (float*)p; p += align256(sizeof(float) * 6 * N);
        conic_opacity = (float4*)p; p += align256(sizeof(float4) * N);
        rect = (int4*)p; p += align256(sizeof(int4) * N);
        tiles_touched = (uint32_t*)p;
    }
};

// The per-tile histogram counters and scatter cursors are split into BIN_SUB sub-counters (sub = Gaussian index & 7): a
// silhouette tile receives 5-9k same-address atomics, which serialise at the L2 (~45 us per pass); eight addresses per tile
// cut that chain eight-fold.  Slots inside a tile come out in a different order -- the per-tile sort fixes the order anyway.
constexpr int BIN_SUB = 8;
struct BinView {
    uint32_t* tile_count;    // [T][BIN_SUB] instances per (tile, sub-counter)  (atomic histogram)
    uint32_t* tile_fill;     // [T][BIN_SUB] scatter cursors
    uint32_t* sub_start;     // [T][BIN_SUB] first slot of each sub-counter's share of the tile segment
    uint32_t* tile_start;    // [T+1] exclusive scan
    uint2* ranges;           // [T]   [start,end) per tile (upstream identifyTileRanges)
    uint32_t* n_runs;        // [4]   {number of sort runs, ...}
    uint2* runs;             // [R]   (first instance, count) of each <= SORT_CHUNK run
    uint64_t* inst_key;      // [P]   (depth bits << 32 | idx), unsorted then sorted per tile
    uint64_t* inst_tmp;      // [P]   merge ping-pong buffer (tiles with > SORT_CHUNK instances)
    uint32_t* inst_tile;     // [P]   tile of each instance slot
    uint64_t* keys_out;      // [P]   (tile << 32 | depth bits)  == upstream sorted keys
    uint32_t* vals_out;      // [P]   Gaussian idx               == upstream sorted values
    Rec* recs;               // [P]
    __host__ __device__ static int64_t run_cap(int64_t P, int T) { return T + P / SORT_CHUNK + 1; }
    __host__ __device__ static size_t header_bytes(int T) {
        return align256(sizeof(uint32_t) * T * BIN_SUB) * 3 + align256(sizeof(uint32_t) * (T + 1)) + align256(sizeof(uint2) * T) + 256;
    }
    __host__ __device__ static size_t bytes(int64_t P, int T) {
        return header_bytes(T) + align256(sizeof(uint2) * run_cap(P, T)) + align256(sizeof(uint64_t) * P) * 3 +
               align256(sizeof(uint32_t) * P) * 2 + align256(sizeof(Rec) * P);
    }
    __host__ __device__ BinView(void* base, int64_t P, int T) {
        char* p = (char*)base;
        tile_count = (uint32_t*)p; p += align256(sizeof(uint32_t) * T * BIN_SUB);
        tile_fill = (uint32_t*)p; p += align256(sizeof(uint32_t) * T * BIN_SUB);
        sub_start = (uint32_t*)p; p += align256(sizeof(uint32_t) * T * BIN_SUB);
        tile_start = (uint32_t*)p; p += align256(sizeof(uint32_t) * (T + 1));
        ranges = (uint2*)p; p += align256(sizeof(uint2) * T);
        n_runs = (uint32_t*)p; p += 256;
        runs = (uint2*)p; p += align256(sizeof(uint2) * run_cap(P, T));
        inst_key = (uint64_t*)p; p += align256(sizeof(uint64_t) * P);
        inst_tmp = (uint64_t*)p; p += align256(sizeof(uint64_t) * P);
        inst_tile = (uint32_t*)p; p += align256(sizeof(uint32_t) * P);
        keys_out = (uint64_t*)p; p += align256(sizeof(uint64_t) * P);
        vals_out = (uint32_t*)p; p += align256(sizeof(uint32_t) * P);
        recs = (Rec*)p;
    }
};

// number of pairwise merge passes a tile segment of n instances needs after the run sort
__host__ __device__ inline int merge_passes(uint32_t n) {
    int p = 0;
    for (uint32_t len = SORT_CHUNK; len < n; len <<= 1) p++;
    return p;
}
constexpr int MAX_MERGE_PASSES = 6;      // tiles up to SORT_CHUNK * 64 = 131072 instances

struct ImgView {
    float* final_T; uint32_t* n_contrib;
    __host__ __device__ static size_t bytes(int H, int W) { return align256(sizeof(float) * H * W) * 2; }
    __host__ __device__ ImgView(void* base, int H, int W) {
        final_T = (float*)base;
        n_contrib = (uint32_t*)((char*)base + align256(sizeof(float) * H * W));
    }
};

// exp(x), x <= 0, as an explicit sequence of IEEE fp32 operations (never contracted): identical,
// bit for bit, to spec_expf() in oracle/oracle_c.c.  ~1 ulp.
__device__ __forceinline__ float spec_expf(float x) {
    const bool tiny = x < -87.0f;                  // result 0 (selected at the end: no branch inside the blend loops)
    const float t = __fmul_rn(x, 1.44269504088896341f);
    const float n = rintf(t);
    float r = __fmaf_rn(n, -0.693145751953125f, x);
    r = __fmaf_rn(n, -1.42860682030941723e-6f, r);
    float p = 1.9875691500e-4f;
    p = __fmaf_rn(p, r, 1.3981999507e-3f);
    p = __fmaf_rn(p, r, 8.3334519073e-3f);
    p = __fmaf_rn(p, r, 4.1665795894e-2f);
    p = __fmaf_rn(p, r, 1.6666665459e-1f);
    p = __fmaf_rn(p, r, 5.0000001201e-1f);
    const float z = __fmul_rn(r, r);
    float y = __fmaf_rn(p, z, r);
    y = __fadd_rn(y, 1.0f);
    const float s = __int_as_float(((int)n + 127) << 23);
    return tiny ? 0.0f : __fmul_rn(y, s);
}

// alpha of one instance at pixel (pxf,pyf); false when the instance is skipped (power > 0 or
// alpha < 1/255).  Operation order is part of the spec (see oracle eval_alpha).
// power < -5.6 implies exp(power) < 0.0037 < 1/255 even with a few ulp of error, hence
// alpha = min(0.99, op*G) <= G < 1/255 for any op <= 1: rejecting early is exactly equivalent
// (for op > 1 the early test is skipped).
__device__ __forceinline__ bool eval_alpha(const Rec& rc, float pxf, float pyf, float& alpha, float& G, float& dx, float& dy) {
    dx = __fsub_rn(rc.x, pxf);
    dy = __fsub_rn(rc.y, pyf);
    const float a = __fmul_rn(__fmul_rn(rc.cx, dx), dx);
    const float b = __fmul_rn(__fmul_rn(rc.cz, dy), dy);
    const float c = __fmul_rn(__fmul_rn(rc.cy, dx), dy);
    const float power = __fsub_rn(__fmul_rn(-0.5f, __fadd_rn(a, b)), c);
    if (power > 0.0f) return false;
    if (power < -5.6f && rc.op <= 1.0f) return false;
    G = spec_expf(power);
    alpha = fminf(0.99f, __fmul_rn(rc.op, G));
    return !(alpha < 1.0f / 255.0f);
}

// Branch-free variant for batched evaluation (several instances in flight per thread): identical operations and
// results for every accepted instance; rejected ones evaluate the exponential of a harmless argument.
__device__ __forceinline__ bool eval_alpha_nb(const Rec& rc, float pxf, float pyf, float& alpha, float& G, float& dx, float& dy) {
    dx = __fsub_rn(rc.x, pxf);
    dy = __fsub_rn(rc.y, pyf);
    const float a = __fmul_rn(__fmul_rn(rc.cx, dx), dx);
    const float b = __fmul_rn(__fmul_rn(rc.cz, dy), dy);
    const float c = __fmul_rn(__fmul_rn(rc.cy, dx), dy);
    const float power = __fsub_rn(__fmul_rn(-0.5f, __fadd_rn(a, b)), c);
    const bool rej = (power > 0.0f) | ((power < -5.6f) & (rc.op <= 1.0f));
    G = spec_expf(rej ? -1.0f : power);
    alpha = fminf(0.99f, __fmul_rn(rc.op, G));
    return !rej & !(alpha < 1.0f / 255.0f);
}

// The same evaluation in two stages, so that a warp can skip the exponential of an instance that NO lane keeps after the
// (cheap) power test: eval_power + eval_finish == eval_alpha_nb for every lane that keeps the instance.
__device__ __forceinline__ bool eval_power(const Rec& rc, float pxf, float pyf, float& power, float& dx, float& dy) {
    dx = __fsub_rn(rc.x, pxf);
    dy = __fsub_rn(rc.y, pyf);
    const float a = __fmul_rn(__fmul_rn(rc.cx, dx), dx);
    const float b = __fmul_rn(__fmul_rn(rc.cz, dy), dy);
    const float c = __fmul_rn(__fmul_rn(rc.cy, dx), dy);
    power = __fsub_rn(__fmul_rn(-0.5f, __fadd_rn(a, b)), c);
    return !((power > 0.0f) | ((power < -5.6f) & (rc.op <= 1.0f)));
}
__device__ __forceinline__ bool eval_finish(float op, float power, bool live, float& alpha, float& G) {
    G = spec_expf(live ? power : -1.0f);
    alpha = fminf(0.99f, __fmul_rn(op, G));
    return live & !(alpha < 1.0f / 255.0f);
}

// Conservative test: can ANY pixel of the strip [x0,x1] x [y0,y1] get alpha >= 1/255 ?
__device__ __forceinline__ bool strip_may_touch(float rx, float ry, uint32_t ext, float x0, float x1, float y0, float y1) {
    const float2 e = __half22float2(*reinterpret_cast<const __half2*>(&ext));
    return !(rx + e.x < x0 || rx - e.x > x1 || ry + e.y < y0 || ry - e.y > y1);
}

// ---- mbarrier + 1-D bulk TMA (cp.async.bulk) helpers ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
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
// global -> shared bulk copy, completion signalled on `bar` (bytes multiple of 16, both 16-B aligned)
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

}  // namespace raster
}  // namespace dwg
