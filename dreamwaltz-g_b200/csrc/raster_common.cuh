// Shared definitions for the tile rasteriser (R11-R13).
#pragma once
#include "common.cuh"

namespace dwg {
namespace raster {

constexpr int TILE = 16;                 // 16x16 pixel tiles, 256-thread CTAs (upstream BLOCK_X/Y)
constexpr int TILE_PIX = TILE * TILE;
constexpr int CHUNK = 256;               // instances staged per TMA bulk copy in the blend loops
constexpr int SORT_CHUNK = 2048;         // keys sorted in shared memory at a time

// One (tile, Gaussian) instance in depth order, gathered once by the sort kernel and then
// streamed (contiguously, 48 B, 16-byte aligned) by the forward and backward blend kernels.
struct __align__(16) Rec {
    float x, y;              // pixel-space mean
    float cx, cy, cz;        // conic
    float op;                // opacity
    float r, g, b;           // colour
    float depth;             // view-space z
    uint32_t idx;            // Gaussian index
    uint32_t pad;
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

struct BinView {
    uint32_t* tile_count;    // [T]   instances per tile (atomic histogram)
    uint32_t* tile_fill;     // [T]   scatter cursors
    uint32_t* tile_start;    // [T+1] exclusive scan
    uint2* ranges;           // [T]   [start,end) per tile (upstream identifyTileRanges)
    uint64_t* inst_key;      // [P]   (depth bits << 32 | idx), unsorted then sorted per tile
    uint64_t* inst_tmp;      // [P]   merge ping-pong buffer (tiles with > SORT_CHUNK instances)
    uint64_t* keys_out;      // [P]   (tile << 32 | depth bits)  == upstream sorted keys
    uint32_t* vals_out;      // [P]   Gaussian idx               == upstream sorted values
    Rec* recs;               // [P]
    __host__ __device__ static size_t header_bytes(int T) {
        return align256(sizeof(uint32_t) * T) * 2 + align256(sizeof(uint32_t) * (T + 1)) + align256(sizeof(uint2) * T);
    }
    __host__ __device__ static size_t bytes(int64_t P, int T) {
        return header_bytes(T) + align256(sizeof(uint64_t) * P) * 3 + align256(sizeof(uint32_t) * P) + align256(sizeof(Rec) * P);
    }
    __host__ __device__ BinView(void* base, int64_t P, int T) {
        char* p = (char*)base;
        tile_count = (uint32_t*)p; p += align256(sizeof(uint32_t) * T);
        tile_fill = (uint32_t*)p; p += align256(sizeof(uint32_t) * T);
        tile_start = (uint32_t*)p; p += align256(sizeof(uint32_t) * (T + 1));
        ranges = (uint2*)p; p += align256(sizeof(uint2) * T);
        inst_key = (uint64_t*)p; p += align256(sizeof(uint64_t) * P);
        inst_tmp = (uint64_t*)p; p += align256(sizeof(uint64_t) * P);
        keys_out = (uint64_t*)p; p += align256(sizeof(uint64_t) * P);
        vals_out = (uint32_t*)p; p += align256(sizeof(uint32_t) * P);
        recs = (Rec*)p;
    }
};

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
    if (x < -87.0f) return 0.0f;
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
    return __fmul_rn(y, s);
}

// alpha of one instance at pixel (pxf,pyf); false when the instance is skipped (power > 0 or
// alpha < 1/255).  Operation order is part of the spec (see oracle eval_alpha).
__device__ __forceinline__ bool eval_alpha(const Rec& rc, float pxf, float pyf, float& alpha, float& G, float& dx, float& dy) {
    dx = __fsub_rn(rc.x, pxf);
    dy = __fsub_rn(rc.y, pyf);
    const float a = __fmul_rn(__fmul_rn(rc.cx, dx), dx);
    const float b = __fmul_rn(__fmul_rn(rc.cz, dy), dy);
    const float c = __fmul_rn(__fmul_rn(rc.cy, dx), dy);
    const float power = __fsub_rn(__fmul_rn(-0.5f, __fadd_rn(a, b)), c);
    if (power > 0.0f) return false;
    G = spec_expf(power);
    alpha = fminf(0.99f, __fmul_rn(rc.op, G));
    return !(alpha < 1.0f / 255.0f);
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
