// Rasteriser stage 3 (R12/R13): per-tile front-to-back alpha blending, and the C-ABI forward.
//
// One 256-thread CTA per 16x16 tile, one pixel per thread (the per-pixel transmittance chain is
// evaluated sequentially in source order so that final_T and n_contrib are bit-exact with the
// oracle).  The tile's depth-sorted 48-byte instance records are contiguous in HBM; they are
// staged into shared memory by 1-D bulk TMA copies (cp.async.bulk + mbarrier complete_tx),
// double-buffered so the copy of chunk c+1 overlaps the blend of chunk c.  All threads read the
// same record at the same time (shared-memory broadcast); warps vote to leave early.
#include "raster_common.cuh"

namespace dwg {
namespace raster {

int launch_pre(const DwgRasterCamera& cam, const DwgRasterCamera* cam_dev, int64_t N, const float* means3D, const float* opacities,
               const float* scales, const float* rots, GeomView g, BinView b, int T, int64_t P_cap,
               int32_t* radii, int32_t* status, cudaStream_t st);
int launch_sort(int T, BinView b, GeomView g, const float* colors, const int32_t* status, int64_t P_cap,
                int write_keys, cudaStream_t st);

constexpr int FWD_BATCH = 4;      // touched instances whose alphas are evaluated together (ILP)

__global__ void __launch_bounds__(TILE_PIX)
render_fwd_kernel(int H, int W, int gx, const uint2* __restrict__ ranges, const Rec* __restrict__ recs,
                  float bg0, float bg1, float bg2, const float* __restrict__ bg_dev,
                  const float* __restrict__ bg_image, float* __restrict__ out_fg,
                  float* __restrict__ out_color, float* __restrict__ out_depth, float* __restrict__ out_alpha,
                  float* __restrict__ final_T, uint32_t* __restrict__ n_contrib) {
    __shared__ __align__(128) Rec s_rec[2][CHUNK];
    __shared__ __align__(8) uint64_t s_bar[2];
    if (bg_dev) { bg0 = bg_dev[0]; bg1 = bg_dev[1]; bg2 = bg_dev[2]; }
    const int tile = blockIdx.y * gx + blockIdx.x;
    const int px = blockIdx.x * TILE + (threadIdx.x & (TILE - 1));
    const int py = blockIdx.y * TILE + (threadIdx.x >> 4);
    const bool inside = px < W && py < H;
    const float pxf = (float)px, pyf = (float)py;
    const uint2 rg = ranges[tile];
    const int n = (int)(rg.y - rg.x);
    const int rounds = (n + CHUNK - 1) / CHUNK;
    const Rec* src = recs + rg.x;
    if (threadIdx.x == 0) {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (threadIdx.x == 0 && rounds > 0) {
        const uint32_t bytes = (uint32_t)(min(CHUNK, n) * sizeof(Rec));
        mbar_expect_tx(&s_bar[0], bytes);
        tma_bulk_g2s(&s_rec[0][0], src, bytes, &s_bar[0]);
    }
    bool done = !inside;
    float T = 1.0f, C0 = 0.f, C1 = 0.f, C2 = 0.f, D = 0.f, A = 0.f;
    uint32_t last = 0;
    const int lane = threadIdx.x & 31;
    // this warp's pixel strip (16 x 2)
    const float sx0 = (float)(blockIdx.x * TILE), sx1 = sx0 + (float)(TILE - 1);
    const float sy0 = (float)(blockIdx.y * TILE + ((threadIdx.x >> 5) << 1)), sy1 = sy0 + 1.0f;
    for (int c = 0; c < rounds; c++) {
        const int buf = c & 1;
        // every thread has finished reading buffer buf^1 (chunk c-1): safe to refill it with chunk c+1
        const int n_done = __syncthreads_count(done);
        if (n_done == TILE_PIX) {
            mbar_wait(&s_bar[buf], (uint32_t)((c >> 1) & 1));   // never exit with a bulk copy in flight
            break;
        }
        if (threadIdx.x == 0 && c + 1 < rounds) {
            const int cnt = min(CHUNK, n - (c + 1) * CHUNK);
            const uint32_t bytes = (uint32_t)(cnt * sizeof(Rec));
            mbar_expect_tx(&s_bar[buf ^ 1], bytes);
            tma_bulk_g2s(&s_rec[buf ^ 1][0], src + (size_t)(c + 1) * CHUNK, bytes, &s_bar[buf ^ 1]);
        }
        mbar_wait(&s_bar[buf], (uint32_t)((c >> 1) & 1));
        const int cnt = min(CHUNK, n - c * CHUNK);
        // Warp-uniform loop: the first 16 bytes of a record decide whether ANY pixel of this warp's
        // 16x2 strip can be touched; most instances of a tile are rejected here for most warps.
        const bool warp_live = __any_sync(0xffffffffu, !done);
        if (warp_live) {
            // Warp-cooperative culling: each lane tests ONE instance header against this warp's
            // 16x2 pixel strip; only the touched instances (ballot bits, ascending order) are
            // visited by the whole warp.  32x fewer serial iterations for the typical small splat.
            for (int j0 = 0; j0 < cnt; j0 += 32) {
                const int jl = j0 + lane;
                bool touch = false;
                if (jl < cnt) {
                    const float4 h = *reinterpret_cast<const float4*>(&s_rec[buf][jl]);
                    touch = strip_may_touch(h.x, h.y, __float_as_uint(h.z), sx0, sx1, sy0, sy1);
                }
                unsigned mask = __ballot_sync(0xffffffffu, touch);
                // The touched instances are taken FWD_BATCH at a time: their alphas (position-only maths, the
                // long dependent chain with the exp) are evaluated independently of each other first, then the
                // short transmittance chain is applied in source order -- same operations per instance, same
                // order, hence the same bits; ~3x shorter critical path per instance on heavy tiles.
                while (mask) {
                    int jb[FWD_BATCH];
                    bool ok[FWD_BATCH];
                    float al[FWD_BATCH], cr[FWD_BATCH], cg[FWD_BATCH], cb[FWD_BATCH], cd[FWD_BATCH];
#pragma unroll
                    for (int u = 0; u < FWD_BATCH; u++) {
                        ok[u] = (mask != 0u) && !done;
                        jb[u] = j0 + (mask ? __ffs(mask) - 1 : 0);
                        mask &= mask - 1;
                    }
#pragma unroll
                    for (int u = 0; u < FWD_BATCH; u++) {                  // straight-line: FWD_BATCH independent chains
                        const Rec rc = s_rec[buf][jb[u]];
                        float G, dx, dy;
                        ok[u] &= eval_alpha_nb(rc, pxf, pyf, al[u], G, dx, dy);
                        cr[u] = rc.r; cg[u] = rc.g; cb[u] = rc.b; cd[u] = rc.depth;
                    }
#pragma unroll
                    for (int u = 0; u < FWD_BATCH; u++) {
                        if (ok[u] && !done) {
                            const float test_T = __fmul_rn(T, __fsub_rn(1.0f, al[u]));
                            if (test_T < 0.0001f) {
                                done = true;
                            } else {
                                const float w = __fmul_rn(al[u], T);
                                C0 = __fadd_rn(C0, __fmul_rn(cr[u], w));
                                C1 = __fadd_rn(C1, __fmul_rn(cg[u], w));
                                C2 = __fadd_rn(C2, __fmul_rn(cb[u], w));
                                D = __fadd_rn(D, __fmul_rn(cd[u], w));
                                A = __fadd_rn(A, w);
                                T = test_T;
                                last = (uint32_t)(c * CHUNK + jb[u] + 1);
                            }
                        }
                    }
                }
            }
        }
    }
    if (inside) {
        const size_t pix = (size_t)py * W + px;
        const size_t HW = (size_t)H * W;
        final_T[pix] = T;
        n_contrib[pix] = last;
        float o0 = __fadd_rn(C0, __fmul_rn(T, bg0)), o1 = __fadd_rn(C1, __fmul_rn(T, bg1)), o2 = __fadd_rn(C2, __fmul_rn(T, bg2));
        if (bg_image) {
            // R13 (core/system/scene.py:153-166) fused into the epilogue: image = image_fg + image_bg * (1 - alpha)
            if (out_fg) { out_fg[pix] = o0; out_fg[HW + pix] = o1; out_fg[2 * HW + pix] = o2; }
            const float k = __fsub_rn(1.0f, A);
            o0 = __fadd_rn(o0, __fmul_rn(bg_image[pix], k));
            o1 = __fadd_rn(o1, __fmul_rn(bg_image[HW + pix], k));
            o2 = __fadd_rn(o2, __fmul_rn(bg_image[2 * HW + pix], k));
        }
        out_color[pix] = o0;
        out_color[HW + pix] = o1;
        out_color[2 * HW + pix] = o2;
        out_depth[pix] = D;
        out_alpha[pix] = A;
    }
}

}  // namespace raster
}  // namespace dwg

using namespace dwg;
using namespace dwg::raster;

extern "C" int64_t dwg_raster_geom_bytes(int64_t N) { return (int64_t)GeomView::bytes(N > 0 ? N : 1); }
extern "C" int64_t dwg_raster_bin_bytes(int64_t P_cap, int H, int W) {
    const int T = ((W + TILE - 1) / TILE) * ((H + TILE - 1) / TILE);
    return (int64_t)BinView::bytes(P_cap > 0 ? P_cap : 1, T);
}
extern "C" int64_t dwg_raster_img_bytes(int H, int W) { return (int64_t)ImgView::bytes(H, W); }

extern "C" int dwg_raster_forward(const DwgRasterCamera* cam, int64_t N, const float* means3D,
                                  const float* colors_precomp, const float* opacities, const float* scales,
                                  const float* rotations, float* out_color, float* out_depth, float* out_alpha,
                                  int32_t* radii, void* geom, void* bin, int64_t P_cap, void* img,
                                  int32_t* status, const void* cam_dev, const float* bg_image, float* out_color_fg, void* stream) {
    DWG_REQUIRE(cam && out_color && out_depth && out_alpha && geom && bin && img && status, "null pointer");
    DWG_REQUIRE(N == 0 || (means3D && colors_precomp && opacities && scales && rotations && radii), "null input");
    DWG_REQUIRE(cam->image_height > 0 && cam->image_width > 0, "bad image size");
    DWG_REQUIRE(P_cap > 0 && P_cap < (1ll << 31), "P_cap must be in (0, 2^31)");
    const int H = cam->image_height, W = cam->image_width;
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
    const int T = gx * gy;
    DWG_REQUIRE(T <= 65535 * 16, "image too large");
    cudaStream_t st = (cudaStream_t)stream;
    GeomView g(geom, N > 0 ? N : 1);
    BinView b(bin, P_cap, T);
    ImgView im(img, H, W);
    const DwgRasterCamera* cd = reinterpret_cast<const DwgRasterCamera*>(cam_dev);
    int rc = launch_pre(*cam, cd, N, means3D, opacities, scales, rotations, g, b, T, P_cap, radii, status, st);
    if (rc != DWG_OK) return rc;
    rc = launch_sort(T, b, g, colors_precomp, status, P_cap, 1, st);
    if (rc != DWG_OK) return rc;
    render_fwd_kernel<<<dim3(gx, gy), TILE_PIX, 0, st>>>(H, W, gx, b.ranges, b.recs, cam->bg[0], cam->bg[1], cam->bg[2], cd ? cd->bg : nullptr,
                                                        bg_image, out_color_fg, out_color, out_depth, out_alpha, im.final_T, im.n_contrib);
    return check_launch("dwg_raster_forward");
}

extern "C" void* dwg_raster_view(int which, void* geom, void* bin, void* img, int64_t N, int64_t P_cap, int H, int W) {
    const int T = ((W + TILE - 1) / TILE) * ((H + TILE - 1) / TILE);
    GeomView g(geom, N > 0 ? N : 1);
    BinView b(bin, P_cap > 0 ? P_cap : 1, T);
    ImgView im(img, H, W);
    switch (which) {
        case 0: return g.xy;
        case 1: return g.depth;
        case 2: return g.cov3D;
        case 3: return g.conic_opacity;
        case 4: return g.rect;
        case 5: return g.tiles_touched;
        case 6: return b.ranges;
        case 7: return b.keys_out;
        case 8: return b.vals_out;
        case 9: return im.final_T;
        case 10: return im.n_contrib;
        case 11: return b.recs;
        default: return nullptr;
    }
}
