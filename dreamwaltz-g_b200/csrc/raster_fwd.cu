// Rasteriser stage 3 (R12/R13): per-tile front-to-back alpha blending, and the C-ABI forward.
//
// One 64-thread CTA per 8x8 QUADRANT of a 16x16 tile (four CTAs walk the same instance list), one pixel per thread,
// a warp = an 8x4 pixel block (the per-pixel transmittance chain is evaluated sequentially in source order so that
// final_T and n_contrib are bit-exact with the oracle).  Why quadrants: every CTA of the launch is resident at once, so
// the kernel lasts as long as its heaviest tile (a few silhouette tiles hold 5-9k instances, tools/raster_probe.py);
// with one CTA per tile that tile's eight warps shared one SM's issue slots and a chunk barrier, now they are spread
// over four SMs.  The tile's depth-sorted 48-byte instance records are contiguous in HBM; they are staged into shared
// memory by 1-D bulk TMA copies (cp.async.bulk + mbarrier complete_tx), double-buffered so the copy of chunk c+1
// overlaps the blend of chunk c.  All threads read the same record at the same time (shared-memory broadcast).
#include "raster_common.cuh"

namespace dwg {
namespace raster {

int launch_pre(const DwgRasterCamera& cam, const DwgRasterCamera* cam_dev, int64_t N, const float* means3D, const float* opacities,
               const float* scales, const float* rots, GeomView g, BinView b, int T, int64_t P_cap,
               int32_t* radii, int32_t* status, cudaStream_t st);
int launch_sort(int T, BinView b, GeomView g, const float* colors, const int32_t* status, int64_t P_cap,
                int write_keys, cudaStream_t st);

constexpr int FWD_BATCH = 4;      // touched instances whose alphas are evaluated together (ILP)

__global__ void __launch_bounds__(QPIX)
render_fwd_kernel(int H, int W, int gx, const uint2* __restrict__ ranges, const Rec* __restrict__ recs,
                  float bg0, float bg1, float bg2, const float* __restrict__ bg_dev,
                  const float* __restrict__ bg_image, float* __restrict__ out_fg,
                  float* __restrict__ out_color, float* __restrict__ out_depth, float* __restrict__ out_alpha,
                  float* __restrict__ final_T, uint32_t* __restrict__ n_contrib, unsigned long long* __restrict__ probe) {
    __shared__ __align__(128) Rec s_rec[2][CHUNK];
    __shared__ unsigned int s_probe[2];
    __shared__ uint8_t s_list[QPIX / 32][CHUNK];              // per warp: instances of the chunk touching the warp's block
    __shared__ uint8_t s_pix[QPIX / 32][CHUNK * 32];          // per lane ([slot][lane]): instances this pixel keeps
    unsigned long long t_probe0 = 0;
    unsigned int n_touch = 0, n_eval = 0;
    if (probe) {
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_probe0));
        if (threadIdx.x < 2) s_probe[threadIdx.x] = 0;
    }
    __shared__ __align__(8) uint64_t s_bar[2];
    if (bg_dev) { bg0 = bg_dev[0]; bg1 = bg_dev[1]; bg2 = bg_dev[2]; }
    const int tile = (blockIdx.y >> 1) * gx + (blockIdx.x >> 1);
    const int lane = threadIdx.x & 31;
    // this warp's 8 x 4 pixel block inside the quadrant
    const int bx0 = (blockIdx.x >> 1) * TILE + (blockIdx.x & 1) * QUAD;
    const int by0 = (blockIdx.y >> 1) * TILE + (blockIdx.y & 1) * QUAD + (threadIdx.x >> 5) * 4;
    const int px = bx0 + (lane & 7);
    const int py = by0 + (lane >> 3);
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
    const float sx0 = (float)bx0, sx1 = sx0 + 7.0f;
    const float sy0 = (float)by0, sy1 = sy0 + 3.0f;
    uint8_t* my_list = s_list[threadIdx.x >> 5];
    uint8_t* my_pix = s_pix[threadIdx.x >> 5];
    for (int c = 0; c < rounds; c++) {
        const int buf = c & 1;
        // every thread has finished reading buffer buf^1 (chunk c-1): safe to refill it with chunk c+1
        const int n_done = __syncthreads_count(done);
        if (n_done == QPIX) {
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
        if (!__any_sync(0xffffffffu, !done)) continue;          // warp-uniform: every pixel of this block is finished
        // Three passes over the chunk (the blend loop is a latency chain on the heaviest block, tools/raster_probe.py: the
        // benchmark's splats cover 1-4 pixels, so a block is "touched" by thousands of instances of which each PIXEL keeps
        // a few per cent):
        //  A. cull: each lane tests ONE record header (16 B) against the block's rectangle; the indices of the touched
        //     instances are compacted, in order, into a per-warp list;
        //  B. power test (the cheap half of the alpha evaluation) of every listed instance at every pixel, four
        //     independent instances at a time; a lane appends the instances IT keeps to its own list;
        //  C. every lane walks its own list: full alpha evaluation + the transmittance recurrence, in source order.  The
        //     trip count is the longest per-pixel list of the block, not the number of instances touching the block.
        // Per pixel and instance the arithmetic of C is exactly that of the oracle, in the same order: same bits.
        int m = 0;
        for (int j0 = 0; j0 < cnt; j0 += 32) {
            const int jl = j0 + lane;
            bool touch = false;
            if (jl < cnt) {
                const float4 h = *reinterpret_cast<const float4*>(&s_rec[buf][jl]);
                touch = strip_may_touch(h.x, h.y, __float_as_uint(h.z), sx0, sx1, sy0, sy1);
            }
            const unsigned mask = __ballot_sync(0xffffffffu, touch);
            if (touch) my_list[m + __popc(mask & ((1u << lane) - 1u))] = (uint8_t)jl;
            m += __popc(mask);
        }
        if (probe) n_touch += m;
        __syncwarp();
        int mine = 0;                                               // length of this lane's list
        for (int i0 = 0; i0 < m; i0 += FWD_BATCH) {
            bool keep[FWD_BATCH];
            int jj[FWD_BATCH];
#pragma unroll
            for (int u = 0; u < FWD_BATCH; u++) {
                jj[u] = my_list[min(i0 + u, m - 1)];
                const float4 h0 = *reinterpret_cast<const float4*>(&s_rec[buf][jj[u]]);
                const float4 h1 = *(reinterpret_cast<const float4*>(&s_rec[buf][jj[u]]) + 1);     // cx, cy, cz, op
                const float dx = __fsub_rn(h0.x, pxf), dy = __fsub_rn(h0.y, pyf);
                const float a = __fmul_rn(__fmul_rn(h1.x, dx), dx);
                const float b = __fmul_rn(__fmul_rn(h1.z, dy), dy);
                const float cc = __fmul_rn(__fmul_rn(h1.y, dx), dy);
                const float power = __fsub_rn(__fmul_rn(-0.5f, __fadd_rn(a, b)), cc);
                keep[u] = (i0 + u < m) & !done & !((power > 0.0f) | ((power < -5.6f) & (h1.w <= 1.0f)));
            }
#pragma unroll
            for (int u = 0; u < FWD_BATCH; u++) {
                if (keep[u]) { my_pix[mine * 32 + lane] = (uint8_t)jj[u]; mine++; }
            }
        }
        const int longest = __reduce_max_sync(0xffffffffu, mine);
        for (int i0 = 0; i0 < longest; i0 += FWD_BATCH) {
            int jb[FWD_BATCH];
            bool ok[FWD_BATCH];
            float al[FWD_BATCH], cr[FWD_BATCH], cg[FWD_BATCH], cb[FWD_BATCH], cd[FWD_BATCH];
#pragma unroll
            for (int u = 0; u < FWD_BATCH; u++) {                  // straight-line: FWD_BATCH independent chains
                const bool valid = i0 + u < mine;
                jb[u] = valid ? (int)my_pix[(i0 + u) * 32 + lane] : 0;
                const Rec rc = s_rec[buf][jb[u]];
                float G, dx, dy;
                ok[u] = eval_alpha_nb(rc, pxf, pyf, al[u], G, dx, dy) & valid;
                cr[u] = rc.r; cg[u] = rc.g; cb[u] = rc.b; cd[u] = rc.depth;
            }
#pragma unroll
            for (int u = 0; u < FWD_BATCH; u++) {                  // the recurrence, in source order, branch-free
                const bool act = ok[u] & !done;
                const float test_T = __fmul_rn(T, __fsub_rn(1.0f, al[u]));
                const bool stop = act & (test_T < 0.0001f);
                const bool take = act & !stop;
                const float w = __fmul_rn(al[u], T);
                C0 = take ? __fadd_rn(C0, __fmul_rn(cr[u], w)) : C0;
                C1 = take ? __fadd_rn(C1, __fmul_rn(cg[u], w)) : C1;
                C2 = take ? __fadd_rn(C2, __fmul_rn(cb[u], w)) : C2;
                D = take ? __fadd_rn(D, __fmul_rn(cd[u], w)) : D;
                A = take ? __fadd_rn(A, w) : A;
                T = take ? test_T : T;
                last = take ? (uint32_t)(c * CHUNK + jb[u] + 1) : last;
                done |= stop;
                if (probe) n_eval += take ? 1u : 0u;
            }
        }
        __syncwarp();                                               // lists are rewritten by the next chunk
    }
    if (probe) {       // debug timeline (dwg_raster_probe): [start ns, end ns, n, max strip touched, max pixel blended, smid]
        atomicMax(&s_probe[0], n_touch);
        atomicMax(&s_probe[1], n_eval);
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned long long t1;
            unsigned int smid;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            unsigned long long* o = probe + 6 * ((size_t)blockIdx.y * gridDim.x + blockIdx.x);
            o[0] = t_probe0; o[1] = t1; o[2] = (unsigned long long)n; o[3] = s_probe[0]; o[4] = s_probe[1]; o[5] = smid;
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

static unsigned long long* g_raster_probe = nullptr;
/* Debug: device pointer to [4 * tiles][6] u64 that every quadrant CTA of the following forward renders fills with
 * (start ns, end ns, instances, max touched per strip, max blended per pixel, smid); NULL = off.  tools/raster_probe.py */
extern "C" int dwg_raster_probe(void* dev_u64) { g_raster_probe = reinterpret_cast<unsigned long long*>(dev_u64); return DWG_OK; }

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
    render_fwd_kernel<<<dim3(2 * gx, 2 * gy), QPIX, 0, st>>>(H, W, gx, b.ranges, b.recs, cam->bg[0], cam->bg[1], cam->bg[2], cd ? cd->bg : nullptr,
                                                        bg_image, out_color_fg, out_color, out_depth, out_alpha, im.final_T, im.n_contrib, g_raster_probe);
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
