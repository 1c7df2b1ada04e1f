// Rasteriser backward (R12): per-tile back-to-front replay + per-Gaussian chain rule.
//
// Blend backward: one 64-thread CTA per 8x8 quadrant of a tile (same split as the forward render, see raster_fwd.cu),
// one pixel per thread, instance records streamed in reverse with double-buffered bulk TMA copies.  Each pixel replays
// only the instances it blended (per-lane lists built by a cheap power-test pass); its gradient contributions are
// accumulated per chunk slot in per-warp shared memory with plain read-modify-writes (lanes that meet on one instance
// take turns; the benchmark's splats cover 1-4 pixels), and only ONE global atomic per (instance, component) leaves the CTA.
// Upstream issues one global atomic per (pixel, instance, component).
//
// Preprocess backward: conic -> cov2D -> cov3D -> (scale, quaternion), mean2D / depth -> mean3D.
#include "raster_common.cuh"

namespace dwg {
namespace raster {

constexpr int NG = 10;      // mean2D.xy, conic.xyz, opacity, colour.rgb, depth
constexpr int BWD_BATCH = 2;

__global__ void __launch_bounds__(QPIX)
render_bwd_kernel(int H, int W, int gx, const uint2* __restrict__ ranges, const Rec* __restrict__ recs,
                  float bg0, float bg1, float bg2, const float* __restrict__ bg_dev,
                  const float* __restrict__ final_T, const uint32_t* __restrict__ n_contrib,
                  const float* __restrict__ dL_dcolor, const float* __restrict__ dL_ddepth,
                  const float* __restrict__ dL_dalpha,
                  const float* __restrict__ bg_image, const float* __restrict__ out_alpha, float* __restrict__ g_bg_image,
                  float* __restrict__ g_mean2D /* [N,3] */, float4* __restrict__ g_conic_depth /* [N] */,
                  float* __restrict__ g_opacity, float* __restrict__ g_color /* [N,3] */) {
    __shared__ __align__(128) Rec s_rec[2][CHUNK];
    __shared__ float s_acc[QPIX / 32][CHUNK][NG + 1];          // per warp: no cross-warp races, no atomics
    __shared__ __align__(8) uint64_t s_bar[2];
    __shared__ uint32_t s_max[QPIX / 32];
    __shared__ uint8_t s_list[QPIX / 32][CHUNK];              // per warp: instances of the chunk touching the warp's block
    __shared__ uint8_t s_pix[QPIX / 32][CHUNK * 32];          // per lane ([slot][lane]): instances this pixel blends
    if (bg_dev) { bg0 = bg_dev[0]; bg1 = bg_dev[1]; bg2 = bg_dev[2]; }
    const int tile = (blockIdx.y >> 1) * gx + (blockIdx.x >> 1);
    const int lane = threadIdx.x & 31;
    const int bx0 = (blockIdx.x >> 1) * TILE + (blockIdx.x & 1) * QUAD;
    const int by0 = (blockIdx.y >> 1) * TILE + (blockIdx.y & 1) * QUAD + (threadIdx.x >> 5) * 4;
    const int px = bx0 + (lane & 7);
    const int py = by0 + (lane >> 3);
    const bool inside = px < W && py < H;
    const float pxf = (float)px, pyf = (float)py;
    const uint2 rg = ranges[tile];
    const int n = (int)(rg.y - rg.x);
    const size_t pix = (size_t)py * W + px;
    const size_t HW = (size_t)H * W;
    float dLp0 = 0.f, dLp1 = 0.f, dLp2 = 0.f, dLd = 0.f, dLa = 0.f;
    if (inside) {
        dLp0 = dL_dcolor[pix]; dLp1 = dL_dcolor[HW + pix]; dLp2 = dL_dcolor[2 * HW + pix];
        if (dL_ddepth) dLd = dL_ddepth[pix];
        if (dL_dalpha) dLa = dL_dalpha[pix];
        if (bg_image) {
            // backward of image = image_fg + image_bg * (1 - alpha)  (scene.py:153-166): the composite only adds a term to
            // dL/dalpha and yields dL/dimage_bg = dL/dimage * (1 - alpha); dL/dimage_fg = dL/dimage unchanged
            dLa -= bg_image[pix] * dLp0 + bg_image[HW + pix] * dLp1 + bg_image[2 * HW + pix] * dLp2;
            if (g_bg_image) {
                const float k = 1.0f - out_alpha[pix];
                g_bg_image[pix] = dLp0 * k; g_bg_image[HW + pix] = dLp1 * k; g_bg_image[2 * HW + pix] = dLp2 * k;
            }
        }
    }
    if (n <= 0) return;
    const Rec* src = recs + rg.x;
    const uint32_t last_contributor = inside ? n_contrib[pix] : 0u;
    const float T_final = inside ? final_T[pix] : 0.f;
    float T = T_final;
    // tile-wide max of n_contrib: chunks beyond it are never touched
    uint32_t mx = last_contributor;
    for (int off = 16; off > 0; off >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, off));
    if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = mx;
    for (int i = threadIdx.x; i < (QPIX / 32) * CHUNK * (NG + 1); i += QPIX) (&s_acc[0][0][0])[i] = 0.f;
    if (threadIdx.x == 0) {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        mbar_fence_init();
    }
    __syncthreads();
    mx = 0;
#pragma unroll
    for (int w = 0; w < QPIX / 32; w++) mx = max(mx, s_max[w]);
    if (mx == 0) return;
    const int rounds = ((int)mx + CHUNK - 1) / CHUNK;        // chunks [0, rounds) hold contributors
    auto chunk_cnt = [&](int c) { return min(CHUNK, n - c * CHUNK); };
    if (threadIdx.x == 0) {
        const int c = rounds - 1;
        const uint32_t bytes = (uint32_t)(chunk_cnt(c) * sizeof(Rec));
        mbar_expect_tx(&s_bar[0], bytes);
        tma_bulk_g2s(&s_rec[0][0], src + (size_t)c * CHUNK, bytes, &s_bar[0]);
    }
    float accum_r = 0.f, accum_g = 0.f, accum_b = 0.f, accum_d = 0.f, accum_a = 0.f;
    float last_alpha = 0.f, last_r = 0.f, last_g = 0.f, last_b = 0.f, last_depth = 0.f;
    const float bg_dot = bg0 * dLp0 + bg1 * dLp1 + bg2 * dLp2;
    const float ddelx_dx = 0.5f * W, ddely_dy = 0.5f * H;
    const float sx0 = (float)bx0, sx1 = sx0 + 7.0f;
    const float sy0 = (float)by0, sy1 = sy0 + 3.0f;
    uint8_t* my_list = s_list[threadIdx.x >> 5];
    uint8_t* my_pix = s_pix[threadIdx.x >> 5];
    float* my_acc = &s_acc[threadIdx.x >> 5][0][0];
    for (int it = 0; it < rounds; it++) {
        const int c = rounds - 1 - it;
        const int buf = it & 1;
        if (threadIdx.x == 0 && it + 1 < rounds) {
            // buffer buf^1 was released by the __syncthreads at the end of iteration it-1
            const int cn = c - 1;
            const uint32_t bytes = (uint32_t)(chunk_cnt(cn) * sizeof(Rec));
            mbar_expect_tx(&s_bar[buf ^ 1], bytes);
            tma_bulk_g2s(&s_rec[buf ^ 1][0], src + (size_t)cn * CHUNK, bytes, &s_bar[buf ^ 1]);
        }
        mbar_wait(&s_bar[buf], (uint32_t)((it >> 1) & 1));
        const int cnt = chunk_cnt(c);
        // Three passes per chunk (same structure as the forward render, raster_fwd.cu): A. cull against the block's rectangle
        // -> ordered per-warp index list; B. power test + "index < this pixel's last contributor" at every pixel, each lane
        // appends the instances IT blends to its own list; C. every lane replays its own list back to front, BWD_BATCH at a
        // time (the alphas -- the long, position-only chain -- are evaluated together, the short T / accumulator
        // recurrences run in order) and adds its ten partial sums to the warp's shared-memory accumulators.  The trip
        // count of C is the longest per-pixel list of the block, not the number of instances touching the block.
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
        __syncwarp();
        int mine = 0;
        const int lc_rel = (int)min(last_contributor, (uint32_t)0x7fffffff) - c * CHUNK;      // slots < lc_rel contribute to this pixel
        for (int i0 = 0; i0 < m; i0 += 4) {
            bool keep[4];
            int jj[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                jj[u] = my_list[min(i0 + u, m - 1)];
                const float4 h0 = *reinterpret_cast<const float4*>(&s_rec[buf][jj[u]]);
                const float4 h1 = *(reinterpret_cast<const float4*>(&s_rec[buf][jj[u]]) + 1);     // cx, cy, cz, op
                const float dx = __fsub_rn(h0.x, pxf), dy = __fsub_rn(h0.y, pyf);
                const float a = __fmul_rn(__fmul_rn(h1.x, dx), dx);
                const float b = __fmul_rn(__fmul_rn(h1.z, dy), dy);
                const float cc = __fmul_rn(__fmul_rn(h1.y, dx), dy);
                const float power = __fsub_rn(__fmul_rn(-0.5f, __fadd_rn(a, b)), cc);
                keep[u] = (i0 + u < m) & (jj[u] < lc_rel) & !((power > 0.0f) | ((power < -5.6f) & (h1.w <= 1.0f)));
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                if (keep[u]) { my_pix[mine * 32 + lane] = (uint8_t)jj[u]; mine++; }
            }
        }
        const int longest = __reduce_max_sync(0xffffffffu, mine);
        for (int it = 0; it < longest; it += BWD_BATCH) {
            int jb[BWD_BATCH];
            bool ok[BWD_BATCH];
            float al[BWD_BATCH], Gv[BWD_BATCH], dxv[BWD_BATCH], dyv[BWD_BATCH];
            Rec rcs[BWD_BATCH];
#pragma unroll
            for (int u = 0; u < BWD_BATCH; u++) {
                const int pos = mine - 1 - it - u;                 // back to front
                jb[u] = pos >= 0 ? (int)my_pix[pos * 32 + lane] : 0;
                rcs[u] = s_rec[buf][jb[u]];
                ok[u] = eval_alpha_nb(rcs[u], pxf, pyf, al[u], Gv[u], dxv[u], dyv[u]) & (pos >= 0);
            }
#pragma unroll
            for (int u = 0; u < BWD_BATCH; u++) {
                float v[NG];
                if (ok[u]) {
                    const Rec& rc = rcs[u];
                    const float alpha = al[u], G = Gv[u], dx = dxv[u], dy = dyv[u];
                    const float inv = 1.f / (1.f - alpha);
                    T = T * inv;
                    const float dchannel_dcolor = alpha * T;
                    float dL_dalpha_ = 0.f;
                    accum_r = last_alpha * last_r + (1.f - last_alpha) * accum_r; last_r = rc.r;
                    dL_dalpha_ += (rc.r - accum_r) * dLp0;
                    accum_g = last_alpha * last_g + (1.f - last_alpha) * accum_g; last_g = rc.g;
                    dL_dalpha_ += (rc.g - accum_g) * dLp1;
                    accum_b = last_alpha * last_b + (1.f - last_alpha) * accum_b; last_b = rc.b;
                    dL_dalpha_ += (rc.b - accum_b) * dLp2;
                    accum_d = last_alpha * last_depth + (1.f - last_alpha) * accum_d; last_depth = rc.depth;
                    dL_dalpha_ += (rc.depth - accum_d) * dLd;
                    accum_a = last_alpha + (1.f - last_alpha) * accum_a;
                    dL_dalpha_ += (1.f - accum_a) * dLa;
                    dL_dalpha_ *= T;
                    last_alpha = alpha;
                    dL_dalpha_ += (-T_final * inv) * bg_dot;
                    const float dL_dG = rc.op * dL_dalpha_;
                    const float gdx = G * dx, gdy = G * dy;
                    const float dG_ddelx = -gdx * rc.cx - gdy * rc.cy;
                    const float dG_ddely = -gdy * rc.cz - gdx * rc.cy;
                    v[0] = dL_dG * dG_ddelx * ddelx_dx;
                    v[1] = dL_dG * dG_ddely * ddely_dy;
                    v[2] = -0.5f * gdx * dx * dL_dG;
                    v[3] = -0.5f * gdx * dy * dL_dG;
                    v[4] = -0.5f * gdy * dy * dL_dG;
                    v[5] = G * dL_dalpha_;
                    v[6] = dchannel_dcolor * dLp0; v[7] = dchannel_dcolor * dLp1; v[8] = dchannel_dcolor * dLp2;
                    v[9] = dchannel_dcolor * dLd;
                }
                // Accumulate into this WARP's per-slot sums with plain read-modify-writes (shared-memory float atomics are
                // CAS loops): lanes that hold the same instance in this step take turns, one per round -- splats cover
                // 1-4 pixels, so one round is the common case.
                const int key = ok[u] ? jb[u] : (CHUNK + lane);
                const unsigned peers = __match_any_sync(0xffffffffu, key);
                const int rank = __popc(peers & ((1u << lane) - 1u));
                const int nround = __reduce_max_sync(0xffffffffu, __popc(peers));
                for (int r = 0; r < nround; r++) {
                    if (ok[u] && rank == r) {
                        float* acc = my_acc + jb[u] * (NG + 1);
#pragma unroll
                        for (int q = 0; q < NG; q++) acc[q] += v[q];
                    }
                    __syncwarp();
                }
            }
        }
        __syncthreads();            // chunk finished by every warp: flush, and release buffer `buf`
        for (int j = threadIdx.x; j < cnt; j += QPIX) {
            const uint32_t gid = s_rec[buf][j].idx;
            float a[NG];
            bool any = false;
#pragma unroll
            for (int q = 0; q < NG; q++) {
                a[q] = s_acc[0][j][q] + s_acc[1][j][q];
                s_acc[0][j][q] = 0.f; s_acc[1][j][q] = 0.f;
                any |= (a[q] != 0.f);
            }
            if (any) {
                atomicAdd(&g_mean2D[3 * (size_t)gid], a[0]);
                atomicAdd(&g_mean2D[3 * (size_t)gid + 1], a[1]);
                float* cd = reinterpret_cast<float*>(g_conic_depth + gid);
                atomicAdd(cd, a[2]); atomicAdd(cd + 1, a[3]); atomicAdd(cd + 2, a[4]); atomicAdd(cd + 3, a[9]);
                atomicAdd(&g_opacity[gid], a[5]);
                atomicAdd(&g_color[3 * (size_t)gid], a[6]);
                atomicAdd(&g_color[3 * (size_t)gid + 1], a[7]);
                atomicAdd(&g_color[3 * (size_t)gid + 2], a[8]);
            }
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256)
preprocess_bwd_kernel(int64_t N, const float* __restrict__ means3D, const float* __restrict__ scales,
                      const float* __restrict__ rots, const DwgRasterCamera cam_val, const DwgRasterCamera* __restrict__ cam_dev, GeomView g,
                      const float* __restrict__ g_mean2D /* [N,3] */, const float4* __restrict__ g_conic_depth,
                      float* __restrict__ g_means3D, float* __restrict__ g_scales, float* __restrict__ g_rots) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    float gm[3] = {0.f, 0.f, 0.f}, gs[3] = {0.f, 0.f, 0.f}, gr[4] = {0.f, 0.f, 0.f, 0.f};
    if (g.tiles_touched[i] > 0) {
        const DwgRasterCamera cam = cam_dev ? *cam_dev : cam_val;
        const int H = cam.image_height, W = cam.image_width;
        const float fx = W / (2.0f * cam.tanfovx), fy = H / (2.0f * cam.tanfovy);
        const float* view = cam.viewmatrix;
        const float* proj = cam.projmatrix;
        const float p[3] = {means3D[3 * i], means3D[3 * i + 1], means3D[3 * i + 2]};
        float c6[6];
#pragma unroll
        for (int k = 0; k < 6; k++) c6[k] = g.cov3D[6 * i + k];
        float t[3];
        t[0] = view[0] * p[0] + view[4] * p[1] + view[8] * p[2] + view[12];
        t[1] = view[1] * p[0] + view[5] * p[1] + view[9] * p[2] + view[13];
        t[2] = view[2] * p[0] + view[6] * p[1] + view[10] * p[2] + view[14];
        const float limx = 1.3f * cam.tanfovx, limy = 1.3f * cam.tanfovy;
        const float txtz = t[0] / t[2], tytz = t[1] / t[2];
        t[0] = fminf(limx, fmaxf(-limx, txtz)) * t[2];
        t[1] = fminf(limy, fmaxf(-limy, tytz)) * t[2];
        const float x_grad_mul = (txtz < -limx || txtz > limx) ? 0.f : 1.f;
        const float y_grad_mul = (tytz < -limy || tytz > limy) ? 0.f : 1.f;
        const float J00 = fx / t[2], J02 = -(fx * t[0]) / (t[2] * t[2]);
        const float J11 = fy / t[2], J12 = -(fy * t[1]) / (t[2] * t[2]);
        float T[2][3];
#pragma unroll
        for (int c = 0; c < 3; c++) {
            T[0][c] = J00 * view[c * 4 + 0] + J02 * view[c * 4 + 2];
            T[1][c] = J11 * view[c * 4 + 1] + J12 * view[c * 4 + 2];
        }
        const float S[3][3] = {{c6[0], c6[1], c6[2]}, {c6[1], c6[3], c6[4]}, {c6[2], c6[4], c6[5]}};
        float TS[2][3];
#pragma unroll
        for (int a = 0; a < 2; a++)
#pragma unroll
            for (int b = 0; b < 3; b++) TS[a][b] = T[a][0] * S[0][b] + T[a][1] * S[1][b] + T[a][2] * S[2][b];
        const float a = TS[0][0] * T[0][0] + TS[0][1] * T[0][1] + TS[0][2] * T[0][2] + 0.3f;
        const float b = TS[0][0] * T[1][0] + TS[0][1] * T[1][1] + TS[0][2] * T[1][2];
        const float c = TS[1][0] * T[1][0] + TS[1][1] * T[1][1] + TS[1][2] * T[1][2] + 0.3f;
        const float denom = a * c - b * b;
        const float denom2inv = 1.0f / ((denom * denom) + 0.0000001f);
        const float4 gcd = g_conic_depth[i];
        float dL_da = 0.f, dL_db = 0.f, dL_dc = 0.f;
        float dcov[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (denom2inv != 0.f) {
            dL_da = denom2inv * (-c * c * gcd.x + 2 * b * c * gcd.y + (denom - a * c) * gcd.z);
            dL_dc = denom2inv * (-a * a * gcd.z + 2 * a * b * gcd.y + (denom - a * c) * gcd.x);
            dL_db = denom2inv * 2 * (b * c * gcd.x - (denom + 2 * b * b) * gcd.y + a * b * gcd.z);
            dcov[0] = T[0][0] * T[0][0] * dL_da + T[0][0] * T[1][0] * dL_db + T[1][0] * T[1][0] * dL_dc;
            dcov[3] = T[0][1] * T[0][1] * dL_da + T[0][1] * T[1][1] * dL_db + T[1][1] * T[1][1] * dL_dc;
            dcov[5] = T[0][2] * T[0][2] * dL_da + T[0][2] * T[1][2] * dL_db + T[1][2] * T[1][2] * dL_dc;
            dcov[1] = 2 * T[0][0] * T[0][1] * dL_da + (T[0][0] * T[1][1] + T[0][1] * T[1][0]) * dL_db + 2 * T[1][0] * T[1][1] * dL_dc;
            dcov[2] = 2 * T[0][0] * T[0][2] * dL_da + (T[0][0] * T[1][2] + T[0][2] * T[1][0]) * dL_db + 2 * T[1][0] * T[1][2] * dL_dc;
            dcov[4] = 2 * T[0][2] * T[0][1] * dL_da + (T[0][1] * T[1][2] + T[0][2] * T[1][1]) * dL_db + 2 * T[1][1] * T[1][2] * dL_dc;
        }
        float dT[2][3];
#pragma unroll
        for (int j = 0; j < 3; j++) {
            const float st0 = T[0][0] * S[j][0] + T[0][1] * S[j][1] + T[0][2] * S[j][2];
            const float st1 = T[1][0] * S[j][0] + T[1][1] * S[j][1] + T[1][2] * S[j][2];
            dT[0][j] = 2.f * st0 * dL_da + st1 * dL_db;
            dT[1][j] = 2.f * st1 * dL_dc + st0 * dL_db;
        }
        float dJ00 = 0.f, dJ02 = 0.f, dJ11 = 0.f, dJ12 = 0.f;
#pragma unroll
        for (int cc = 0; cc < 3; cc++) {
            dJ00 += view[cc * 4 + 0] * dT[0][cc];
            dJ02 += view[cc * 4 + 2] * dT[0][cc];
            dJ11 += view[cc * 4 + 1] * dT[1][cc];
            dJ12 += view[cc * 4 + 2] * dT[1][cc];
        }
        const float tz = 1.f / t[2], tz2 = tz * tz, tz3 = tz2 * tz;
        const float dtx = x_grad_mul * -fx * tz2 * dJ02;
        const float dty = y_grad_mul * -fy * tz2 * dJ12;
        const float dtz = -fx * tz2 * dJ00 - fy * tz2 * dJ11 + (2 * fx * t[0]) * tz3 * dJ02 + (2 * fy * t[1]) * tz3 * dJ12;
#pragma unroll
        for (int k = 0; k < 3; k++) gm[k] = view[k * 4 + 0] * dtx + view[k * 4 + 1] * dty + view[k * 4 + 2] * dtz;
        // mean2D -> mean3D (perspective divide)
        const float mh3 = proj[3] * p[0] + proj[7] * p[1] + proj[11] * p[2] + proj[15];
        const float m_w = 1.0f / (mh3 + 0.0000001f);
        const float mul1 = (proj[0] * p[0] + proj[4] * p[1] + proj[8] * p[2] + proj[12]) * m_w * m_w;
        const float mul2 = (proj[1] * p[0] + proj[5] * p[1] + proj[9] * p[2] + proj[13]) * m_w * m_w;
        const float d2x = g_mean2D[3 * i], d2y = g_mean2D[3 * i + 1];
        gm[0] += (proj[0] * m_w - proj[3] * mul1) * d2x + (proj[1] * m_w - proj[3] * mul2) * d2y;
        gm[1] += (proj[4] * m_w - proj[7] * mul1) * d2x + (proj[5] * m_w - proj[7] * mul2) * d2y;
        gm[2] += (proj[8] * m_w - proj[11] * mul1) * d2x + (proj[9] * m_w - proj[11] * mul2) * d2y;
        // depth -> mean3D
        gm[0] += view[2] * gcd.w; gm[1] += view[6] * gcd.w; gm[2] += view[10] * gcd.w;
        // cov3D -> scale, rotation
        const float qr = rots[4 * i], qx = rots[4 * i + 1], qy = rots[4 * i + 2], qz = rots[4 * i + 3];
        const float R[3][3] = {
            {1.f - 2.f * (qy * qy + qz * qz), 2.f * (qx * qy - qr * qz), 2.f * (qx * qz + qr * qy)},
            {2.f * (qx * qy + qr * qz), 1.f - 2.f * (qx * qx + qz * qz), 2.f * (qy * qz - qr * qx)},
            {2.f * (qx * qz - qr * qy), 2.f * (qy * qz + qr * qx), 1.f - 2.f * (qx * qx + qy * qy)}};
        const float s[3] = {cam.scale_modifier * scales[3 * i], cam.scale_modifier * scales[3 * i + 1], cam.scale_modifier * scales[3 * i + 2]};
        const float dS[3][3] = {{dcov[0], 0.5f * dcov[1], 0.5f * dcov[2]},
                                {0.5f * dcov[1], dcov[3], 0.5f * dcov[4]},
                                {0.5f * dcov[2], 0.5f * dcov[4], dcov[5]}};
        float M[3][3], dM[3][3], dR[3][3];
#pragma unroll
        for (int a_ = 0; a_ < 3; a_++)
#pragma unroll
            for (int k = 0; k < 3; k++) M[a_][k] = R[a_][k] * s[k];
#pragma unroll
        for (int a_ = 0; a_ < 3; a_++)
#pragma unroll
            for (int k = 0; k < 3; k++) dM[a_][k] = 2.f * (dS[a_][0] * M[0][k] + dS[a_][1] * M[1][k] + dS[a_][2] * M[2][k]);
#pragma unroll
        for (int k = 0; k < 3; k++) {
            gs[k] = cam.scale_modifier * (R[0][k] * dM[0][k] + R[1][k] * dM[1][k] + R[2][k] * dM[2][k]);
#pragma unroll
            for (int a_ = 0; a_ < 3; a_++) dR[a_][k] = dM[a_][k] * s[k];
        }
        gr[0] = 2 * qz * (dR[1][0] - dR[0][1]) + 2 * qy * (dR[0][2] - dR[2][0]) + 2 * qx * (dR[2][1] - dR[1][2]);
        gr[1] = 2 * qy * (dR[0][1] + dR[1][0]) + 2 * qz * (dR[0][2] + dR[2][0]) + 2 * qr * (dR[2][1] - dR[1][2]) - 4 * qx * (dR[2][2] + dR[1][1]);
        gr[2] = 2 * qx * (dR[0][1] + dR[1][0]) + 2 * qr * (dR[0][2] - dR[2][0]) + 2 * qz * (dR[2][1] + dR[1][2]) - 4 * qy * (dR[2][2] + dR[0][0]);
        gr[3] = 2 * qr * (dR[1][0] - dR[0][1]) + 2 * qx * (dR[0][2] + dR[2][0]) + 2 * qy * (dR[2][1] + dR[1][2]) - 4 * qz * (dR[1][1] + dR[0][0]);
    }
#pragma unroll
    for (int k = 0; k < 3; k++) { g_means3D[3 * i + k] = gm[k]; g_scales[3 * i + k] = gs[k]; }
#pragma unroll
    for (int k = 0; k < 4; k++) g_rots[4 * i + k] = gr[k];
}

}  // namespace raster
}  // namespace dwg

using namespace dwg;
using namespace dwg::raster;

extern "C" int64_t dwg_raster_bwd_scratch_bytes(int64_t N) { return (int64_t)align256(sizeof(float4) * (size_t)(N > 0 ? N : 1)); }

extern "C" int dwg_raster_backward(const DwgRasterCamera* cam, int64_t N, const float* means3D,
                                   const float* colors_precomp, const float* opacities, const float* scales,
                                   const float* rotations, const void* geom, const void* bin, int64_t P_cap,
                                   const void* img, const float* dL_dcolor, const float* dL_ddepth,
                                   const float* dL_dalpha, float* g_means3D, float* g_means2D, float* g_colors,
                                   float* g_opacities, float* g_scales, float* g_rotations, void* scratch,
                                   const void* cam_dev, const float* bg_image, const float* out_alpha, float* g_bg_image, void* stream) {
    (void)colors_precomp; (void)opacities;
    DWG_REQUIRE(cam && geom && bin && img && dL_dcolor && scratch, "null pointer");
    DWG_REQUIRE(!g_bg_image || (bg_image && out_alpha), "g_bg_image needs bg_image and the forward's out_alpha");
    DWG_REQUIRE(N == 0 || (means3D && scales && rotations && g_means3D && g_means2D && g_colors && g_opacities && g_scales && g_rotations),
                "null gradient buffer");
    if (N == 0) return DWG_OK;
    const int H = cam->image_height, W = cam->image_width;
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
    const int T = gx * gy;
    cudaStream_t st = (cudaStream_t)stream;
    GeomView g(const_cast<void*>(geom), N);
    BinView b(const_cast<void*>(bin), P_cap, T);
    ImgView im(const_cast<void*>(img), H, W);
    float4* gcd = reinterpret_cast<float4*>(scratch);
    const DwgRasterCamera* cd = reinterpret_cast<const DwgRasterCamera*>(cam_dev);
    cudaMemsetAsync(g_means2D, 0, sizeof(float) * 3 * N, st);
    cudaMemsetAsync(g_colors, 0, sizeof(float) * 3 * N, st);
    cudaMemsetAsync(g_opacities, 0, sizeof(float) * N, st);
    cudaMemsetAsync(gcd, 0, sizeof(float4) * N, st);
    render_bwd_kernel<<<dim3(2 * gx, 2 * gy), QPIX, 0, st>>>(H, W, gx, b.ranges, b.recs, cam->bg[0], cam->bg[1], cam->bg[2], cd ? cd->bg : nullptr,
                                                        im.final_T, im.n_contrib, dL_dcolor, dL_ddepth, dL_dalpha,
                                                        bg_image, out_alpha, g_bg_image, g_means2D, gcd, g_opacities, g_colors);
    preprocess_bwd_kernel<<<(unsigned)ceil_div(N, 256), 256, 0, st>>>(N, means3D, scales, rotations, *cam, cd, g, g_means2D, gcd,
                                                                     g_means3D, g_scales, g_rotations);
    return check_launch("dwg_raster_backward");
}
