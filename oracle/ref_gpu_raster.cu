/* ref_gpu_raster.cu -- BASELINE STAND-IN, TEST/BENCH INFRASTRUCTURE ONLY (never linked into libdwg_sm100.so).
 *
 * A plain SIMT restatement of the PUBLISHED 3DGS rasteriser algorithm (the un-vendored third-party
 * ashawkey/diff-gaussian-rasterization the reference calls at core/gaussian/gaussian_renderer.py:186-195, SURVEY.md
 * appendix B) in the structure the public implementation uses: thread-per-Gaussian preprocess, inclusive scan of
 * tiles_touched, key duplication, ONE global radix sort of (tile << 32 | depth) keys, tile-range detection, a
 * 16x16-thread CTA per tile that fetches 256 instances at a time into shared memory and blends front to back, and a
 * backward that walks each pixel back to front with per-Gaussian global atomics.  No TMA, no culling, no packing, no
 * warp-level reduction: this is the "reference-equivalent GPU" raster arm of bench.py (--impl reference-gpu), labelled
 * as a stand-in because the real package cannot be installed here (no network).  The scan and the sort are done by the
 * caller with torch.cumsum / torch.sort (CUB radix sort underneath, as upstream).
 *
 * Built by oracle/build_ref.py -> oracle/_ref/libref_raster_simt.so.  Arithmetic follows oracle/oracle_c.c (same
 * formulas, ordinary fp32 with FMA contraction and expf, as upstream).
 */
#include <cuda_runtime.h>
#include <stdint.h>

#define TILE 16
#define BLOCK (TILE * TILE)

struct RefCamera {
    int H, W;
    float tanfovx, tanfovy;
    float view[16];
    float proj[16];
    float bg[3];
    float scale_modifier;
};

__device__ __forceinline__ void xf43(const float* p, const float* m, float* o) {
    o[0] = m[0] * p[0] + m[4] * p[1] + m[8] * p[2] + m[12];
    o[1] = m[1] * p[0] + m[5] * p[1] + m[9] * p[2] + m[13];
    o[2] = m[2] * p[0] + m[6] * p[1] + m[10] * p[2] + m[14];
}
__device__ __forceinline__ void xf44(const float* p, const float* m, float* o) {
    xf43(p, m, o);
    o[3] = m[3] * p[0] + m[7] * p[1] + m[11] * p[2] + m[15];
}

__device__ void cov3d(const float* s3, float mod, const float* q, float* c6) {
    float r = q[0], x = q[1], y = q[2], z = q[3];
    float R[3][3] = {{1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y)},
                     {2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x)},
                     {2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y)}};
    float M[3][3];
    for (int i = 0; i < 3; i++) for (int k = 0; k < 3; k++) M[i][k] = R[i][k] * (mod * s3[k]);
    int t = 0;
    for (int i = 0; i < 3; i++) for (int j = i; j < 3; j++) c6[t++] = M[i][0] * M[j][0] + M[i][1] * M[j][1] + M[i][2] * M[j][2];
}

__device__ void cov2d(const float* tv, float fx, float fy, float tfx, float tfy, const float* c6, const float* view, float* abc, float T[2][3]) {
    float tx = tv[0], ty = tv[1], tz = tv[2];
    float limx = 1.3f * tfx, limy = 1.3f * tfy;
    tx = fminf(limx, fmaxf(-limx, tx / tz)) * tz;
    ty = fminf(limy, fmaxf(-limy, ty / tz)) * tz;
    float J00 = fx / tz, J02 = -(fx * tx) / (tz * tz), J11 = fy / tz, J12 = -(fy * ty) / (tz * tz);
    for (int c = 0; c < 3; c++) {
        T[0][c] = J00 * view[c * 4 + 0] + J02 * view[c * 4 + 2];
        T[1][c] = J11 * view[c * 4 + 1] + J12 * view[c * 4 + 2];
    }
    float S[3][3] = {{c6[0], c6[1], c6[2]}, {c6[1], c6[3], c6[4]}, {c6[2], c6[4], c6[5]}};
    float TS[2][3];
    for (int i = 0; i < 2; i++) for (int j = 0; j < 3; j++) TS[i][j] = T[i][0] * S[0][j] + T[i][1] * S[1][j] + T[i][2] * S[2][j];
    abc[0] = TS[0][0] * T[0][0] + TS[0][1] * T[0][1] + TS[0][2] * T[0][2] + 0.3f;
    abc[1] = TS[0][0] * T[1][0] + TS[0][1] * T[1][1] + TS[0][2] * T[1][2];
    abc[2] = TS[1][0] * T[1][0] + TS[1][1] * T[1][1] + TS[1][2] * T[1][2] + 0.3f;
}

__global__ void k_preprocess(int N, const float* means3D, const float* scales, const float* rots, const float* opac, RefCamera cam,
                             int* radii, float2* xy, float* depth, float* cov3D, float4* conic_op, int4* rect, int* tiles_touched) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    radii[i] = 0; tiles_touched[i] = 0;
    const int gx = (cam.W + TILE - 1) / TILE, gy = (cam.H + TILE - 1) / TILE;
    const float fx = cam.W / (2.0f * cam.tanfovx), fy = cam.H / (2.0f * cam.tanfovy);
    const float* p = means3D + 3 * (size_t)i;
    float pv[3]; xf43(p, cam.view, pv);
    if (pv[2] <= 0.2f) return;
    float ph[4]; xf44(p, cam.proj, ph);
    float pw = 1.0f / (ph[3] + 0.0000001f);
    float c6[6]; cov3d(scales + 3 * (size_t)i, cam.scale_modifier, rots + 4 * (size_t)i, c6);
    for (int k = 0; k < 6; k++) cov3D[6 * (size_t)i + k] = c6[k];
    float abc[3], T[2][3]; cov2d(pv, fx, fy, cam.tanfovx, cam.tanfovy, c6, cam.view, abc, T);
    float det = abc[0] * abc[2] - abc[1] * abc[1];
    if (det == 0.0f) return;
    float di = 1.f / det;
    float mid = 0.5f * (abc[0] + abc[2]);
    float sq = sqrtf(fmaxf(0.1f, mid * mid - det));
    float rad = ceilf(3.f * sqrtf(fmaxf(mid + sq, mid - sq)));
    float px = ((ph[0] * pw + 1.0) * cam.W - 1.0) * 0.5;
    float py = ((ph[1] * pw + 1.0) * cam.H - 1.0) * 0.5;
    int r = (int)rad;
    int x0 = min(gx, max(0, (int)((px - r) / TILE))), y0 = min(gy, max(0, (int)((py - r) / TILE)));
    int x1 = min(gx, max(0, (int)((px + r + TILE - 1) / TILE))), y1 = min(gy, max(0, (int)((py + r + TILE - 1) / TILE)));
    if ((x1 - x0) * (y1 - y0) == 0) return;
    depth[i] = pv[2]; radii[i] = r; xy[i] = make_float2(px, py);
    conic_op[i] = make_float4(abc[2] * di, -abc[1] * di, abc[0] * di, opac[i]);
    rect[i] = make_int4(x0, y0, x1, y1);
    tiles_touched[i] = (x1 - x0) * (y1 - y0);
}

__global__ void k_duplicate(int N, const int* radii, const float* depth, const int4* rect, const int64_t* offsets_incl, int gx,
                            int64_t* keys, int* vals) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N || radii[i] <= 0) return;
    int64_t off = i == 0 ? 0 : offsets_incl[i - 1];
    int4 rc = rect[i];
    for (int y = rc.y; y < rc.w; y++)
        for (int x = rc.x; x < rc.z; x++) {
            keys[off] = ((int64_t)(y * gx + x) << 32) | (uint32_t)__float_as_uint(depth[i]);
            vals[off] = i;
            off++;
        }
}

__global__ void k_ranges(int64_t P, const int64_t* keys, uint2* ranges) {
    int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (j >= P) return;
    uint32_t t = (uint32_t)(keys[j] >> 32);
    if (j == 0) ranges[t].x = 0;
    else {
        uint32_t tp = (uint32_t)(keys[j - 1] >> 32);
        if (t != tp) { ranges[tp].y = (uint32_t)j; ranges[t].x = (uint32_t)j; }
    }
    if (j == P - 1) ranges[t].y = (uint32_t)P;
}

__global__ void __launch_bounds__(BLOCK) k_render(RefCamera cam, const uint2* ranges, const int* vals, const float2* xy, const float4* conic_op,
                                                  const float* colors, const float* depth, float* out_color, float* out_depth, float* out_alpha,
                                                  float* final_T, uint32_t* n_contrib) {
    const int gx = (cam.W + TILE - 1) / TILE;
    const int px = blockIdx.x * TILE + threadIdx.x, py = blockIdx.y * TILE + threadIdx.y;
    const int tid = threadIdx.y * TILE + threadIdx.x;
    const bool inside = px < cam.W && py < cam.H;
    const int pix = py * cam.W + px;
    uint2 range = ranges[blockIdx.y * gx + blockIdx.x];
    int todo = range.y - range.x;
    const int rounds = (todo + BLOCK - 1) / BLOCK;
    __shared__ int s_id[BLOCK];
    __shared__ float2 s_xy[BLOCK];
    __shared__ float4 s_co[BLOCK];
    bool done = !inside;
    float T = 1.f, C[3] = {0, 0, 0}, D = 0, A = 0;
    uint32_t contributor = 0, last = 0;
    for (int i = 0; i < rounds; i++, todo -= BLOCK) {
        if (__syncthreads_count(done) == BLOCK) break;
        int progress = i * BLOCK + tid;
        if (range.x + progress < range.y) {
            int g = vals[range.x + progress];
            s_id[tid] = g; s_xy[tid] = xy[g]; s_co[tid] = conic_op[g];
        }
        __syncthreads();
        for (int j = 0; !done && j < min(BLOCK, todo); j++) {
            contributor++;
            float2 c = s_xy[j]; float4 co = s_co[j];
            float dx = c.x - (float)px, dy = c.y - (float)py;
            float power = -0.5f * (co.x * dx * dx + co.z * dy * dy) - co.y * dx * dy;
            if (power > 0.0f) continue;
            float alpha = fminf(0.99f, co.w * expf(power));
            if (alpha < 1.0f / 255.0f) continue;
            float test_T = T * (1 - alpha);
            if (test_T < 0.0001f) { done = true; continue; }
            int g = s_id[j];
            float w = alpha * T;
            for (int ch = 0; ch < 3; ch++) C[ch] += colors[3 * (size_t)g + ch] * w;
            D += depth[g] * w; A += w;
            T = test_T; last = contributor;
        }
    }
    if (inside) {
        final_T[pix] = T; n_contrib[pix] = last;
        for (int ch = 0; ch < 3; ch++) out_color[(size_t)ch * cam.H * cam.W + pix] = C[ch] + T * cam.bg[ch];
        out_depth[pix] = D; out_alpha[pix] = A;
    }
}

__global__ void __launch_bounds__(BLOCK) k_render_bwd(RefCamera cam, const uint2* ranges, const int* vals, const float2* xy, const float4* conic_op,
                                                      const float* colors, const float* depth, const float* final_T, const uint32_t* n_contrib,
                                                      const float* dL_dpix, const float* dL_ddepth, const float* dL_dalpha_pix,
                                                      float* g_mean2D /*[N,2]*/, float* g_conic /*[N,3]*/, float* g_opac, float* g_color, float* g_depth) {
    const int gx = (cam.W + TILE - 1) / TILE;
    const int px = blockIdx.x * TILE + threadIdx.x, py = blockIdx.y * TILE + threadIdx.y;
    const int tid = threadIdx.y * TILE + threadIdx.x;
    const bool inside = px < cam.W && py < cam.H;
    const int pix = py * cam.W + px;
    const size_t HW = (size_t)cam.H * cam.W;
    uint2 range = ranges[blockIdx.y * gx + blockIdx.x];
    int todo = range.y - range.x;
    const int rounds = (todo + BLOCK - 1) / BLOCK;
    __shared__ int s_id[BLOCK];
    __shared__ float2 s_xy[BLOCK];
    __shared__ float4 s_co[BLOCK];
    __shared__ float s_col[3 * BLOCK];
    __shared__ float s_dep[BLOCK];
    bool done = !inside;
    const float T_final = inside ? final_T[pix] : 0;
    float T = T_final;
    uint32_t contributor = todo;
    const int last_contributor = inside ? n_contrib[pix] : 0;
    float dLp[3] = {0, 0, 0}, dLd = 0, dLa = 0;
    if (inside) {
        for (int ch = 0; ch < 3; ch++) dLp[ch] = dL_dpix[ch * HW + pix];
        dLd = dL_ddepth ? dL_ddepth[pix] : 0.f;
        dLa = dL_dalpha_pix ? dL_dalpha_pix[pix] : 0.f;
    }
    float accum_rec[3] = {0, 0, 0}, accum_d = 0, accum_a = 0, last_alpha = 0, last_color[3] = {0, 0, 0}, last_depth = 0;
    const float bg_dot = cam.bg[0] * dLp[0] + cam.bg[1] * dLp[1] + cam.bg[2] * dLp[2];
    const float ddelx_dx = 0.5f * cam.W, ddely_dy = 0.5f * cam.H;
    for (int i = 0; i < rounds; i++, todo -= BLOCK) {
        __syncthreads();
        const int progress = i * BLOCK + tid;
        if (range.x + progress < range.y) {
            const int g = vals[range.y - progress - 1];           // back to front
            s_id[tid] = g; s_xy[tid] = xy[g]; s_co[tid] = conic_op[g];
            for (int ch = 0; ch < 3; ch++) s_col[ch * BLOCK + tid] = colors[3 * (size_t)g + ch];
            s_dep[tid] = depth[g];
        }
        __syncthreads();
        for (int j = 0; !done && j < min(BLOCK, todo); j++) {
            contributor--;
            if (contributor >= (uint32_t)last_contributor) continue;
            const float2 c = s_xy[j]; const float4 co = s_co[j];
            const float dx = c.x - (float)px, dy = c.y - (float)py;
            const float power = -0.5f * (co.x * dx * dx + co.z * dy * dy) - co.y * dx * dy;
            if (power > 0.0f) continue;
            const float G = expf(power);
            const float alpha = fminf(0.99f, co.w * G);
            if (alpha < 1.0f / 255.0f) continue;
            T = T / (1.f - alpha);
            const float dch = alpha * T;
            float dL_dalpha = 0.f;
            const int g = s_id[j];
            for (int ch = 0; ch < 3; ch++) {
                const float col = s_col[ch * BLOCK + j];
                accum_rec[ch] = last_alpha * last_color[ch] + (1.f - last_alpha) * accum_rec[ch];
                last_color[ch] = col;
                dL_dalpha += (col - accum_rec[ch]) * dLp[ch];
                atomicAdd(&g_color[3 * (size_t)g + ch], dch * dLp[ch]);
            }
            const float dep = s_dep[j];
            accum_d = last_alpha * last_depth + (1.f - last_alpha) * accum_d;
            last_depth = dep;
            dL_dalpha += (dep - accum_d) * dLd;
            atomicAdd(&g_depth[g], dch * dLd);
            accum_a = last_alpha + (1.f - last_alpha) * accum_a;
            dL_dalpha += (1.f - accum_a) * dLa;
            dL_dalpha *= T;
            last_alpha = alpha;
            dL_dalpha += (-T_final / (1.f - alpha)) * bg_dot;
            const float dL_dG = co.w * dL_dalpha;
            const float gdx = G * dx, gdy = G * dy;
            const float dG_ddelx = -gdx * co.x - gdy * co.y, dG_ddely = -gdy * co.z - gdx * co.y;
            atomicAdd(&g_mean2D[2 * (size_t)g + 0], dL_dG * dG_ddelx * ddelx_dx);
            atomicAdd(&g_mean2D[2 * (size_t)g + 1], dL_dG * dG_ddely * ddely_dy);
            atomicAdd(&g_conic[3 * (size_t)g + 0], -0.5f * gdx * dx * dL_dG);
            atomicAdd(&g_conic[3 * (size_t)g + 1], -0.5f * gdx * dy * dL_dG);
            atomicAdd(&g_conic[3 * (size_t)g + 2], -0.5f * gdy * dy * dL_dG);
            atomicAdd(&g_opac[g], G * dL_dalpha);
        }
    }
}

__global__ void k_preprocess_bwd(int N, const float* means3D, const float* scales, const float* rots, RefCamera cam, const int* radii,
                                 const float* cov3D, const float* g_mean2D, const float* g_conic, const float* g_depth,
                                 float* g_means3D, float* g_scales, float* g_rots) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    for (int k = 0; k < 3; k++) { g_means3D[3 * (size_t)i + k] = 0; g_scales[3 * (size_t)i + k] = 0; }
    for (int k = 0; k < 4; k++) g_rots[4 * (size_t)i + k] = 0;
    if (radii[i] <= 0) return;
    const float fx = cam.W / (2.0f * cam.tanfovx), fy = cam.H / (2.0f * cam.tanfovy);
    const float* view = cam.view; const float* proj = cam.proj;
    const float* p = means3D + 3 * (size_t)i;
    const float* c6 = cov3D + 6 * (size_t)i;
    float t[3]; xf43(p, view, t);
    const float limx = 1.3f * cam.tanfovx, limy = 1.3f * cam.tanfovy;
    const float txtz = t[0] / t[2], tytz = t[1] / t[2];
    t[0] = fminf(limx, fmaxf(-limx, txtz)) * t[2];
    t[1] = fminf(limy, fmaxf(-limy, tytz)) * t[2];
    const float xm = (txtz < -limx || txtz > limx) ? 0.f : 1.f, ym = (tytz < -limy || tytz > limy) ? 0.f : 1.f;
    float J00 = fx / t[2], J02 = -(fx * t[0]) / (t[2] * t[2]), J11 = fy / t[2], J12 = -(fy * t[1]) / (t[2] * t[2]);
    float T[2][3];
    for (int c = 0; c < 3; c++) {
        T[0][c] = J00 * view[c * 4 + 0] + J02 * view[c * 4 + 2];
        T[1][c] = J11 * view[c * 4 + 1] + J12 * view[c * 4 + 2];
    }
    float S[3][3] = {{c6[0], c6[1], c6[2]}, {c6[1], c6[3], c6[4]}, {c6[2], c6[4], c6[5]}};
    float TS[2][3];
    for (int r = 0; r < 2; r++) for (int j = 0; j < 3; j++) TS[r][j] = T[r][0] * S[0][j] + T[r][1] * S[1][j] + T[r][2] * S[2][j];
    float a = TS[0][0] * T[0][0] + TS[0][1] * T[0][1] + TS[0][2] * T[0][2] + 0.3f;
    float b = TS[0][0] * T[1][0] + TS[0][1] * T[1][1] + TS[0][2] * T[1][2];
    float c = TS[1][0] * T[1][0] + TS[1][1] * T[1][1] + TS[1][2] * T[1][2] + 0.3f;
    float denom = a * c - b * b;
    float dL_da = 0, dL_db = 0, dL_dc = 0;
    float d2i = 1.0f / ((denom * denom) + 0.0000001f);
    float dcx = g_conic[3 * (size_t)i], dcy = g_conic[3 * (size_t)i + 1], dcz = g_conic[3 * (size_t)i + 2];
    float dcv[6] = {0, 0, 0, 0, 0, 0};
    if (d2i != 0) {
        dL_da = d2i * (-c * c * dcx + 2 * b * c * dcy + (denom - a * c) * dcz);
        dL_dc = d2i * (-a * a * dcz + 2 * a * b * dcy + (denom - a * c) * dcx);
        dL_db = d2i * 2 * (b * c * dcx - (denom + 2 * b * b) * dcy + a * b * dcz);
        dcv[0] = T[0][0] * T[0][0] * dL_da + T[0][0] * T[1][0] * dL_db + T[1][0] * T[1][0] * dL_dc;
        dcv[3] = T[0][1] * T[0][1] * dL_da + T[0][1] * T[1][1] * dL_db + T[1][1] * T[1][1] * dL_dc;
        dcv[5] = T[0][2] * T[0][2] * dL_da + T[0][2] * T[1][2] * dL_db + T[1][2] * T[1][2] * dL_dc;
        dcv[1] = 2 * T[0][0] * T[0][1] * dL_da + (T[0][0] * T[1][1] + T[0][1] * T[1][0]) * dL_db + 2 * T[1][0] * T[1][1] * dL_dc;
        dcv[2] = 2 * T[0][0] * T[0][2] * dL_da + (T[0][0] * T[1][2] + T[0][2] * T[1][0]) * dL_db + 2 * T[1][0] * T[1][2] * dL_dc;
        dcv[4] = 2 * T[0][2] * T[0][1] * dL_da + (T[0][1] * T[1][2] + T[0][2] * T[1][1]) * dL_db + 2 * T[1][1] * T[1][2] * dL_dc;
    }
    float dT[2][3];
    for (int j = 0; j < 3; j++) {
        dT[0][j] = 2 * (T[0][0] * S[j][0] + T[0][1] * S[j][1] + T[0][2] * S[j][2]) * dL_da + (T[1][0] * S[j][0] + T[1][1] * S[j][1] + T[1][2] * S[j][2]) * dL_db;
        dT[1][j] = 2 * (T[1][0] * S[j][0] + T[1][1] * S[j][1] + T[1][2] * S[j][2]) * dL_dc + (T[0][0] * S[j][0] + T[0][1] * S[j][1] + T[0][2] * S[j][2]) * dL_db;
    }
    float dJ00 = 0, dJ02 = 0, dJ11 = 0, dJ12 = 0;
    for (int cc = 0; cc < 3; cc++) {
        dJ00 += view[cc * 4 + 0] * dT[0][cc]; dJ02 += view[cc * 4 + 2] * dT[0][cc];
        dJ11 += view[cc * 4 + 1] * dT[1][cc]; dJ12 += view[cc * 4 + 2] * dT[1][cc];
    }
    float tz = 1.f / t[2], tz2 = tz * tz, tz3 = tz2 * tz;
    float dtx = xm * -fx * tz2 * dJ02, dty = ym * -fy * tz2 * dJ12;
    float dtz = -fx * tz2 * dJ00 - fy * tz2 * dJ11 + (2 * fx * t[0]) * tz3 * dJ02 + (2 * fy * t[1]) * tz3 * dJ12;
    float gm[3];
    for (int k = 0; k < 3; k++) gm[k] = view[k * 4 + 0] * dtx + view[k * 4 + 1] * dty + view[k * 4 + 2] * dtz;
    float mh[4]; xf44(p, proj, mh);
    float mw = 1.0f / (mh[3] + 0.0000001f);
    float mul1 = mh[0] * mw * mw, mul2 = mh[1] * mw * mw;
    float d2x = g_mean2D[2 * (size_t)i], d2y = g_mean2D[2 * (size_t)i + 1];
    gm[0] += (proj[0] * mw - proj[3] * mul1) * d2x + (proj[1] * mw - proj[3] * mul2) * d2y;
    gm[1] += (proj[4] * mw - proj[7] * mul1) * d2x + (proj[5] * mw - proj[7] * mul2) * d2y;
    gm[2] += (proj[8] * mw - proj[11] * mul1) * d2x + (proj[9] * mw - proj[11] * mul2) * d2y;
    float gd = g_depth[i];
    gm[0] += view[2] * gd; gm[1] += view[6] * gd; gm[2] += view[10] * gd;
    for (int k = 0; k < 3; k++) g_means3D[3 * (size_t)i + k] = gm[k];
    const float* q = rots + 4 * (size_t)i;
    float r = q[0], x = q[1], y = q[2], z = q[3];
    float R[3][3] = {{1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y)},
                     {2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x)},
                     {2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y)}};
    float s[3] = {cam.scale_modifier * scales[3 * (size_t)i], cam.scale_modifier * scales[3 * (size_t)i + 1], cam.scale_modifier * scales[3 * (size_t)i + 2]};
    float dS[3][3] = {{dcv[0], 0.5f * dcv[1], 0.5f * dcv[2]}, {0.5f * dcv[1], dcv[3], 0.5f * dcv[4]}, {0.5f * dcv[2], 0.5f * dcv[4], dcv[5]}};
    float M[3][3], dM[3][3], dR[3][3];
    for (int a_ = 0; a_ < 3; a_++) for (int k = 0; k < 3; k++) M[a_][k] = R[a_][k] * s[k];
    for (int a_ = 0; a_ < 3; a_++) for (int k = 0; k < 3; k++) dM[a_][k] = 2.f * (dS[a_][0] * M[0][k] + dS[a_][1] * M[1][k] + dS[a_][2] * M[2][k]);
    for (int k = 0; k < 3; k++) {
        g_scales[3 * (size_t)i + k] = cam.scale_modifier * (R[0][k] * dM[0][k] + R[1][k] * dM[1][k] + R[2][k] * dM[2][k]);
        for (int a_ = 0; a_ < 3; a_++) dR[a_][k] = dM[a_][k] * s[k];
    }
    g_rots[4 * (size_t)i + 0] = 2 * z * (dR[1][0] - dR[0][1]) + 2 * y * (dR[0][2] - dR[2][0]) + 2 * x * (dR[2][1] - dR[1][2]);
    g_rots[4 * (size_t)i + 1] = 2 * y * (dR[0][1] + dR[1][0]) + 2 * z * (dR[0][2] + dR[2][0]) + 2 * r * (dR[2][1] - dR[1][2]) - 4 * x * (dR[2][2] + dR[1][1]);
    g_rots[4 * (size_t)i + 2] = 2 * x * (dR[0][1] + dR[1][0]) + 2 * r * (dR[0][2] - dR[2][0]) + 2 * z * (dR[2][1] + dR[1][2]) - 4 * y * (dR[2][2] + dR[0][0]);
    g_rots[4 * (size_t)i + 3] = 2 * r * (dR[1][0] - dR[0][1]) + 2 * x * (dR[0][2] + dR[2][0]) + 2 * y * (dR[2][1] + dR[1][2]) - 4 * z * (dR[1][1] + dR[0][0]);
}

extern "C" {
int refr_preprocess(int N, const float* means3D, const float* scales, const float* rots, const float* opac, const RefCamera* cam,
                    int* radii, float* xy, float* depth, float* cov3D, float* conic_op, int* rect, int* tiles_touched, void* stream) {
    k_preprocess<<<(N + 255) / 256, 256, 0, (cudaStream_t)stream>>>(N, means3D, scales, rots, opac, *cam, radii, (float2*)xy, depth, cov3D,
                                                                    (float4*)conic_op, (int4*)rect, tiles_touched);
    return (int)cudaGetLastError();
}
int refr_duplicate(int N, const int* radii, const float* depth, const int* rect, const int64_t* offsets_incl, int gx, int64_t* keys, int* vals,
                   void* stream) {
    k_duplicate<<<(N + 255) / 256, 256, 0, (cudaStream_t)stream>>>(N, radii, depth, (const int4*)rect, offsets_incl, gx, keys, vals);
    return (int)cudaGetLastError();
}
int refr_ranges(int64_t P, const int64_t* keys, uint32_t* ranges, void* stream) {
    if (P > 0) k_ranges<<<(unsigned)((P + 255) / 256), 256, 0, (cudaStream_t)stream>>>(P, keys, (uint2*)ranges);
    return (int)cudaGetLastError();
}
int refr_render(const RefCamera* cam, const uint32_t* ranges, const int* vals, const float* xy, const float* conic_op, const float* colors,
                const float* depth, float* out_color, float* out_depth, float* out_alpha, float* final_T, uint32_t* n_contrib, void* stream) {
    dim3 grid((cam->W + TILE - 1) / TILE, (cam->H + TILE - 1) / TILE), block(TILE, TILE);
    k_render<<<grid, block, 0, (cudaStream_t)stream>>>(*cam, (const uint2*)ranges, vals, (const float2*)xy, (const float4*)conic_op, colors, depth,
                                                       out_color, out_depth, out_alpha, final_T, n_contrib);
    return (int)cudaGetLastError();
}
int refr_render_bwd(const RefCamera* cam, const uint32_t* ranges, const int* vals, const float* xy, const float* conic_op, const float* colors,
                    const float* depth, const float* final_T, const uint32_t* n_contrib, const float* dL_dpix, const float* dL_ddepth,
                    const float* dL_dalpha, float* g_mean2D, float* g_conic, float* g_opac, float* g_color, float* g_depth, void* stream) {
    dim3 grid((cam->W + TILE - 1) / TILE, (cam->H + TILE - 1) / TILE), block(TILE, TILE);
    k_render_bwd<<<grid, block, 0, (cudaStream_t)stream>>>(*cam, (const uint2*)ranges, vals, (const float2*)xy, (const float4*)conic_op, colors,
                                                           depth, final_T, n_contrib, dL_dpix, dL_ddepth, dL_dalpha, g_mean2D, g_conic, g_opac,
                                                           g_color, g_depth);
    return (int)cudaGetLastError();
}
int refr_preprocess_bwd(int N, const float* means3D, const float* scales, const float* rots, const RefCamera* cam, const int* radii,
                        const float* cov3D, const float* g_mean2D, const float* g_conic, const float* g_depth, float* g_means3D, float* g_scales,
                        float* g_rots, void* stream) {
    k_preprocess_bwd<<<(N + 255) / 256, 256, 0, (cudaStream_t)stream>>>(N, means3D, scales, rots, *cam, radii, cov3D, g_mean2D, g_conic, g_depth,
                                                                        g_means3D, g_scales, g_rots);
    return (int)cudaGetLastError();
}
}
