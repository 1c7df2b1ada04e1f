// Rasteriser stage 1 (R12): per-Gaussian preprocess + tile histogram, tile scan, instance scatter.
//
// THIS FILE IS COMPILED WITH -fmad=false: every fp32 operation below is a single IEEE
// operation in source order, which is the specification shared with oracle/oracle_c.c
// (orc_raster_preprocess) and makes radii, tile rectangles, depths (= sort keys) bit-exact.
//
// B200-first binning (instead of upstream's global 64-bit radix sort + host readback of
// num_rendered): (1) preprocess also histograms instances per tile with atomics,
// (2) one CTA scans the <= 4096 tile counts -> tile ranges for free, total P stays on the
// device, (3) instances are scattered into their tile's segment, (4) each tile's segment is
// sorted locally by (depth bits, idx) in shared memory (raster_sort.cu).  The result equals
// upstream's stable radix sort of (tile << 32 | depth) keys.
#include "raster_common.cuh"

namespace dwg {
namespace raster {

__global__ void __launch_bounds__(256)
preprocess_kernel(int64_t N, const float* __restrict__ means3D, const float* __restrict__ scales,
                  const float* __restrict__ rots, const float* __restrict__ opacities,
                  const DwgRasterCamera cam_val, const DwgRasterCamera* __restrict__ cam_dev, GeomView g,
                  int32_t* __restrict__ radii, uint32_t* __restrict__ tile_count) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    // a device-resident camera (updated between CUDA-graph replays) overrides the by-value copy
    const DwgRasterCamera cam = cam_dev ? *cam_dev : cam_val;
    const int H = cam.image_height, W = cam.image_width;
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
    const float* view = cam.viewmatrix;
    const float* proj = cam.projmatrix;
    radii[i] = 0;
    g.tiles_touched[i] = 0;
    g.rect[i] = make_int4(0, 0, 0, 0);
    g.xy[i] = make_float2(0.f, 0.f);
    g.depth[i] = 0.f;
    g.conic_opacity[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    const float px_ = means3D[3 * i], py_ = means3D[3 * i + 1], pz_ = means3D[3 * i + 2];
    float pv[3];
    pv[0] = view[0] * px_ + view[4] * py_ + view[8] * pz_ + view[12];
    pv[1] = view[1] * px_ + view[5] * py_ + view[9] * pz_ + view[13];
    pv[2] = view[2] * px_ + view[6] * py_ + view[10] * pz_ + view[14];
    float c6[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    bool ok = pv[2] > 0.2f;
    float conic[3] = {0.f, 0.f, 0.f};
    float pixx = 0.f, pixy = 0.f;
    int r = 0, x0 = 0, y0 = 0, x1 = 0, y1 = 0;
    if (ok) {
        float ph[4];
        ph[0] = proj[0] * px_ + proj[4] * py_ + proj[8] * pz_ + proj[12];
        ph[1] = proj[1] * px_ + proj[5] * py_ + proj[9] * pz_ + proj[13];
        ph[3] = proj[3] * px_ + proj[7] * py_ + proj[11] * pz_ + proj[15];
        const float pw = 1.0f / (ph[3] + 0.0000001f);
        const float ppx = ph[0] * pw, ppy = ph[1] * pw;
        // cov3D = R diag(s)^2 R^T
        {
            const float qr = rots[4 * i], qx = rots[4 * i + 1], qy = rots[4 * i + 2], qz = rots[4 * i + 3];
            const float R[3][3] = {
                {1.f - 2.f * (qy * qy + qz * qz), 2.f * (qx * qy - qr * qz), 2.f * (qx * qz + qr * qy)},
                {2.f * (qx * qy + qr * qz), 1.f - 2.f * (qx * qx + qz * qz), 2.f * (qy * qz - qr * qx)},
                {2.f * (qx * qz - qr * qy), 2.f * (qy * qz + qr * qx), 1.f - 2.f * (qx * qx + qy * qy)}};
            const float s[3] = {cam.scale_modifier * scales[3 * i], cam.scale_modifier * scales[3 * i + 1],
                                cam.scale_modifier * scales[3 * i + 2]};
            float M[3][3];
#pragma unroll
            for (int a = 0; a < 3; a++)
#pragma unroll
                for (int k = 0; k < 3; k++) M[a][k] = R[a][k] * s[k];
            int t = 0;
#pragma unroll
            for (int a = 0; a < 3; a++)
#pragma unroll
                for (int b = a; b < 3; b++) c6[t++] = M[a][0] * M[b][0] + M[a][1] * M[b][1] + M[a][2] * M[b][2];
        }
        // EWA cov2D
        const float fx = W / (2.0f * cam.tanfovx), fy = H / (2.0f * cam.tanfovy);
        float tx = pv[0], ty = pv[1];
        const float tz = pv[2];
        const float limx = 1.3f * cam.tanfovx, limy = 1.3f * cam.tanfovy;
        const float txtz = tx / tz, tytz = ty / tz;
        tx = fminf(limx, fmaxf(-limx, txtz)) * tz;
        ty = fminf(limy, fmaxf(-limy, tytz)) * tz;
        const float J00 = fx / tz, J02 = -(fx * tx) / (tz * tz);
        const float J11 = fy / tz, J12 = -(fy * ty) / (tz * tz);
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
        const float ca = TS[0][0] * T[0][0] + TS[0][1] * T[0][1] + TS[0][2] * T[0][2] + 0.3f;
        const float cb = TS[0][0] * T[1][0] + TS[0][1] * T[1][1] + TS[0][2] * T[1][2];
        const float cc = TS[1][0] * T[1][0] + TS[1][1] * T[1][1] + TS[1][2] * T[1][2] + 0.3f;
        const float det = ca * cc - cb * cb;
        if (det == 0.0f) ok = false;
        if (ok) {
            const float det_inv = 1.f / det;
            conic[0] = cc * det_inv; conic[1] = -cb * det_inv; conic[2] = ca * det_inv;
            const float mid = 0.5f * (ca + cc);
            const float sq = sqrtf(fmaxf(0.1f, mid * mid - det));
            const float l1 = mid + sq, l2 = mid - sq;
            const float my_radius = ceilf(3.f * sqrtf(fmaxf(l1, l2)));
            pixx = (float)((((double)ppx + 1.0) * (double)W - 1.0) * 0.5);      // ndc2Pix in double
            pixy = (float)((((double)ppy + 1.0) * (double)H - 1.0) * 0.5);
            r = (int)my_radius;
            x0 = (int)((pixx - r) / TILE); y0 = (int)((pixy - r) / TILE);
            x1 = (int)((pixx + r + TILE - 1) / TILE); y1 = (int)((pixy + r + TILE - 1) / TILE);
            x0 = min(gx, max(0, x0)); y0 = min(gy, max(0, y0));
            x1 = min(gx, max(0, x1)); y1 = min(gy, max(0, y1));
            if ((x1 - x0) * (y1 - y0) == 0) ok = false;
        }
    }
#pragma unroll
    for (int k = 0; k < 6; k++) g.cov3D[6 * i + k] = c6[k];
    if (!ok) return;
    radii[i] = r;
    g.depth[i] = pv[2];
    g.xy[i] = make_float2(pixx, pixy);
    g.conic_opacity[i] = make_float4(conic[0], conic[1], conic[2], opacities[i]);
    g.rect[i] = make_int4(x0, y0, x1, y1);
    g.tiles_touched[i] = (uint32_t)((x1 - x0) * (y1 - y0));
    const int sub = (int)(i & (BIN_SUB - 1));
    for (int y = y0; y < y1; y++)
        for (int x = x0; x < x1; x++) atomicAdd(&tile_count[(y * gx + x) * BIN_SUB + sub], 1u);
}

// One CTA: exclusive scan of the tile counts (sum of the BIN_SUB sub-counters of each tile), ranges, total P and overflow flag.
__global__ void __launch_bounds__(1024)
tile_scan_kernel(int T, BinView b, int64_t P_cap, int32_t* __restrict__ status) {
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    uint32_t max_load = 0;
    for (int base = 0; base < T; base += 1024) {
        const int t = base + threadIdx.x;
        uint32_t sc[BIN_SUB];
        uint32_t v = 0u;
#pragma unroll
        for (int k = 0; k < BIN_SUB; k++) { sc[k] = t < T ? b.tile_count[t * BIN_SUB + k] : 0u; v += sc[k]; }
        max_load = max(max_load, v);
        uint32_t incl = v;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const uint32_t n = __shfl_up_sync(0xffffffffu, incl, off);
            if ((threadIdx.x & 31) >= off) incl += n;
        }
        if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = incl;
        __syncthreads();
        if (threadIdx.x < 32) {
            uint32_t w = s_warp[threadIdx.x];
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const uint32_t n = __shfl_up_sync(0xffffffffu, w, off);
                if (threadIdx.x >= off) w += n;
            }
            s_warp[threadIdx.x] = w;
        }
        __syncthreads();
        const uint32_t warp_off = (threadIdx.x >> 5) ? s_warp[(threadIdx.x >> 5) - 1] : 0u;
        const uint32_t carry = s_carry;
        const uint32_t excl = carry + warp_off + incl - v;
        if (t < T) {
            b.tile_start[t] = excl;
            uint32_t run = excl;
#pragma unroll
            for (int k = 0; k < BIN_SUB; k++) { b.sub_start[t * BIN_SUB + k] = run; run += sc[k]; }
            const uint32_t e = excl + v;
            // clamp to the granted capacity so later stages never touch memory beyond it
            const uint32_t cs = (uint32_t)min((int64_t)excl, P_cap), ce = (uint32_t)min((int64_t)e, P_cap);
            b.ranges[t] = v ? make_uint2(cs, ce) : make_uint2(0u, 0u);
        }
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = carry + warp_off + incl;
        __syncthreads();
    }
    // block max of tile load
    for (int off = 16; off > 0; off >>= 1) max_load = max(max_load, __shfl_xor_sync(0xffffffffu, max_load, off));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = max_load;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t m = 0;
        for (int w = 0; w < 32; w++) m = max(m, s_warp[w]);
        const uint32_t P = s_carry;
        b.tile_start[T] = P;
        status[0] = ((int64_t)P > P_cap) ? 1 : 0;
        status[1] = (int32_t)P;
        status[2] = (int32_t)m;
        status[3] = 0;
        s_carry = 0;
    }
    __syncthreads();
    // ---- sort-run table: every tile segment is cut into runs of <= SORT_CHUNK instances ----
    for (int base = 0; base < T; base += 1024) {
        const int t = base + threadIdx.x;
        uint2 rg = make_uint2(0u, 0u);
        if (t < T) rg = b.ranges[t];
        const uint32_t n = rg.y - rg.x;
        const uint32_t v = (n + SORT_CHUNK - 1) / SORT_CHUNK;
        uint32_t incl = v;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const uint32_t nb = __shfl_up_sync(0xffffffffu, incl, off);
            if ((threadIdx.x & 31) >= off) incl += nb;
        }
        if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = incl;
        __syncthreads();
        if (threadIdx.x < 32) {
            uint32_t w = s_warp[threadIdx.x];
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const uint32_t nb = __shfl_up_sync(0xffffffffu, w, off);
                if (threadIdx.x >= off) w += nb;
            }
            s_warp[threadIdx.x] = w;
        }
        __syncthreads();
        const uint32_t warp_off = (threadIdx.x >> 5) ? s_warp[(threadIdx.x >> 5) - 1] : 0u;
        const uint32_t carry = s_carry;
        uint32_t r0 = carry + warp_off + incl - v;
        for (uint32_t k = 0; k < v; k++)
            b.runs[r0 + k] = make_uint2(rg.x + k * SORT_CHUNK, min((uint32_t)SORT_CHUNK, n - k * SORT_CHUNK));
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = carry + warp_off + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) b.n_runs[0] = s_carry;
}

__global__ void __launch_bounds__(256)
scatter_kernel(int64_t N, int gx, GeomView g, BinView b, int64_t P_cap) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    if (g.tiles_touched[i] == 0) return;
    const int4 rc = g.rect[i];
    const uint64_t key = ((uint64_t)__float_as_uint(g.depth[i]) << 32) | (uint64_t)(uint32_t)i;
    for (int y = rc.y; y < rc.w; y++)
        for (int x = rc.x; x < rc.z; x++) {
            const int t = y * gx + x;
            const int ts = t * BIN_SUB + (int)(i & (BIN_SUB - 1));
            const uint32_t slot = b.sub_start[ts] + atomicAdd(&b.tile_fill[ts], 1u);
            if ((int64_t)slot < P_cap) { b.inst_key[slot] = key; b.inst_tile[slot] = (uint32_t)t; }
        }
}

int launch_pre(const DwgRasterCamera& cam, const DwgRasterCamera* cam_dev, int64_t N, const float* means3D, const float* opacities,
               const float* scales, const float* rots, GeomView g, BinView b, int T, int64_t P_cap,
               int32_t* radii, int32_t* status, cudaStream_t st) {
    const int gx = (cam.image_width + TILE - 1) / TILE;
    cudaMemsetAsync(b.tile_count, 0, BinView::header_bytes(T), st);      // counts, cursors, starts, ranges
    if (N > 0) preprocess_kernel<<<(unsigned)ceil_div(N, 256), 256, 0, st>>>(N, means3D, scales, rots, opacities, cam, cam_dev, g, radii, b.tile_count);
    tile_scan_kernel<<<1, 1024, 0, st>>>(T, b, P_cap, status);
    if (N > 0) scatter_kernel<<<(unsigned)ceil_div(N, 256), 256, 0, st>>>(N, gx, g, b, P_cap);
    return check_launch("raster preprocess/scan/scatter");
}

}  // namespace raster
}  // namespace dwg
