// Rasteriser stage 2 (R12): per-tile depth sort + instance record packing, load-balanced.
//
// After the scatter every tile owns a contiguous, unsorted segment of (depth bits << 32 | idx)
// keys.  Sorting is split so that no CTA ever owns a whole heavy tile:
//   1. run sort   - one CTA per run of <= SORT_CHUNK keys (run table built by the scan kernel):
//                   bitonic network in shared memory;
//   2. merge pass - one THREAD per instance: rank = own index + binary search in the sibling run
//                   (keys are unique), written to the ping-pong buffer; log2(runs) passes, each
//                   a fixed launch that exits immediately when no tile needs it;
//   3. pack       - one thread per instance gathers the per-Gaussian data ONCE into contiguous
//                   48-byte records (+ the upstream-compatible sorted key / value lists), so the
//                   blend kernels stream them with bulk TMA copies instead of indexed gathers.
// The resulting order equals upstream's stable radix sort of (tile << 32 | depth) keys emitted in
// ascending Gaussian order.
#include "raster_common.cuh"

namespace dwg {
namespace raster {

__device__ __forceinline__ void bitonic_sort_smem(uint64_t* s, int n_pow2) {
    for (int k = 2; k <= n_pow2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < (n_pow2 >> 1); t += blockDim.x) {
                const int lo = ((t & ~(j - 1)) << 1) | (t & (j - 1));     // insert a 0 bit at log2(j)
                const int hi = lo | j;
                const bool up = ((lo & k) == 0);
                const uint64_t a = s[lo], b = s[hi];
                if ((a > b) == up) { s[lo] = b; s[hi] = a; }
            }
            __syncthreads();
        }
    }
}

__global__ void __launch_bounds__(256)
run_sort_kernel(BinView b) {
    __shared__ uint64_t s_keys[SORT_CHUNK];
    const uint32_t n_runs = b.n_runs[0];
    for (uint32_t r = blockIdx.x; r < n_runs; r += gridDim.x) {
        const uint2 run = b.runs[r];
        const int m = (int)run.y;
        uint64_t* seg = b.inst_key + run.x;
        int p2 = 32;
        while (p2 < m) p2 <<= 1;
        for (int i = threadIdx.x; i < p2; i += blockDim.x) s_keys[i] = i < m ? seg[i] : ~0ull;
        __syncthreads();
        bitonic_sort_smem(s_keys, p2);
        for (int i = threadIdx.x; i < m; i += blockDim.x) seg[i] = s_keys[i];
        __syncthreads();
    }
}

// pass p merges sibling runs of length SORT_CHUNK << p inside every tile that still needs it
__global__ void __launch_bounds__(256)
merge_pass_kernel(BinView b, int pass, const int32_t* __restrict__ status, int64_t P_cap) {
    const uint32_t run = (uint32_t)SORT_CHUNK << pass;
    if ((uint32_t)status[2] <= run) return;                    // no tile has more than `run` instances
    const int64_t P = min((int64_t)status[1], P_cap);
    const uint64_t* __restrict__ src = (pass & 1) ? b.inst_tmp : b.inst_key;
    uint64_t* __restrict__ dst = (pass & 1) ? b.inst_key : b.inst_tmp;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P; i += (int64_t)gridDim.x * blockDim.x) {
        const uint2 rg = b.ranges[b.inst_tile[i]];
        const uint32_t n = rg.y - rg.x;
        if (n <= run) continue;                                // this tile is already fully sorted
        const uint32_t li = (uint32_t)i - rg.x;
        const uint32_t a0 = (li / (2 * run)) * (2 * run);
        const uint32_t a1 = min(a0 + run, n), b1 = min(a0 + 2 * run, n);
        const uint64_t* sp = src + rg.x;
        const uint64_t key = sp[li];
        uint32_t pos;
        if (li < a1) {                                          // run A: count keys of B below
            uint32_t lo = a1, hi = b1;
            while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (sp[mid] < key) lo = mid + 1; else hi = mid; }
            pos = li + (lo - a1);
        } else {                                                // run B: count keys of A below
            uint32_t lo = a0, hi = a1;
            while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (sp[mid] < key) lo = mid + 1; else hi = mid; }
            pos = a0 + (li - a1) + (lo - a0);
        }
        dst[rg.x + pos] = key;
    }
}

__global__ void __launch_bounds__(256)
pack_kernel(BinView b, GeomView g, const float* __restrict__ colors, const int32_t* __restrict__ status,
            int64_t P_cap, int write_keys) {
    const int64_t P = min((int64_t)status[1], P_cap);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P; i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t tile = b.inst_tile[i];
        const uint2 rg = b.ranges[tile];
        const int np = merge_passes(rg.y - rg.x);
        const uint64_t key = (np & 1) ? b.inst_tmp[i] : b.inst_key[i];
        const uint32_t idx = (uint32_t)key;
        const uint32_t dbits = (uint32_t)(key >> 32);
        const float2 xy = g.xy[idx];
        const float4 co = g.conic_opacity[idx];
        // conservative extent of the alpha >= 1/255 ellipse:  Q(d) <= q = 2 ln(255 op)
        float ex = 0.f, ey = 0.f;
        const float q = 2.0f * __logf(255.0f * co.w) + 0.05f;
        if (q > 0.f) {
            const float det = co.x * co.z - co.y * co.y;
            if (det > 0.f) {
                ex = fminf(sqrtf(q * co.z / det) * 1.01f + 0.05f, 60000.f);
                ey = fminf(sqrtf(q * co.x / det) * 1.01f + 0.05f, 60000.f);
            } else {
                ex = ey = 60000.f;
            }
        } else if (!(co.w <= 1.0f / 255.0f)) {
            ex = ey = 60000.f;                                   // NaN or unexpected: never cull
        } else {
            ex = ey = -1.0f;                                     // op < 1/255: can never contribute
        }
        const __half2 eh = __halves2half2(__float2half_ru(ex), __float2half_ru(ey));
        Rec rc;
        rc.x = xy.x; rc.y = xy.y; rc.ext = *reinterpret_cast<const uint32_t*>(&eh); rc.idx = idx;
        rc.cx = co.x; rc.cy = co.y; rc.cz = co.z; rc.op = co.w;
        rc.r = colors[3 * (size_t)idx]; rc.g = colors[3 * (size_t)idx + 1]; rc.b = colors[3 * (size_t)idx + 2];
        rc.depth = __uint_as_float(dbits);
        b.recs[i] = rc;
        if (write_keys) {
            b.keys_out[i] = ((uint64_t)tile << 32) | dbits;
            b.vals_out[i] = idx;
        }
    }
}

int launch_sort(int T, BinView b, GeomView g, const float* colors, const int32_t* status, int64_t P_cap,
                int write_keys, cudaStream_t st) {
    const int64_t R_cap = BinView::run_cap(P_cap, T);
    const int64_t rs_grid = R_cap < 8 * kNumSMs ? R_cap : 8 * kNumSMs;
    run_sort_kernel<<<(unsigned)rs_grid, 256, 0, st>>>(b);
    const int64_t want = ceil_div(P_cap, 256);
    const unsigned grid = (unsigned)(want < 8 * kNumSMs ? want : 8 * kNumSMs);
    for (int p = 0; p < MAX_MERGE_PASSES; p++) merge_pass_kernel<<<grid, 256, 0, st>>>(b, p, status, P_cap);
    pack_kernel<<<grid, 256, 0, st>>>(b, g, colors, status, P_cap, write_keys);
    return check_launch("raster sort/pack");
}

}  // namespace raster
}  // namespace dwg
