// Rasteriser stage 2 (R12): per-tile depth sort + instance record packing.
//
// One CTA per tile.  The tile's (depth bits << 32 | idx) keys are sorted in shared memory
// (bitonic network on SORT_CHUNK-key runs; tiles with more instances merge their runs through
// the L2-resident ping-pong buffer with rank-by-binary-search passes).  Keys within a tile are
// unique (idx is unique), so the order equals upstream's stable radix sort on
// (tile << 32 | depth) with ascending-idx emission order.  The CTA then gathers the
// per-Gaussian data ONCE into contiguous 48-byte records so that the forward and backward
// blend kernels can stream them with bulk TMA copies instead of indexed gathers.
#include "raster_common.cuh"

namespace dwg {
namespace raster {

__device__ __forceinline__ void bitonic_sort_smem(uint64_t* s, int n_pow2) {
    for (int k = 2; k <= n_pow2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < (n_pow2 >> 1); t += blockDim.x) {
                // pair index: insert a zero bit at position log2(j)
                const int lo = ((t & ~(j - 1)) << 1) | (t & (j - 1));
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
tile_sort_pack_kernel(int T, BinView b, GeomView g, const float* __restrict__ colors, int write_keys) {
    __shared__ uint64_t s_keys[SORT_CHUNK];
    const int tile = blockIdx.x;
    const uint2 rg = b.ranges[tile];
    const int n = (int)(rg.y - rg.x);
    if (n <= 0) return;
    uint64_t* seg = b.inst_key + rg.x;
    uint64_t* tmp = b.inst_tmp + rg.x;
    // ---- sort runs of SORT_CHUNK in shared memory ----
    for (int base = 0; base < n; base += SORT_CHUNK) {
        const int m = min(SORT_CHUNK, n - base);
        int p2 = 32;
        while (p2 < m) p2 <<= 1;
        for (int i = threadIdx.x; i < p2; i += blockDim.x) s_keys[i] = i < m ? seg[base + i] : ~0ull;
        __syncthreads();
        bitonic_sort_smem(s_keys, p2);
        for (int i = threadIdx.x; i < m; i += blockDim.x) seg[base + i] = s_keys[i];
        __syncthreads();
    }
    // ---- merge runs (only for tiles with n > SORT_CHUNK) ----
    uint64_t* src = seg;
    uint64_t* dst = tmp;
    for (int run = SORT_CHUNK; run < n; run <<= 1) {
        __threadfence_block();
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const int pair = i / (2 * run);
            const int a0 = pair * 2 * run, a1 = min(a0 + run, n), b1 = min(a0 + 2 * run, n);
            const uint64_t key = src[i];
            int pos;
            if (i < a1) {                       // element of run A: count elements of B smaller than key
                int lo = a1, hi = b1;
                while (lo < hi) { const int mid = (lo + hi) >> 1; if (src[mid] < key) lo = mid + 1; else hi = mid; }
                pos = a0 + (i - a0) + (lo - a1);
            } else {                            // element of run B: count elements of A smaller than key
                int lo = a0, hi = a1;
                while (lo < hi) { const int mid = (lo + hi) >> 1; if (src[mid] < key) lo = mid + 1; else hi = mid; }
                pos = a0 + (i - a1) + (lo - a0);
            }
            dst[pos] = key;
        }
        uint64_t* sw = src; src = dst; dst = sw;
    }
    __threadfence_block();
    __syncthreads();
    // ---- pack records (and upstream-compatible sorted key/value lists) ----
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const uint64_t key = src[i];
        const uint32_t idx = (uint32_t)key;
        const uint32_t dbits = (uint32_t)(key >> 32);
        if (src != seg) seg[i] = key;
        const float2 xy = g.xy[idx];
        const float4 co = g.conic_opacity[idx];
        Rec rc;
        rc.x = xy.x; rc.y = xy.y; rc.cx = co.x; rc.cy = co.y; rc.cz = co.z; rc.op = co.w;
        rc.r = colors[3 * (size_t)idx]; rc.g = colors[3 * (size_t)idx + 1]; rc.b = colors[3 * (size_t)idx + 2];
        rc.depth = __uint_as_float(dbits); rc.idx = idx; rc.pad = 0;
        b.recs[rg.x + i] = rc;
        if (write_keys) {
            b.keys_out[rg.x + i] = ((uint64_t)tile << 32) | dbits;
            b.vals_out[rg.x + i] = idx;
        }
    }
}

int launch_sort(int T, BinView b, GeomView g, const float* colors, int write_keys, cudaStream_t st) {
    tile_sort_pack_kernel<<<T, 256, 0, st>>>(T, b, g, colors, write_keys);
    return check_launch("raster tile sort");
}

}  // namespace raster
}  // namespace dwg
