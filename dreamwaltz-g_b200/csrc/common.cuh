// Shared helpers for libdwg_sm100.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/dwg.h"

namespace dwg {

void set_error(const char* fmt, ...);

inline int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return DWG_ERR_CUDA;
    }
    return DWG_OK;
}

#define DWG_REQUIRE(cond, msg)                                   \
    do {                                                         \
        if (!(cond)) {                                           \
            ::dwg::set_error("%s: %s", __func__, msg);           \
            return DWG_ERR_INVALID;                              \
        }                                                        \
    } while (0)

constexpr int kNumSMs = 148;   // B200

__host__ __device__ inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// 128-bit streaming load / store (read-once data: bypass L1 allocation)
__device__ __forceinline__ float4 ld_stream_f4(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream_f4(float4* p, const float4& v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w));
}

// Copy `n` floats from global `src` (16-byte aligned when `aligned`) into shared `dst`
// with the whole CTA; coalesced 128-bit loads where alignment allows.
__device__ __forceinline__ void cta_load_floats(float* dst, const float* src, int n, bool aligned) {
    if (aligned) {
        const int n4 = n >> 2;
        const float4* s4 = reinterpret_cast<const float4*>(src);
        float4* d4 = reinterpret_cast<float4*>(dst);
        for (int i = threadIdx.x; i < n4; i += blockDim.x) d4[i] = ld_stream_f4(s4 + i);
        for (int i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
    } else {
        for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
    }
}
__device__ __forceinline__ void cta_store_floats(float* dst, const float* src, int n, bool aligned) {
    if (aligned) {
        const int n4 = n >> 2;
        const float4* s4 = reinterpret_cast<const float4*>(src);
        float4* d4 = reinterpret_cast<float4*>(dst);
        for (int i = threadIdx.x; i < n4; i += blockDim.x) st_stream_f4(d4 + i, s4[i]);
        for (int i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
    } else {
        for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
    }
}

}  // namespace dwg
