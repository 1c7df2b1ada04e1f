// (f4) Inference output stage: the per-frame post-processing of Trainer.evaluate (core/trainer.py:1068-1084) +
// tensor2image (utils/image.py:52-61) fused into ONE kernel: planar fp32 render outputs -> interleaved uint8 frames
//   image    [3,H,W]            -> RGB  u8 [H,W,3]        (x * 255).clip(0, 255).astype(uint8)  (truncation, as numpy)
//   image_fg [3,H,W] + alpha    -> RGBA u8 [H,W,4]        concat_alpha (trainer.py:51-57)
//   depth    [H,W] / depth_div  -> L    u8 [H,W]          trainer.py:1076-1077 (depth / 3.0)
//   alpha    [H,W]              -> L    u8 [H,W]
// Any output may be NULL.  Pure streaming: 4 pixels per thread, 32-bit / 96-bit / 128-bit stores.
#include "common.cuh"

namespace dwg {
namespace {

__device__ __forceinline__ uint32_t to_u8(float x) {
    const float v = fminf(fmaxf(x * 255.0f, 0.0f), 255.0f);      // NaN -> 0 like numpy's clip+astype on most platforms is undefined; 0 is the safe choice
    return (uint32_t)(int)v;                                     // truncation toward zero == astype(np.uint8) on [0, 255]
}

__global__ void __launch_bounds__(256)
frame_pack_kernel(const float* __restrict__ image, const float* __restrict__ image_fg, const float* __restrict__ depth,
                  const float* __restrict__ alpha, uint8_t* __restrict__ rgb, uint8_t* __restrict__ rgba, uint8_t* __restrict__ depth_u8,
                  uint8_t* __restrict__ alpha_u8, int64_t HW, float inv_depth_div) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;      // group of 4 pixels
    const int64_t p0 = q * 4;
    if (p0 >= HW) return;
    const bool full = p0 + 4 <= HW && (HW & 3) == 0;          // vector path needs 16-byte aligned planes (plane stride = HW floats)
    float a4[4] = {0.f, 0.f, 0.f, 0.f};
    if (alpha) {
        if (full) { const float4 t = *reinterpret_cast<const float4*>(alpha + p0); a4[0] = t.x; a4[1] = t.y; a4[2] = t.z; a4[3] = t.w; }
        else for (int i = 0; i < 4 && p0 + i < HW; i++) a4[i] = alpha[p0 + i];
    }
    auto load4 = [&](const float* src, float (&v)[4]) {
        if (full) { const float4 t = *reinterpret_cast<const float4*>(src + p0); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
        else { for (int i = 0; i < 4; i++) v[i] = (p0 + i < HW) ? src[p0 + i] : 0.f; }
    };
    if (rgb && image) {
        float r[4], g[4], b[4];
        load4(image, r); load4(image + HW, g); load4(image + 2 * HW, b);
        uint8_t o[12];
#pragma unroll
        for (int i = 0; i < 4; i++) { o[3 * i] = (uint8_t)to_u8(r[i]); o[3 * i + 1] = (uint8_t)to_u8(g[i]); o[3 * i + 2] = (uint8_t)to_u8(b[i]); }
        if (full) {
            uint32_t* dst = reinterpret_cast<uint32_t*>(rgb + p0 * 3);         // p0 % 4 == 0 -> 12-byte groups are 4-byte aligned
            dst[0] = *reinterpret_cast<uint32_t*>(o); dst[1] = *reinterpret_cast<uint32_t*>(o + 4); dst[2] = *reinterpret_cast<uint32_t*>(o + 8);
        } else {
            for (int i = 0; i < 4 && p0 + i < HW; i++) { rgb[(p0 + i) * 3] = o[3 * i]; rgb[(p0 + i) * 3 + 1] = o[3 * i + 1]; rgb[(p0 + i) * 3 + 2] = o[3 * i + 2]; }
        }
    }
    if (rgba && image_fg) {
        float r[4], g[4], b[4];
        load4(image_fg, r); load4(image_fg + HW, g); load4(image_fg + 2 * HW, b);
        uint32_t o[4];
#pragma unroll
        for (int i = 0; i < 4; i++) o[i] = to_u8(r[i]) | (to_u8(g[i]) << 8) | (to_u8(b[i]) << 16) | (to_u8(a4[i]) << 24);
        if (full) *reinterpret_cast<uint4*>(rgba + p0 * 4) = make_uint4(o[0], o[1], o[2], o[3]);
        else for (int i = 0; i < 4 && p0 + i < HW; i++) reinterpret_cast<uint32_t*>(rgba)[p0 + i] = o[i];
    }
    if (depth_u8 && depth) {
        float d[4];
        load4(depth, d);
        const uint32_t o = to_u8(d[0] * inv_depth_div) | (to_u8(d[1] * inv_depth_div) << 8) | (to_u8(d[2] * inv_depth_div) << 16) | (to_u8(d[3] * inv_depth_div) << 24);
        if (full) *reinterpret_cast<uint32_t*>(depth_u8 + p0) = o;
        else for (int i = 0; i < 4 && p0 + i < HW; i++) depth_u8[p0 + i] = (uint8_t)(o >> (8 * i));
    }
    if (alpha_u8 && alpha) {
        const uint32_t o = to_u8(a4[0]) | (to_u8(a4[1]) << 8) | (to_u8(a4[2]) << 16) | (to_u8(a4[3]) << 24);
        if (full) *reinterpret_cast<uint32_t*>(alpha_u8 + p0) = o;
        else for (int i = 0; i < 4 && p0 + i < HW; i++) alpha_u8[p0 + i] = (uint8_t)(o >> (8 * i));
    }
}

}  // namespace
}  // namespace dwg

using namespace dwg;

extern "C" int dwg_frame_pack(const float* image, const float* image_fg, const float* depth, const float* alpha,
                              uint8_t* rgb, uint8_t* rgba_fg, uint8_t* depth_u8, uint8_t* alpha_u8,
                              int H, int W, float depth_div, void* stream) {
    DWG_REQUIRE(H > 0 && W > 0 && depth_div != 0.f, "bad arguments");
    DWG_REQUIRE(!rgb || image, "rgb output needs image");
    DWG_REQUIRE(!rgba_fg || (image_fg && alpha), "rgba output needs image_fg and alpha");
    DWG_REQUIRE(!depth_u8 || depth, "depth output needs depth");
    DWG_REQUIRE(!alpha_u8 || alpha, "alpha output needs alpha");
    auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    DWG_REQUIRE(al16(image) && al16(image_fg) && al16(depth) && al16(alpha) && al16(rgb) && al16(rgba_fg) && al16(depth_u8) && al16(alpha_u8),
                "buffers must be 16-byte aligned");
    const int64_t HW = (int64_t)H * W;
    const int64_t groups = (HW + 3) / 4;
    frame_pack_kernel<<<(unsigned)((groups + 255) / 256), 256, 0, (cudaStream_t)stream>>>(image, image_fg, depth, alpha, rgb, rgba_fg, depth_u8,
                                                                                          alpha_u8, HW, 1.0f / depth_div);
    return check_launch("dwg_frame_pack");
}
