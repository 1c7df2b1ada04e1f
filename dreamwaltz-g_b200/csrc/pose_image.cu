// (f1) Per-step ControlNet condition producer on the GPU: SMPL-X keypoints -> projected, occlusion-tested 2-D keypoints ->
// OpenPose-style skeleton image, without leaving the device.
//
// Replaces the CPU stage the reference runs for every training view (core/human/smpl_condition.py:191-235 export_pose:
// numpy projection, Embree ray casting against the SMPL-X mesh for occlusion culling (utils/open3d.py:8-45,
// smpl_condition.py:82-141), cv2 drawing (core/human/open_pose.py:48-333 adaptive_draw_poses), PIL -> tensor
// (core/guidance/controlnet.py:33-55)).
//   * pose_project_kernel: world keypoints -> camera -> pixels (K . X / Z, utils/point3d.py:32-42); behind the camera =>
//     absent; occlusion = depth test against the avatar's OWN rendered depth / alpha (the rasteriser's outputs of the same
//     view) instead of a BVH over the template mesh: occluded when the render is opaque there and closer by more than the
//     reference's thresholds (body 0.2, face 0.02, hand 0.2).
//   * pose_draw_kernel: one thread per pixel replays the reference's draw ORDER with uint8 semantics: filled body circles,
//     17 limb ellipses alpha-blended one after the other (cv2.addWeighted 0.4 / 0.6, rounded), blue hand circles + hsv
//     coloured edges, white face dots.  cv2's circle is exactly dx^2 + dy^2 <= r^2; its 1-degree ellipse polygon and thick
//     lines are matched analytically (semi-axes + 0.5; capsule), so images agree with cv2 up to boundary pixels
//     (tests/test_pose_condition.py: mismatching pixels < 0.5 % of the drawn area ... see the stated bound there).
#include "common.cuh"

namespace dwg {
namespace pose {

constexpr int NK = 128;      // 18 body + 21 left hand + 21 right hand + 51 landmarks + 17 contour (smpl_condition.py:22)

__global__ void pose_project_kernel(const float* __restrict__ kp_world, int K, const float* __restrict__ ext, const float* __restrict__ intr_dev,
                                    float fx, float fy, float cx, float cy,
                                    const float* __restrict__ depth, const float* __restrict__ alpha, int Hd, int Wd, float cond_w, float cond_h,
                                    float thr_body, float thr_face, float thr_hand, float* __restrict__ kp2d) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    if (intr_dev) { fx = intr_dev[0]; fy = intr_dev[1]; cx = intr_dev[2]; cy = intr_dev[3]; }      // device-resident camera (CUDA-graph replay)
    const float X = kp_world[3 * k], Y = kp_world[3 * k + 1], Z = kp_world[3 * k + 2];
    const float xc = ext[0] * X + ext[1] * Y + ext[2] * Z + ext[3];
    const float yc = ext[4] * X + ext[5] * Y + ext[6] * Z + ext[7];
    const float zc = ext[8] * X + ext[9] * Y + ext[10] * Z + ext[11];
    float u = nanf(""), v = nanf("");
    if (!(zc < 0.f)) {                                           // smpl_condition.py:208
        u = (fx * xc + cx * zc) / zc;
        v = (fy * yc + cy * zc) / zc;
        if (depth && alpha) {
            // OcclusionCulling group of keypoint k (smpl_condition.py:88-91)
            const bool face = k == 0 || (k >= 14 && k <= 17) || k >= 60;
            const bool hand = k >= 18 && k < 60;
            const float thr = face ? thr_face : (hand ? thr_hand : thr_body);
            const int px = (int)(u / cond_w * (float)Wd), py = (int)(v / cond_h * (float)Hd);
            if (px >= 0 && px < Wd && py >= 0 && py < Hd) {
                const float a = alpha[(size_t)py * Wd + px];
                if (a > 0.5f && (zc - depth[(size_t)py * Wd + px] / a) > thr) { u = nanf(""); v = nanf(""); }
            }
        }
    }
    kp2d[2 * k] = u; kp2d[2 * k + 1] = v;
}

__device__ __forceinline__ float blend8(float old, float col) { return rintf(0.4f * old + 0.6f * col); }      // cv2.addWeighted + saturate_cast<uchar>

__global__ void __launch_bounds__(256)
pose_draw_kernel(const float* __restrict__ kp2d, int H, int W, int flags, const uint8_t* __restrict__ hand_colors, float* __restrict__ out) {
    __shared__ float s_x[NK], s_y[NK];           // pixel coordinates (NaN = absent)
    __shared__ int s_ix[NK], s_iy[NK];           // int(x), int(y) as cv2 receives them; -1 = not drawn
    for (int k = threadIdx.x; k < NK; k += blockDim.x) {
        const float x = kp2d[2 * k], y = kp2d[2 * k + 1];
        const bool ok = !(isnan(x) || isnan(y));
        s_x[k] = x; s_y[k] = y;
        const float xn = x / (float)W * (float)W, yn = y / (float)H * (float)H;        // Keypoint.x * W (smpl_condition.py:35, open_pose.py:118)
        const int ix = ok ? (int)xn : -1, iy = ok ? (int)yn : -1;
        const bool draw = ok && ix > 0 && iy > 0;                                      // `x > eps and y > eps` on ints
        s_ix[k] = draw ? ix : -1; s_iy[k] = draw ? iy : -1;
    }
    __shared__ int s_hit;
    if (threadIdx.x == 0) s_hit = 0;
    __syncthreads();
    const int px = blockIdx.x * 16 + (threadIdx.x & 15), py = blockIdx.y * 16 + (threadIdx.x >> 4);
    // adaptive_draw_poses (open_pose.py:305-318)
    int body_r = 4, stick = 4, hand_r = 4, hand_th = 2, face_r = 3;
    if (H != 512 || W != 512) {
        const float r = (float)(H + W) / 2.f / 512.f;
        body_r = max((int)(body_r * r), 1); stick = max((int)(stick * r), 1); hand_r = max((int)(hand_r * r), 1);
        hand_th = max((int)(hand_th * r), 1); face_r = max((int)(face_r * r), 1);
    }
    {
        // Per-tile early out: the skeleton covers a few per cent of the image.  A tile draws nothing unless the bounding box of
        // some primitive (keypoint disc; limb ellipse and hand edge = box of their two end points), grown by the largest
        // radius / half-thickness + 2 pixels, meets it -- conservative, so the drawn image is unchanged.
        const float m = (float)(max(max(body_r, stick), max(max(hand_r, hand_th), face_r)) + 2);
        const float tx0 = (float)(blockIdx.x * 16) - m, tx1 = (float)(blockIdx.x * 16 + 15) + m;
        const float ty0 = (float)(blockIdx.y * 16) - m, ty1 = (float)(blockIdx.y * 16 + 15) + m;
        auto seg_hits = [&](int k1, int k2) {
            const float ax = s_x[k1], ay = s_y[k1], bx = s_x[k2], by = s_y[k2];
            if (isnan(ax) || isnan(ay) || isnan(bx) || isnan(by)) return false;
            return !(fmaxf(ax, bx) < tx0 || fminf(ax, bx) > tx1 || fmaxf(ay, by) < ty0 || fminf(ay, by) > ty1);
        };
        bool hit = false;
        const int t = threadIdx.x;
        if (t < NK) {
            const float x = s_x[t], y = s_y[t];
            hit = !(isnan(x) || isnan(y)) && x >= tx0 && x <= tx1 && y >= ty0 && y <= ty1;
        } else if (t < NK + 17) {
            const int limb[17][2] = {{2, 3}, {2, 6}, {3, 4}, {4, 5}, {6, 7}, {7, 8}, {2, 9}, {9, 10}, {10, 11}, {2, 12}, {12, 13}, {13, 14}, {2, 1},
                                     {1, 15}, {15, 17}, {1, 16}, {16, 18}};
            // both orientations (the flip map permutes body keypoints among themselves: test the limb's own and its mirror's box)
            const int l = t - NK;
            hit = seg_hits(limb[l][0] - 1, limb[l][1] - 1);
            const int flipmap[18] = {0, 1, 5, 6, 7, 2, 3, 4, 11, 12, 13, 8, 9, 10, 15, 14, 17, 16};
            hit |= seg_hits(flipmap[limb[l][0] - 1], flipmap[limb[l][1] - 1]);
        } else if (t < NK + 17 + 40) {
            const int edges[20][2] = {{0, 1}, {1, 2}, {2, 3}, {3, 4}, {0, 5}, {5, 6}, {6, 7}, {7, 8}, {0, 9}, {9, 10}, {10, 11}, {11, 12}, {0, 13},
                                      {13, 14}, {14, 15}, {15, 16}, {0, 17}, {17, 18}, {18, 19}, {19, 20}};
            const int e = (t - NK - 17) % 20, base = 18 + 21 * ((t - NK - 17) / 20);
            hit = seg_hits(base + edges[e][0], base + edges[e][1]);
        }
        if (hit) s_hit = 1;
        __syncthreads();
    }
    if (px >= W || py >= H) return;
    if (!s_hit) {
        const size_t HW0 = (size_t)H * W, pix0 = (size_t)py * W + px;
        out[pix0] = 0.f; out[HW0 + pix0] = 0.f; out[2 * HW0 + pix0] = 0.f;
        return;
    }
    float v0 = 0.f, v1 = 0.f, v2 = 0.f;
    const int body_col[18][3] = {{255, 0, 0}, {255, 85, 0}, {255, 170, 0}, {255, 255, 0}, {170, 255, 0}, {85, 255, 0}, {0, 255, 0}, {0, 255, 85},
                                 {0, 255, 170}, {0, 255, 255}, {0, 170, 255}, {0, 85, 255}, {0, 0, 255}, {85, 0, 255}, {170, 0, 255}, {255, 0, 255},
                                 {255, 0, 170}, {255, 0, 85}};
    if (flags & 1) {
        const int flipmap[18] = {0, 1, 5, 6, 7, 2, 3, 4, 11, 12, 13, 8, 9, 10, 15, 14, 17, 16};            // open_pose.py:92-101
        const bool flip = (flags & 8) != 0;
        for (int i = 0; i < 18; i++) {
            const int k = flip ? flipmap[i] : i;
            if (s_ix[k] < 0) continue;
            const int dx = px - s_ix[k], dy = py - s_iy[k];
            if (dx * dx + dy * dy <= body_r * body_r) { v0 = (float)body_col[i][0]; v1 = (float)body_col[i][1]; v2 = (float)body_col[i][2]; }
        }
        const int limb[17][2] = {{2, 3}, {2, 6}, {3, 4}, {4, 5}, {6, 7}, {7, 8}, {2, 9}, {9, 10}, {10, 11}, {2, 12}, {12, 13}, {13, 14}, {2, 1},
                                 {1, 15}, {15, 17}, {1, 16}, {16, 18}};
        for (int l = 0; l < 17; l++) {
            int k1 = limb[l][0] - 1, k2 = limb[l][1] - 1;
            if (flip) { k1 = flipmap[k1]; k2 = flipmap[k2]; }
            if (isnan(s_x[k1]) || isnan(s_x[k2])) continue;                              // keypoint is None (limbs do not apply the > eps test)
            const float Y0 = s_x[k1], Y1 = s_x[k2], X0 = s_y[k1], X1 = s_y[k2];          // reference naming: X = rows, Y = columns
            const float mX = 0.5f * (X0 + X1), mY = 0.5f * (Y0 + Y1);
            const float length = sqrtf((X0 - X1) * (X0 - X1) + (Y0 - Y1) * (Y0 - Y1));
            const float angle = atan2f(X0 - X1, Y0 - Y1) * 57.29577951308232f;
            const int cx = (int)mY, cy = (int)mX, a = (int)(length * 0.5f), ang = (int)angle;
            const float t = (float)ang * 0.017453292519943295f;
            float sn, cs;
            sincosf(t, &sn, &cs);
            const float dx = (float)(px - cx), dy = (float)(py - cy);
            const float u = dx * cs + dy * sn, w = -dx * sn + dy * cs;
            const float ea = (float)a + 0.5f, eb = (float)stick + 0.5f;
            if ((u * u) / (ea * ea) + (w * w) / (eb * eb) <= 1.0f) {
                v0 = blend8(v0, (float)body_col[l][0]); v1 = blend8(v1, (float)body_col[l][1]); v2 = blend8(v2, (float)body_col[l][2]);
            }
        }
    }
    if (flags & 2) {
        const int edges[20][2] = {{0, 1}, {1, 2}, {2, 3}, {3, 4}, {0, 5}, {5, 6}, {6, 7}, {7, 8}, {0, 9}, {9, 10}, {10, 11}, {11, 12}, {0, 13},
                                  {13, 14}, {14, 15}, {15, 16}, {0, 17}, {17, 18}, {18, 19}, {19, 20}};
        const float hw0 = hand_th <= 1 ? 0.5f : (float)((hand_th + 1) / 2);
        for (int hnd = 0; hnd < 2; hnd++) {
            const int base = 18 + 21 * hnd;
            bool any = false;
            for (int k = 0; k < 21; k++) any |= !isnan(s_x[base + k]);
            if (!any) continue;
            for (int k = 0; k < 21; k++) {
                if (s_ix[base + k] < 0) continue;
                const int dx = px - s_ix[base + k], dy = py - s_iy[base + k];
                if (dx * dx + dy * dy <= hand_r * hand_r) { v0 = 0.f; v1 = 0.f; v2 = 255.f; }
            }
            for (int e = 0; e < 20; e++) {
                const int k1 = base + edges[e][0], k2 = base + edges[e][1];
                if (isnan(s_x[k1]) || isnan(s_x[k2])) continue;
                const int x1 = (int)s_x[k1], y1 = (int)s_y[k1], x2 = (int)s_x[k2], y2 = (int)s_y[k2];
                if (!(x1 > 0 && y1 > 0 && x2 > 0 && y2 > 0)) continue;
                const float ex = (float)(x2 - x1), ey = (float)(y2 - y1);
                const float L2 = ex * ex + ey * ey;
                float tt = L2 > 0.f ? ((float)(px - x1) * ex + (float)(py - y1) * ey) / L2 : 0.f;
                tt = fminf(fmaxf(tt, 0.f), 1.f);
                const float qx = (float)(px - x1) - tt * ex, qy = (float)(py - y1) - tt * ey;
                const float mn = fminf(fabsf(ex), fabsf(ey)), mxv = fmaxf(fabsf(ex), fabsf(ey));
                const float hw = hw0 + (mxv > 0.f ? 0.5f * mn / mxv : 0.f);             // cv2's thick lines are wider along diagonals
                if (qx * qx + qy * qy <= hw * hw) { v0 = (float)hand_colors[3 * e]; v1 = (float)hand_colors[3 * e + 1]; v2 = (float)hand_colors[3 * e + 2]; }
            }
        }
    }
    if (flags & 4) {
        for (int k = 60; k < NK; k++) {
            if (s_ix[k] < 0) continue;
            const int dx = px - s_ix[k], dy = py - s_iy[k];
            if (dx * dx + dy * dy <= face_r * face_r) { v0 = 255.f; v1 = 255.f; v2 = 255.f; }
        }
    }
    const size_t HW = (size_t)H * W, pix = (size_t)py * W + px;
    out[pix] = v0 * (1.0f / 255.0f); out[HW + pix] = v1 * (1.0f / 255.0f); out[2 * HW + pix] = v2 * (1.0f / 255.0f);      // controlnet.py:45
}

}  // namespace pose
}  // namespace dwg

using namespace dwg;
using namespace dwg::pose;

extern "C" int dwg_pose_keypoints_2d(const float* kp_world, int K, const float* extrinsic_dev, const float* intrinsics_dev,
                                     float fx, float fy, float cx, float cy,
                                     const float* depth, const float* alpha, int Hd, int Wd, float cond_w, float cond_h,
                                     float thres_body, float thres_face, float thres_hand, float* kp2d, void* stream) {
    DWG_REQUIRE(kp_world && extrinsic_dev && kp2d && K > 0, "null pointer");
    DWG_REQUIRE((depth == nullptr) == (alpha == nullptr), "depth and alpha come together");
    pose_project_kernel<<<(K + 127) / 128, 128, 0, (cudaStream_t)stream>>>(kp_world, K, extrinsic_dev, intrinsics_dev, fx, fy, cx, cy, depth, alpha, Hd, Wd, cond_w, cond_h,
                                                                         thres_body, thres_face, thres_hand, kp2d);
    return check_launch("dwg_pose_keypoints_2d");
}

extern "C" int dwg_pose_image(const float* kp2d, int H, int W, int flags, const uint8_t* hand_edge_colors_dev, float* out, void* stream) {
    DWG_REQUIRE(kp2d && out && H > 0 && W > 0, "bad arguments");
    DWG_REQUIRE(!(flags & 2) || hand_edge_colors_dev, "hand drawing needs the 20 edge colours");
    pose_draw_kernel<<<dim3((W + 15) / 16, (H + 15) / 16), 256, 0, (cudaStream_t)stream>>>(kp2d, H, W, flags, hand_edge_colors_dev, out);
    return check_launch("dwg_pose_image");
}
