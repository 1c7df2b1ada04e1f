// R1 (GeneralLinearBlendSkinning.forward as the avatar uses it) and R5 (mesh-bound Gaussians) as kernels, so that
// DreamWaltzG.animate issues no eager torch ops for them.
//
// R1  core/human/inverse_lbs.py:570-784 + smplx.lbs.{batch_rodrigues, blend_shapes, vertices2joints, batch_rigid_transform}.
//     The reference evaluates the blend shapes for all 10 475 vertices (two GEMVs over 106 MB of shape / pose directions)
//     only to use (a) the 55 joint transforms and (b) the transforms of the few thousand PREDEFINED mesh vertices.
//     Here: (a) glbs_joints_kernel, ONE CTA: full pose assembly, Rodrigues, joints from a pre-multiplied regressor
//         JS = J_regressor . shapedirs  [165, 400]  (J = J_template + JS . shape: the regressor is linear),
//         the kinematic chain (parents precede children), relative transforms A_j, pose feature (R_1.. - I);
//     (b) glbs_vertices_kernel, one warp per predefined vertex over per-part, per-vertex contiguous slices of the shape /
//         pose directions (gathered once at construction): shape + pose offsets, blended 3x4, p' = M (p + so + po) + t.
// R5  core/system/avatar.py:1016-1079 + utils/mesh.py:34-97.  mesh_normals_kernel gathers face normals over a static
//     vertex -> triangle adjacency (no atomics, no memset); mesh_points_kernel evaluates position, tangent frame, scales and
//     quaternion of every mesh-bound Gaussian.  Its backward is the SAME code on forward-mode dual numbers: a point has
//     5 differentiable inputs (3 barycentric weights, 2 scale multipliers) and 9 outputs, so the 9x5 Jacobian is carried
//     along and contracted with the upstream gradient -- no hand-derived adjoint of the cross-product / quaternion chain.
#include "common.cuh"

namespace dwg {
namespace glbs {

constexpr int NJ = 55, NP = 165, NS = 400, NPF = 486;

struct JointArgs {
    const float* pose_part[7];     // global_orient[3], body[63], jaw[3], leye[3], reye[3], lhand[45], rhand[45]
    const float* pose_mean;        // [165]
    const float* betas;            // [nb]   (betas, already including any extra_betas)
    const float* expression;       // [ne]
    int nb, ne;
    const float* J_template;       // [55,3]
    const float* JS;               // [165, nb+ne]
    const int* parents;            // [55]
    const float* transl;           // [3] or null
    float* A;                      // [55,16]  relative rigid transforms (smplx A)
    float* A_t;                    // [55,16]  transl o A  (joint transforms used for the Gaussians)
    float* pose_feature;           // [486]
    float* shape_out;              // [nb+ne]  assembled shape vector (consumed by the vertex kernel)
    float* joints;                 // [55,3]   rest joints (debug / parity)
    float* posed_joints;           // [55,3]   posed joints incl. transl (smplx posed_joints; keypoint source of the condition producer)
};

__global__ void __launch_bounds__(256) glbs_joints_kernel(const JointArgs a) {
    __shared__ float s_pose[NP], s_shape[NS], s_R[NJ][9], s_J[NJ][3], s_chain[NJ][12];
    const int t = threadIdx.x;
    const int part_len[7] = {3, 63, 3, 3, 3, 45, 45};
    if (t < NP) {
        int off = 0, p = 0;
        while (t >= off + part_len[p]) { off += part_len[p]; p++; }
        s_pose[t] = a.pose_part[p][t - off] + a.pose_mean[t];
    }
    const int ns = a.nb + a.ne;
    for (int i = t; i < ns; i += blockDim.x) {
        const float v = i < a.nb ? a.betas[i] : a.expression[i - a.nb];
        s_shape[i] = v;
        a.shape_out[i] = v;
    }
    __syncthreads();
    if (t < NJ) {                                                // smplx.lbs.batch_rodrigues
        const float rx = s_pose[3 * t], ry = s_pose[3 * t + 1], rz = s_pose[3 * t + 2];
        const float ax = rx + 1e-8f, ay = ry + 1e-8f, az = rz + 1e-8f;
        const float angle = sqrtf(ax * ax + ay * ay + az * az);
        const float dx = rx / angle, dy = ry / angle, dz = rz / angle;
        float s, c;
        sincosf(angle, &s, &c);
        const float K[9] = {0.f, -dz, dy, dz, 0.f, -dx, -dy, dx, 0.f};
        float KK[9];
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) KK[3 * i + j] = K[3 * i] * K[j] + K[3 * i + 1] * K[3 + j] + K[3 * i + 2] * K[6 + j];
        for (int i = 0; i < 9; i++) s_R[t][i] = ((i % 4) == 0 ? 1.f : 0.f) + s * K[i] + (1.f - c) * KK[i];
    }
    if (t < NP) {                                                // J = J_template + JS . shape   (vertices2joints o blend_shapes)
        const float* row = a.JS + (size_t)t * ns;
        float acc = 0.f;
        for (int k = 0; k < ns; k++) acc = fmaf(row[k], s_shape[k], acc);
        (&s_J[0][0])[t] = a.J_template[t] + acc;
    }
    __syncthreads();
    for (int i = t; i < NPF; i += blockDim.x) {                  // pose_feature = (R[1:] - I).flatten()
        const int j = 1 + i / 9, e = i % 9;
        a.pose_feature[i] = s_R[j][e] - ((e % 4) == 0 ? 1.f : 0.f);
    }
    if (t < NP) a.joints[t] = (&s_J[0][0])[t];
    // kinematic chain: chain_j = chain_parent . [R_j | J_j - J_parent]; parents precede children (SMPL-X ordering)
    if (t < 12) {
        const int r = t / 4, c = t % 4;
        s_chain[0][t] = c < 3 ? s_R[0][3 * r + c] : s_J[0][r];
    }
    __syncwarp();
    if (t < 32) {
        for (int j = 1; j < NJ; j++) {
            const int p = a.parents[j];
            float v = 0.f;
            if (t < 12) {
                const int r = t / 4, c = t % 4;
                const float* P = s_chain[p];
                if (c < 3) v = P[4 * r] * s_R[j][c] + P[4 * r + 1] * s_R[j][3 + c] + P[4 * r + 2] * s_R[j][6 + c];
                else v = P[4 * r] * (s_J[j][0] - s_J[p][0]) + P[4 * r + 1] * (s_J[j][1] - s_J[p][1]) + P[4 * r + 2] * (s_J[j][2] - s_J[p][2]) + P[4 * r + 3];
            }
            __syncwarp();
            if (t < 12) s_chain[j][t] = v;
            __syncwarp();
        }
    }
    __syncthreads();
    if (t < NP) a.posed_joints[t] = s_chain[t / 3][4 * (t % 3) + 3] + (a.transl ? a.transl[t % 3] : 0.f);
    for (int i = t; i < NJ * 16; i += blockDim.x) {              // A_j = [R_chain | t_chain - R_chain J_j]
        const int j = i / 16, r = (i % 16) / 4, c = i % 4;
        float v;
        if (r == 3) v = c == 3 ? 1.f : 0.f;
        else if (c < 3) v = s_chain[j][4 * r + c];
        else v = s_chain[j][4 * r + 3] - (s_chain[j][4 * r] * s_J[j][0] + s_chain[j][4 * r + 1] * s_J[j][1] + s_chain[j][4 * r + 2] * s_J[j][2]);
        a.A[i] = v;
        a.A_t[i] = (r < 3 && c == 3 && a.transl) ? v + a.transl[r] : v;
    }
}

// one warp per predefined vertex.  sdirs [Vp,3,ns], pdirs [Vp,3,486], w [Vp,55], p [Vp,3] -> out [Vp,3]
__global__ void __launch_bounds__(256)
glbs_vertices_kernel(int Vp, int ns, const float* __restrict__ shape, const float* __restrict__ pose_feature, const float* __restrict__ A,
                     const float* __restrict__ transl, const float* __restrict__ sdirs, const float* __restrict__ pdirs,
                     const float* __restrict__ w, const float* __restrict__ p, float* __restrict__ out) {
    __shared__ float s_A[NJ * 12];
    for (int i = threadIdx.x; i < NJ * 12; i += blockDim.x) s_A[i] = A[(i / 12) * 16 + (i % 12)];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int v = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (v >= Vp) return;
    float off[3] = {0.f, 0.f, 0.f};
    const float* sd = sdirs + (size_t)v * 3 * ns;
    const float* pd = pdirs + (size_t)v * 3 * NPF;
    for (int k = lane; k < ns; k += 32) {
        const float s = shape[k];
        off[0] = fmaf(sd[k], s, off[0]); off[1] = fmaf(sd[ns + k], s, off[1]); off[2] = fmaf(sd[2 * ns + k], s, off[2]);
    }
    for (int k = lane; k < NPF; k += 32) {
        const float f = pose_feature[k];
        off[0] = fmaf(pd[k], f, off[0]); off[1] = fmaf(pd[NPF + k], f, off[1]); off[2] = fmaf(pd[2 * NPF + k], f, off[2]);
    }
    float M[12];
#pragma unroll
    for (int i = 0; i < 12; i++) M[i] = 0.f;
    for (int j = lane; j < NJ; j += 32) {
        const float wj = w[(size_t)v * NJ + j];
#pragma unroll
        for (int i = 0; i < 12; i++) M[i] = fmaf(wj, s_A[j * 12 + i], M[i]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int i = 0; i < 3; i++) off[i] += __shfl_xor_sync(0xffffffffu, off[i], o);
#pragma unroll
        for (int i = 0; i < 12; i++) M[i] += __shfl_xor_sync(0xffffffffu, M[i], o);
    }
    if (lane < 3) {
        const float q0 = p[3 * v] + off[0], q1 = p[3 * v + 1] + off[1], q2 = p[3 * v + 2] + off[2];
        float r = M[4 * lane] * q0 + M[4 * lane + 1] * q1 + M[4 * lane + 2] * q2 + M[4 * lane + 3];
        if (transl) r += transl[lane];
        out[3 * v + lane] = r;
    }
}

// ------------------------------------------------------------------------------------------ R5
__device__ __forceinline__ void cross3(const float* a, const float* b, float* o) {
    o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0];
}

// vertex normals of utils/mesh.py:34-97 (compute_normal): gather over the static vertex -> triangle adjacency
__global__ void __launch_bounds__(256)
mesh_normals_kernel(int Vp, const float* __restrict__ vc, const int* __restrict__ tri, const int* __restrict__ adj_ptr,
                    const int* __restrict__ adj_tri, float* __restrict__ vn) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= Vp) return;
    float acc[3] = {0.f, 0.f, 0.f};
    for (int e = adj_ptr[v]; e < adj_ptr[v + 1]; e++) {
        const int f = adj_tri[e];
        const int i0 = tri[3 * f], i1 = tri[3 * f + 1], i2 = tri[3 * f + 2];
        const float a[3] = {vc[3 * i1] - vc[3 * i0], vc[3 * i1 + 1] - vc[3 * i0 + 1], vc[3 * i1 + 2] - vc[3 * i0 + 2]};
        const float b[3] = {vc[3 * i2] - vc[3 * i0], vc[3 * i2 + 1] - vc[3 * i0 + 1], vc[3 * i2 + 2] - vc[3 * i0 + 2]};
        float n[3];
        cross3(a, b, n);
        const float inv = 1.0f / sqrtf(fmaxf(n[0] * n[0] + n[1] * n[1] + n[2] * n[2], 1e-20f));
        acc[0] += n[0] * inv; acc[1] += n[1] * inv; acc[2] += n[2] * inv;
    }
    float d = acc[0] * acc[0] + acc[1] * acc[1] + acc[2] * acc[2];
    if (!(d > 1e-20f)) { acc[0] = 0.f; acc[1] = 0.f; acc[2] = 1.f; d = 1.f; }
    const float inv = 1.0f / sqrtf(fmaxf(d, 1e-20f));
    vn[3 * v] = acc[0] * inv; vn[3 * v + 1] = acc[1] * inv; vn[3 * v + 2] = acc[2] * inv;
}

// forward-mode scalar: value + ND partial derivatives (ND = 0 is a plain float)
template <int ND>
struct Dual {
    float v;
    float d[ND > 0 ? ND : 1];
    __device__ Dual() {}
    __device__ Dual(float x) : v(x) { for (int i = 0; i < ND; i++) d[i] = 0.f; }
};
template <int ND> __device__ __forceinline__ Dual<ND> operator+(const Dual<ND>& a, const Dual<ND>& b) { Dual<ND> r; r.v = a.v + b.v; for (int i = 0; i < ND; i++) r.d[i] = a.d[i] + b.d[i]; return r; }
template <int ND> __device__ __forceinline__ Dual<ND> operator-(const Dual<ND>& a, const Dual<ND>& b) { Dual<ND> r; r.v = a.v - b.v; for (int i = 0; i < ND; i++) r.d[i] = a.d[i] - b.d[i]; return r; }
template <int ND> __device__ __forceinline__ Dual<ND> operator-(const Dual<ND>& a) { Dual<ND> r; r.v = -a.v; for (int i = 0; i < ND; i++) r.d[i] = -a.d[i]; return r; }
template <int ND> __device__ __forceinline__ Dual<ND> operator*(const Dual<ND>& a, const Dual<ND>& b) { Dual<ND> r; r.v = a.v * b.v; for (int i = 0; i < ND; i++) r.d[i] = a.d[i] * b.v + a.v * b.d[i]; return r; }
template <int ND> __device__ __forceinline__ Dual<ND> operator/(const Dual<ND>& a, const Dual<ND>& b) {
    Dual<ND> r; const float ib = 1.0f / b.v; r.v = a.v * ib;
    for (int i = 0; i < ND; i++) r.d[i] = (a.d[i] - r.v * b.d[i]) * ib;
    return r;
}
template <int ND> __device__ __forceinline__ Dual<ND> dsqrt(const Dual<ND>& a) { Dual<ND> r; r.v = sqrtf(a.v); const float k = 0.5f / r.v; for (int i = 0; i < ND; i++) r.d[i] = a.d[i] * k; return r; }
template <int ND> __device__ __forceinline__ Dual<ND> dabs(const Dual<ND>& a) { return a.v < 0.f ? -a : a; }      // torch.abs: sign(x) (0 at 0 is irrelevant here)
template <int ND> __device__ __forceinline__ Dual<ND> dclamp(const Dual<ND>& a, float lo, float hi) {
    if (a.v < lo) return Dual<ND>(lo);
    if (a.v > hi) return Dual<ND>(hi);
    return a;
}

// One mesh-bound Gaussian.  Inputs: raw barycentric weights b[3], scale multipliers sc1 / sc2, the triangle's vertex
// coordinates V[3][3] and vertex normals N[3][3] (constants).  Outputs: pos[3], s1, s2, quaternion q[4].
template <int ND>
__device__ __forceinline__ void mesh_point(const Dual<ND> b[3], const Dual<ND>& sc1, const Dual<ND>& sc2, const float V[3][3],
                                           const float N[3][3], float inv_n_per_tri, Dual<ND> pos[3], Dual<ND>& s1, Dual<ND>& s2,
                                           Dual<ND> q[4]) {
    typedef Dual<ND> D;
    const float eps = 1e-9f;
    const D bs = b[0] + b[1] + b[2];
    D pn[3];
    for (int c = 0; c < 3; c++) {
        pos[c] = (b[0] * D(V[0][c]) + b[1] * D(V[1][c]) + b[2] * D(V[2][c])) / bs;          // get_positions: normalised weights
        pn[c] = b[0] * D(N[0][c]) + b[1] * D(N[1][c]) + b[2] * D(N[2][c]);                  // point normal: RAW weights (avatar.py:1057)
    }
    auto nrm = [&](const D* v) { return dsqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); };
    D v0[3], v1[3], v2[3];
    { const D n = nrm(pn) + D(eps); for (int c = 0; c < 3; c++) v0[c] = pn[c] / n; }
    v1[0] = D(0.f); v1[1] = v0[2]; v1[2] = -v0[1];                                          // cross(v0, (1,0,0))
    { const D n = nrm(v1) + D(eps); for (int c = 0; c < 3; c++) v1[c] = v1[c] / n; }
    v2[0] = v0[1] * v1[2] - v0[2] * v1[1]; v2[1] = v0[2] * v1[0] - v0[0] * v1[2]; v2[2] = v0[0] * v1[1] - v0[1] * v1[0];
    { const D n = nrm(v2) + D(eps); for (int c = 0; c < 3; c++) v2[c] = v2[c] / n; }
    D a1(0.f), a2(0.f);
    for (int k = 0; k < 3; k++) {
        D d1(0.f), d2(0.f);
        for (int c = 0; c < 3; c++) { const D e = D(V[k][c]) - pos[c]; d1 = d1 + e * v1[c]; d2 = d2 + e * v2[c]; }
        a1 = a1 + dabs(d1); a2 = a2 + dabs(d2);
    }
    s1 = a1 * D(inv_n_per_tri) * dclamp(sc1, 0.5f, 2.0f);
    s2 = a2 * D(inv_n_per_tri) * dclamp(sc2, 0.5f, 2.0f);
    // R = [v0 v1 v2] (columns) with rows 1, 2 negated (avatar.py:1068); pytorch3d matrix_to_quaternion + standardize
    const float fl[3] = {1.f, -1.f, -1.f};
    D m[3][3];
    for (int r = 0; r < 3; r++) { m[r][0] = v0[r] * D(fl[r]); m[r][1] = v1[r] * D(fl[r]); m[r][2] = v2[r] * D(fl[r]); }
    D e[4] = {D(1.f) + m[0][0] + m[1][1] + m[2][2], D(1.f) + m[0][0] - m[1][1] - m[2][2], D(1.f) - m[0][0] + m[1][1] - m[2][2],
              D(1.f) - m[0][0] - m[1][1] + m[2][2]};
    D qa[4];
    int sel = 0;
    for (int i = 0; i < 4; i++) {
        if (e[i].v > 0.f) { D t = e[i]; if (t.v < 1e-38f) t = D(1e-38f); qa[i] = dsqrt(t); } else qa[i] = D(0.f);
        if (qa[i].v > qa[sel].v) sel = i;
    }
    D cand[4];
    if (sel == 0) { cand[0] = qa[0] * qa[0]; cand[1] = m[2][1] - m[1][2]; cand[2] = m[0][2] - m[2][0]; cand[3] = m[1][0] - m[0][1]; }
    else if (sel == 1) { cand[0] = m[2][1] - m[1][2]; cand[1] = qa[1] * qa[1]; cand[2] = m[1][0] + m[0][1]; cand[3] = m[0][2] + m[2][0]; }
    else if (sel == 2) { cand[0] = m[0][2] - m[2][0]; cand[1] = m[1][0] + m[0][1]; cand[2] = qa[2] * qa[2]; cand[3] = m[1][2] + m[2][1]; }
    else { cand[0] = m[1][0] - m[0][1]; cand[1] = m[2][0] + m[0][2]; cand[2] = m[2][1] + m[1][2]; cand[3] = qa[3] * qa[3]; }
    D den = qa[sel];
    if (den.v < 0.1f) den = D(0.1f);
    den = den * D(2.f);
    const bool neg = (cand[0] / den).v < 0.f;
    for (int i = 0; i < 4; i++) { q[i] = cand[i] / den; if (neg) q[i] = -q[i]; }
}

__device__ __forceinline__ void load_tri(int f, const int* __restrict__ tri, const float* __restrict__ vc, const float* __restrict__ vn,
                                         float V[3][3], float N[3][3]) {
    for (int k = 0; k < 3; k++) {
        const int i = tri[3 * f + k];
        for (int c = 0; c < 3; c++) { V[k][c] = vc[3 * i + c]; N[k][c] = vn[3 * i + c]; }
    }
}

__global__ void __launch_bounds__(128)
mesh_points_fwd_kernel(int P, int n_per_tri, const float* __restrict__ bary, const float* __restrict__ scales_param,
                       const float* __restrict__ vc, const float* __restrict__ vn, const int* __restrict__ tri,
                       float* __restrict__ pos, float* __restrict__ scales, float* __restrict__ quat) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const int f = i / n_per_tri;
    float V[3][3], N[3][3];
    load_tri(f, tri, vc, vn, V, N);
    typedef Dual<0> D;
    const D b[3] = {D(bary[3 * i]), D(bary[3 * i + 1]), D(bary[3 * i + 2])};
    D p[3], s1, s2, q[4];
    mesh_point<0>(b, D(scales_param[3 * i + 1]), D(scales_param[3 * i + 2]), V, N, 1.0f / (float)n_per_tri, p, s1, s2, q);
    for (int c = 0; c < 3; c++) pos[3 * i + c] = p[c].v;
    scales[3 * i] = 0.f; scales[3 * i + 1] = s1.v; scales[3 * i + 2] = s2.v;
    for (int c = 0; c < 4; c++) quat[4 * i + c] = q[c].v;
}

__global__ void __launch_bounds__(128)
mesh_points_bwd_kernel(int P, int n_per_tri, const float* __restrict__ bary, const float* __restrict__ scales_param,
                       const float* __restrict__ vc, const float* __restrict__ vn, const int* __restrict__ tri,
                       const float* __restrict__ g_pos, const float* __restrict__ g_scales, const float* __restrict__ g_quat,
                       float* __restrict__ g_bary, float* __restrict__ g_scales_param) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const int f = i / n_per_tri;
    float V[3][3], N[3][3];
    load_tri(f, tri, vc, vn, V, N);
    typedef Dual<5> D;
    D b[3] = {D(bary[3 * i]), D(bary[3 * i + 1]), D(bary[3 * i + 2])};
    D sc1(scales_param[3 * i + 1]), sc2(scales_param[3 * i + 2]);
    b[0].d[0] = 1.f; b[1].d[1] = 1.f; b[2].d[2] = 1.f; sc1.d[3] = 1.f; sc2.d[4] = 1.f;
    D p[3], s1, s2, q[4];
    mesh_point<5>(b, sc1, sc2, V, N, 1.0f / (float)n_per_tri, p, s1, s2, q);
    float g[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    for (int k = 0; k < 5; k++) {
        float acc = 0.f;
        if (g_pos) for (int c = 0; c < 3; c++) acc = fmaf(g_pos[3 * i + c], p[c].d[k], acc);
        if (g_scales) acc = fmaf(g_scales[3 * i + 1], s1.d[k], fmaf(g_scales[3 * i + 2], s2.d[k], acc));
        if (g_quat) for (int c = 0; c < 4; c++) acc = fmaf(g_quat[4 * i + c], q[c].d[k], acc);
        g[k] = acc;
    }
    g_bary[3 * i] = g[0]; g_bary[3 * i + 1] = g[1]; g_bary[3 * i + 2] = g[2];
    g_scales_param[3 * i] = 0.f; g_scales_param[3 * i + 1] = g[3]; g_scales_param[3 * i + 2] = g[4];
}

}  // namespace glbs
}  // namespace dwg

using namespace dwg;
using namespace dwg::glbs;

extern "C" int dwg_glbs_joints(const float* global_orient, const float* body_pose, const float* jaw_pose, const float* leye_pose,
                               const float* reye_pose, const float* left_hand_pose, const float* right_hand_pose, const float* pose_mean,
                               const float* betas, int n_betas, const float* expression, int n_expr,
                               const float* J_template, const float* JS, const int32_t* parents, const float* transl,
                               float* A, float* A_transl, float* pose_feature, float* shape_out, float* joints, float* posed_joints, void* stream) {
    DWG_REQUIRE(global_orient && body_pose && jaw_pose && leye_pose && reye_pose && left_hand_pose && right_hand_pose && pose_mean && betas &&
                expression && J_template && JS && parents && A && A_transl && pose_feature && shape_out && joints && posed_joints, "null pointer");
    DWG_REQUIRE(n_betas > 0 && n_expr >= 0 && n_betas + n_expr <= NS, "at most 400 shape components");
    JointArgs a;
    a.pose_part[0] = global_orient; a.pose_part[1] = body_pose; a.pose_part[2] = jaw_pose; a.pose_part[3] = leye_pose;
    a.pose_part[4] = reye_pose; a.pose_part[5] = left_hand_pose; a.pose_part[6] = right_hand_pose;
    a.pose_mean = pose_mean; a.betas = betas; a.expression = expression; a.nb = n_betas; a.ne = n_expr;
    a.J_template = J_template; a.JS = JS; a.parents = parents; a.transl = transl;
    a.A = A; a.A_t = A_transl; a.pose_feature = pose_feature; a.shape_out = shape_out; a.joints = joints; a.posed_joints = posed_joints;
    glbs_joints_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(a);
    return check_launch("dwg_glbs_joints");
}

extern "C" int dwg_glbs_vertices(int Vp, int n_shape, const float* shape, const float* pose_feature, const float* A, const float* transl,
                                 const float* shapedirs_sel, const float* posedirs_sel, const float* weights_sel, const float* points,
                                 float* out, void* stream) {
    DWG_REQUIRE(shape && pose_feature && A && shapedirs_sel && posedirs_sel && weights_sel && points && out, "null pointer");
    DWG_REQUIRE(Vp >= 0 && n_shape > 0 && n_shape <= NS, "bad sizes");
    if (Vp == 0) return DWG_OK;
    glbs_vertices_kernel<<<(Vp + 7) / 8, 256, 0, (cudaStream_t)stream>>>(Vp, n_shape, shape, pose_feature, A, transl, shapedirs_sel, posedirs_sel,
                                                                       weights_sel, points, out);
    return check_launch("dwg_glbs_vertices");
}

extern "C" int dwg_mesh_gaussians_fwd(int Vp, int F, int n_per_tri, const float* vertex_coords, const int32_t* triangles,
                                      const int32_t* adj_ptr, const int32_t* adj_tri, const float* bary, const float* scales_param,
                                      float* vertex_normals, float* positions, float* scales, float* quaternions, void* stream) {
    DWG_REQUIRE(vertex_coords && triangles && adj_ptr && adj_tri && bary && scales_param && vertex_normals && positions && scales && quaternions,
                "null pointer");
    DWG_REQUIRE(Vp > 0 && F > 0 && n_per_tri > 0, "bad sizes");
    cudaStream_t st = (cudaStream_t)stream;
    mesh_normals_kernel<<<(Vp + 255) / 256, 256, 0, st>>>(Vp, vertex_coords, triangles, adj_ptr, adj_tri, vertex_normals);
    const int P = F * n_per_tri;
    mesh_points_fwd_kernel<<<(P + 127) / 128, 128, 0, st>>>(P, n_per_tri, bary, scales_param, vertex_coords, vertex_normals, triangles, positions,
                                                          scales, quaternions);
    return check_launch("dwg_mesh_gaussians_fwd");
}

extern "C" int dwg_mesh_gaussians_bwd(int F, int n_per_tri, const float* vertex_coords, const float* vertex_normals, const int32_t* triangles,
                                      const float* bary, const float* scales_param, const float* g_positions, const float* g_scales,
                                      const float* g_quaternions, float* g_bary, float* g_scales_param, void* stream) {
    DWG_REQUIRE(vertex_coords && vertex_normals && triangles && bary && scales_param && g_bary && g_scales_param, "null pointer");
    const int P = F * n_per_tri;
    mesh_points_bwd_kernel<<<(P + 127) / 128, 128, 0, (cudaStream_t)stream>>>(P, n_per_tri, bary, scales_param, vertex_coords, vertex_normals, triangles,
                                                                            g_positions, g_scales, g_quaternions, g_bary, g_scales_param);
    return check_launch("dwg_mesh_gaussians_bwd");
}
