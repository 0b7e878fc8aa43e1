// Fused vertex pipeline: displacement map -> template vertices -> camera space, one launch
// forward and one backward instead of ~25 ATen launches in front of the rasteriser.
//
// Reference chain (under /root/reference/code):
//   MeshTemplate.get_vertex_positions   rendering/mesh_template.py:125-149
//       circpad / seam column (adjust_uv_and_texture :151-170) -> grid_sample (bilinear, align_corners) at the per-vertex
//       UV sites of the non-negative-x vertices -> deform: local (normal, tangent, bitangent) -> object space (:106-111)
//       -> mirror the +x half onto the -x half, pin the x = 0 vertices to the plane (:139-147) -> + template vertices
//   transform_vertices                  run_reconstruction.py:237-252
//       v' = qrot(q, (s + ds) v) + (t + dt);  (x, y, z) -> (x, -y, -z);  optional perspective x,y *= (z0 + z/2)/(z0 - z/2)
//   qrot                                rendering/utils.py:36-46     v + 2 (w (u x v) + u x (u x v))
// The sampling sites are constants of the template, so the host resolves them once per map size into four texel
// offsets of the UNPADDED map (the circular wrap / seam column folded in) and four bilinear weights per vertex.
#include "b3d_common.cuh"

namespace {
constexpr int NT = 128;

struct VRec {            // per template vertex, 80 bytes
    int tap[4];          // texel offsets y*w + x into one channel plane of the displacement map
    float wgt[4];        // bilinear weights (0 for taps that fall outside the padded map)
    float frame[9];      // (normal, tangent, bitangent) rows of the sampled (non-negative-x partner) vertex
    float v0[3];         // template vertex position
};
struct VSign { float sx, pad0, pad1, pad2; };   // x factor: -1 mirrored vertex, 0 vertex on the symmetry plane, +1 otherwise

struct Pose {
    const float* scale;      // [B]   s (+ ds added by the host when optimise_deltas)
    const float* trans;      // [B,3]
    const float* rot;        // [B,4] (w,x,y,z), used as given (unit quaternions in the reference's data)
    const float* z0;         // [B] or nullptr: perspective correction
    int apply;               // 0: raw vertices only
};

__device__ __forceinline__ void cross3(const float a[3], const float b[3], float o[3]) {
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}

// D [B,3,h,w] with element strides (sn, sc, sy*w folded into tap offsets by `sy`, `sx` strides): tap offset t = y*w + x
// is expanded as (t / w) * sy + (t % w) * sx.
__global__ void __launch_bounds__(NT)
vertex_fwd_kernel(const float* __restrict__ D, long long sn, long long sc, long long sy, long long sx, int w,
                  const VRec* __restrict__ rec, const VSign* __restrict__ sgn, int B, int V, Pose P,
                  float* __restrict__ raw, float* __restrict__ vtx) {
    const int i = blockIdx.x * NT + threadIdx.x;
    if (i >= B * V) return;
    const int b = i / V, v = i - b * V;
    const VRec r = rec[v];
    float local[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const float wt = r.wgt[t];
        if (wt != 0.f) {
            const long long off = (long long)b * sn + (long long)(r.tap[t] / w) * sy + (long long)(r.tap[t] % w) * sx;
#pragma unroll
            for (int k = 0; k < 3; ++k) local[k] = fmaf(wt, __ldg(D + off + k * sc), local[k]);
        }
    }
    float m[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) m[c] = local[0] * r.frame[c] + local[1] * r.frame[3 + c] + local[2] * r.frame[6 + c];
    m[0] *= sgn[v].sx;
    float p[3] = {r.v0[0] + m[0], r.v0[1] + m[1], r.v0[2] + m[2]};
    float* ro = raw + (long long)i * 3;
    ro[0] = p[0]; ro[1] = p[1]; ro[2] = p[2];
    if (!P.apply) return;
    const float s = P.scale[b];
    const float q0 = P.rot[b * 4], u[3] = {P.rot[b * 4 + 1], P.rot[b * 4 + 2], P.rot[b * 4 + 3]};
    float sp[3] = {s * p[0], s * p[1], s * p[2]}, t1[3], t2[3];
    cross3(u, sp, t1);
    cross3(u, t1, t2);
    float o[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) o[c] = sp[c] + 2.f * (q0 * t1[c] + t2[c]) + P.trans[b * 3 + c];
    o[1] = -o[1];
    o[2] = -o[2];
    if (P.z0) {
        const float z0 = P.z0[b];
        const float f = (z0 + o[2] * 0.5f) / (z0 - o[2] * 0.5f);
        o[0] *= f;
        o[1] *= f;
    }
    float* vo = vtx + (long long)i * 3;
    vo[0] = o[0]; vo[1] = o[1]; vo[2] = o[2];
}

// Backward.  g_raw (nullable) = gradient reaching the raw vertices directly (flat-loss path), g_vtx (nullable) through the
// pose.  Outputs: dD (atomics, caller zeroes), and per-sample d scale [B], d trans [B,3], d z0 [B] (atomics, caller zeroes,
// nullable).  One CTA handles NT vertices of ONE sample (grid.y = sample) so the pose gradients reduce per block.
__global__ void __launch_bounds__(NT)
vertex_bwd_kernel(const float* __restrict__ g_raw, const float* __restrict__ g_vtx, const float* __restrict__ raw,
                  const VRec* __restrict__ rec, const VSign* __restrict__ sgn, int B, int V, Pose P, int w, int hw,
                  float* __restrict__ dD, float* __restrict__ d_scale, float* __restrict__ d_trans, float* __restrict__ d_z0) {
    __shared__ float red[32];
    const int b = blockIdx.y, v = blockIdx.x * NT + threadIdx.x;
    const bool on = v < V;
    float g[3] = {0.f, 0.f, 0.f};           // gradient w.r.t. the raw vertex
    float ds = 0.f, dt[3] = {0.f, 0.f, 0.f}, dz0 = 0.f;
    const long long i = (long long)b * V + v;
    if (on && g_raw) { g[0] = g_raw[i * 3]; g[1] = g_raw[i * 3 + 1]; g[2] = g_raw[i * 3 + 2]; }
    if (on && g_vtx && P.apply) {
        float go[3] = {g_vtx[i * 3], g_vtx[i * 3 + 1], g_vtx[i * 3 + 2]};
        const float p[3] = {raw[i * 3], raw[i * 3 + 1], raw[i * 3 + 2]};
        const float s = P.scale[b];
        const float q0 = P.rot[b * 4], u[3] = {P.rot[b * 4 + 1], P.rot[b * 4 + 2], P.rot[b * 4 + 3]};
        if (P.z0) {                          // recompute the pre-perspective camera-space point
            float sp[3] = {s * p[0], s * p[1], s * p[2]}, t1[3], t2[3], o[3];
            cross3(u, sp, t1);
            cross3(u, t1, t2);
#pragma unroll
            for (int c = 0; c < 3; ++c) o[c] = sp[c] + 2.f * (q0 * t1[c] + t2[c]) + P.trans[b * 3 + c];
            o[1] = -o[1];
            o[2] = -o[2];
            const float z0 = P.z0[b], den = z0 - 0.5f * o[2], num = z0 + 0.5f * o[2];
            const float f = num / den;
            const float gf = go[0] * o[0] + go[1] * o[1];              // d f
            // f = num/den: df/dz = (0.5 den + 0.5 num)/den^2 = z0/den^2 ; df/dz0 = (den - num)/den^2 = -z/den^2
            go[2] += gf * z0 / (den * den);
            dz0 = gf * (-o[2]) / (den * den);
            go[0] *= f;
            go[1] *= f;
        }
        go[1] = -go[1];
        go[2] = -go[2];
        dt[0] = go[0]; dt[1] = go[1]; dt[2] = go[2];
        // rotation is linear: adjoint = rotation by the conjugate quaternion
        const float un[3] = {-u[0], -u[1], -u[2]};
        float t1[3], t2[3], gr[3];
        cross3(un, go, t1);
        cross3(un, t1, t2);
#pragma unroll
        for (int c = 0; c < 3; ++c) gr[c] = go[c] + 2.f * (q0 * t1[c] + t2[c]);
        ds = gr[0] * p[0] + gr[1] * p[1] + gr[2] * p[2];
#pragma unroll
        for (int c = 0; c < 3; ++c) g[c] += s * gr[c];
    }
    if (on && dD) {
        const VRec r = rec[v];
        g[0] *= sgn[v].sx;
        float gl[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) gl[k] = g[0] * r.frame[3 * k] + g[1] * r.frame[3 * k + 1] + g[2] * r.frame[3 * k + 2];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            if (r.wgt[t] != 0.f) {
#pragma unroll
                for (int k = 0; k < 3; ++k) atomicAdd(dD + ((long long)b * 3 + k) * hw + r.tap[t], r.wgt[t] * gl[k]);
            }
        }
    }
    if (d_scale || d_trans || d_z0) {
        float vals[5] = {ds, dt[0], dt[1], dt[2], dz0};
        float* dst[5] = {d_scale ? d_scale + b : nullptr, d_trans ? d_trans + b * 3 : nullptr, d_trans ? d_trans + b * 3 + 1 : nullptr,
                         d_trans ? d_trans + b * 3 + 2 : nullptr, (d_z0 && P.z0) ? d_z0 + b : nullptr};
#pragma unroll
        for (int j = 0; j < 5; ++j) {
            if (dst[j] == nullptr) continue;           // uniform across the block
            const float tot = b3d::block_sum(vals[j], red);
            if (threadIdx.x == 0) atomicAdd(dst[j], tot);
        }
    }
}
}  // namespace

extern "C" {

int b3d_vertex_record_bytes(void) { return (int)sizeof(VRec); }

// dmap [B,3,h,w] addressed with element strides (sn, sc, sy, sx) — NCHW or channels-last storage alike.
// rec [V] records of b3d_vertex_record_bytes() bytes (tap[4] int32, wgt[4], frame[9], v0[3]), sgn [V] float4 (x factor).
// scale [B] / trans [B,3] / rot [B,4] / z0 [B]: pose (all nullable together with vtx: raw vertices only).
// raw [B,V,3] out: MeshTemplate.get_vertex_positions; vtx [B,V,3] out: transform_vertices of it.
int b3d_vertex_pipeline_fwd(const float* dmap, long long sn, long long sc, long long sy, long long sx, int h, int w,
                            const void* rec, const void* sgn, int B, int V, const float* scale, const float* trans,
                            const float* rot, const float* z0, float* raw, float* vtx, void* stream) {
    B3D_REQUIRE(B >= 0 && V > 0 && h > 0 && w > 0, B3D_EINVAL, "b3d_vertex_pipeline_fwd: bad sizes");
    if (B == 0) return B3D_OK;
    B3D_REQUIRE(dmap && rec && sgn && raw, B3D_EINVAL, "b3d_vertex_pipeline_fwd: null pointer");
    B3D_REQUIRE((vtx != nullptr) == (scale && trans && rot), B3D_EINVAL, "b3d_vertex_pipeline_fwd: pose and vtx go together");
    Pose P{scale, trans, rot, z0, vtx != nullptr};
    vertex_fwd_kernel<<<b3d::ceil_div(B * V, NT), NT, 0, (cudaStream_t)stream>>>(dmap, sn, sc, sy, sx, w, static_cast<const VRec*>(rec),
                                                                               static_cast<const VSign*>(sgn), B, V, P, raw, vtx);
    B3D_LAUNCH_OK();
    return B3D_OK;
}

// g_raw / g_vtx [B,V,3] (either nullable).  d_dmap [B,3,h,w] contiguous NCHW (nullable), d_scale [B], d_trans [B,3],
// d_z0 [B] (nullable): all ACCUMULATED into — the caller zeroes them.
int b3d_vertex_pipeline_bwd(const float* g_raw, const float* g_vtx, const float* raw, const void* rec, const void* sgn, int B, int V,
                            int h, int w, const float* scale, const float* trans, const float* rot, const float* z0,
                            float* d_dmap, float* d_scale, float* d_trans, float* d_z0, void* stream) {
    B3D_REQUIRE(B >= 0 && V > 0 && h > 0 && w > 0, B3D_EINVAL, "b3d_vertex_pipeline_bwd: bad sizes");
    if (B == 0) return B3D_OK;
    B3D_REQUIRE(raw && rec && sgn && (g_raw || g_vtx), B3D_EINVAL, "b3d_vertex_pipeline_bwd: null pointer");
    B3D_REQUIRE(!g_vtx || (scale && trans && rot), B3D_EINVAL, "b3d_vertex_pipeline_bwd: g_vtx needs the pose");
    Pose P{scale, trans, rot, z0, g_vtx != nullptr};
    dim3 grid(b3d::ceil_div(V, NT), B);
    vertex_bwd_kernel<<<grid, NT, 0, (cudaStream_t)stream>>>(g_raw, g_vtx, raw, static_cast<const VRec*>(rec),
                                                           static_cast<const VSign*>(sgn), B, V, P, w, h * w, d_dmap, d_scale,
                                                           d_trans, d_z0);
    B3D_LAUNCH_OK();
    return B3D_OK;
}
}
