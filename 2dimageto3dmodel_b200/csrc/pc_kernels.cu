// Point-cloud "effective loss" path for sm_100a.
//
//   pc_project_kernel        a1+a2+a3(index part): quaternion rotate, perspective divide, grid coords,
//                            in-bounds mask, floor index buffer.  Uses round-to-nearest intrinsics in the
//                            reference's operation order (no FMA contraction) so that floor(g) — the
//                            reference's int64 index buffer — is reproduced bit for bit.
//   pc_sil_fwd_kernel        a3+a4+a5 fused, mode R: one CTA owns a TY x 16 patch of (y,x) columns over
//                            the whole depth in shared memory; points are splatted with shared-memory
//                            atomics (no global atomics, the V^3 grid never touches HBM), then each thread
//                            walks one column: clamp -> 21-tap z blur (register-blocked, 8 outputs per
//                            window) -> scale/clamp -> ray termination -> silhouette.
//   pc_sil_bwd_kernel        recomputes the patch (+1 halo so every point is owned by exactly one CTA),
//                            runs the ray-march / clamp / blur adjoints column-wise in shared memory and
//                            gathers d/d(grid coords) per point: no dense gradient grid in HBM either.
//   pc_project_bwd_kernel    adjoint of projection + rotation + quaternion normalisation.
//   pc_splat_grid_kernel     materialised occupancy grid (parity tests, mode P).
//
// Reference lines restated: quaternions/points_quaternions.py:41-81, quaternions/operations.py:68-136,
// camera/coordinate_system_transformation.py:20-39, utils/trilinear_interpolation.py:17-74,
// utils/smooth_voxels.py:44-84, utils/effective_loss_function.py:18-81 (all under /root/reference/code).
#include "b3d_common.cuh"

namespace {

using b3d::clamp_nan;

constexpr int TX = 16;         // patch width in x (columns are contiguous over x in shared memory)
constexpr int ZB = 8;          // blur outputs produced per register window
constexpr int MAX_TAPS = 63;
constexpr int NTHREADS = 256;
// field_of_view = 1.875 and camera_view_distance = 2.0 are what effective_loss_function.py:69-70 passes
constexpr float TERM_EPS = 1e-5f;      // effective_loss_function.py:18

struct Taps {
    float w[MAX_TAPS + 1];
    int n;
};

// ----------------------------------------------------------------------------------------------
// exact-order fp32 helpers (mirror torch's separate rounding of every elementwise op)
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }

struct Quat {
    float w, x, y, z;
};

// operations.py:83-86, evaluated left to right like the Python expressions
__device__ __forceinline__ Quat hamilton_exact(const Quat a, const Quat b) {
    Quat r;
    r.w = sub(sub(sub(mul(a.w, b.w), mul(a.x, b.x)), mul(a.y, b.y)), mul(a.z, b.z));
    r.x = sub(add(add(mul(a.w, b.x), mul(a.x, b.w)), mul(a.y, b.z)), mul(a.z, b.y));
    r.y = sub(add(add(mul(a.w, b.y), mul(a.y, b.w)), mul(a.z, b.x)), mul(a.x, b.z));
    r.z = sub(add(add(mul(a.w, b.z), mul(a.z, b.w)), mul(a.x, b.y)), mul(a.y, b.x));
    return r;
}

__device__ __forceinline__ Quat hamilton(const Quat a, const Quat b) {
    Quat r;
    r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
    r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
    r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
    r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
    return r;
}
__device__ __forceinline__ Quat conj(const Quat a) { return Quat{a.w, -a.x, -a.y, -a.z}; }

// F.normalize(q, dim=-1): q / max(||q||_2, 1e-12)   (points_quaternions.py:53-56)
__device__ __forceinline__ Quat normalize_quat(const float* q, float* norm_out) {
    const float w = q[0], x = q[1], y = q[2], z = q[3];
    float n = __fsqrt_rn(add(add(add(mul(w, w), mul(x, x)), mul(y, y)), mul(z, z)));
    n = n < 1e-12f ? 1e-12f : n;
    if (norm_out) *norm_out = n;
    return Quat{__fdiv_rn(w, n), __fdiv_rn(x, n), __fdiv_rn(y, n), __fdiv_rn(z, n)};
}

// ----------------------------------------------------------------------------------------------
// projection
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NTHREADS)
pc_project_kernel(const float* __restrict__ points, const float* __restrict__ quat, int N, int V,
                  float FOV, float CAM_DIST, float4* __restrict__ pg, float* __restrict__ coords, int32_t* __restrict__ base,
                  uint8_t* __restrict__ inb) {
    const int b = blockIdx.y;
    const int n = blockIdx.x * NTHREADS + threadIdx.x;
    if (n >= N) return;
    const Quat q = normalize_quat(quat + 4 * b, nullptr);
    const size_t i = (size_t)b * N + n;
    const float p0 = points[3 * i + 0], p1 = points[3 * i + 1], p2 = points[3 * i + 2];
    const Quat P{0.f, p0, p1, p2};                       // points_quaternions.py:33 (pad scalar 0)
    const Quat r = hamilton_exact(hamilton_exact(q, P), conj(q));   // :72-75 (forward direction)
    // coordinate_system_transformation.py:25-39: columns are (z, y, x)
    const float z = r.x;
    const float den = add(z, CAM_DIST);
    const float y = __fdiv_rn(mul(r.y, FOV), den);
    const float x = __fdiv_rn(mul(r.z, FOV), den);
    // trilinear_interpolation.py:23-25, scalars cast to fp32 as torch does
    const float hi = (float)(0.5 - 1e-6), lo = (float)(-0.5 + 1e-6);
    const bool ok = (z < hi) && (z > lo) && (y < hi) && (y > lo) && (x < hi) && (x > lo);
    // trilinear_interpolation.py:34
    const float vm1 = (float)(V - 1);
    const float gz = mul(vm1, add(z, 0.5f)), gy = mul(vm1, add(y, 0.5f)), gx = mul(vm1, add(x, 0.5f));
    pg[i] = make_float4(gz, gy, gx, ok ? 1.f : 0.f);
    if (coords) {
        coords[3 * i + 0] = z;
        coords[3 * i + 1] = y;
        coords[3 * i + 2] = x;
    }
    if (base) {
        base[3 * i + 0] = (int32_t)floorf(gz);
        base[3 * i + 1] = (int32_t)floorf(gy);
        base[3 * i + 2] = (int32_t)floorf(gx);
    }
    if (inb) inb[i] = ok ? 1 : 0;
}

// trilinear weights of one axis: trilinear_interpolation.py:66 (mode R keeps `1.0 - grid - floor`)
__device__ __forceinline__ void axis_weights(float g, float f, int mode, float& w0, float& w1) {
    w1 = sub(g, f);
    w0 = (mode == B3D_MODE_REFERENCE) ? sub(sub(1.0f, g), f) : sub(1.0f, w1);
}

// ----------------------------------------------------------------------------------------------
// register-blocked 1-D blur along z of one shared-memory column (stride cs between depths).
// out[j] = sum_k taps[k] * in[zb + j + k - KT/2], zero padding (smooth_voxels.py:69-73).
// ----------------------------------------------------------------------------------------------
template <int KT, bool CLAMP_IN, bool REVERSED>
__device__ __forceinline__ void blur_window(const float* colp, int cs, int V, int zb, const Taps& taps,
                                            float (&out)[ZB]) {
    if constexpr (KT > 0) {
        constexpr int H = KT / 2, W = ZB + KT - 1;
        float in[W];
#pragma unroll
        for (int i = 0; i < W; ++i) {
            const int z = zb - H + i;
            float v = (z >= 0 && z < V) ? colp[z * cs] : 0.f;
            if (CLAMP_IN) v = clamp_nan(v, 0.f, 1.f);        // trilinear_interpolation.py:74
            in[i] = v;
        }
#pragma unroll
        for (int j = 0; j < ZB; ++j) {
            float s = 0.f;
#pragma unroll
            for (int k = 0; k < KT; ++k) s = fmaf(taps.w[REVERSED ? KT - 1 - k : k], in[j + k], s);
            out[j] = s;
        }
    } else {
        const int kt = taps.n, H = kt / 2;
#pragma unroll
        for (int j = 0; j < ZB; ++j) out[j] = 0.f;
        for (int k = 0; k < kt; ++k) {
            const float t = taps.w[REVERSED ? kt - 1 - k : k];
#pragma unroll
            for (int j = 0; j < ZB; ++j) {
                const int z = zb + j + k - H;
                float v = (z >= 0 && z < V) ? colp[z * cs] : 0.f;
                if (CLAMP_IN) v = clamp_nan(v, 0.f, 1.f);
                out[j] = fmaf(t, v, out[j]);
            }
        }
    }
}

// occupancy after smooth_voxels.py:80-82 and effective_loss_function.py:33
__device__ __forceinline__ float scaled(float S, bool has_scale, float sc) {
    return has_scale ? clamp_nan(S * sc, 0.f, 1.f) : S;
}

// ----------------------------------------------------------------------------------------------
// forward: splat into a shared-memory patch, then the column walk
// ----------------------------------------------------------------------------------------------
template <int KT>
__global__ void __launch_bounds__(NTHREADS)
pc_sil_fwd_kernel(const float4* __restrict__ pg, const Taps taps, const float* __restrict__ scale, int N,
                  int V, int TY, int mode, float* __restrict__ sil) {
    extern __shared__ float sm[];
    const int b = blockIdx.z, ty0 = blockIdx.y * TY, tx0 = blockIdx.x * TX;
    const int ncol = TY * TX;
    const int tid = threadIdx.x;
    for (int i = tid; i < V * ncol; i += NTHREADS) sm[i] = 0.f;
    __syncthreads();

    const float4* p = pg + (size_t)b * N;
    for (int n0 = 0; n0 < N; n0 += 4 * NTHREADS) {
        float4 g[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int n = n0 + u * NTHREADS + tid;
            g[u] = n < N ? __ldg(p + n) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (g[u].w == 0.f) continue;
            const float fzf = floorf(g[u].x), fyf = floorf(g[u].y), fxf = floorf(g[u].z);
            const int ly = (int)fyf - ty0, lx = (int)fxf - tx0;
            if (ly < -1 || ly >= TY || lx < -1 || lx >= TX) continue;
            const int fz = (int)fzf;
            float wz[2], wy[2], wx[2];
            axis_weights(g[u].x, fzf, mode, wz[0], wz[1]);
            axis_weights(g[u].y, fyf, mode, wy[0], wy[1]);
            axis_weights(g[u].z, fxf, mode, wx[0], wx[1]);
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int cy = ly + j;
                if (cy < 0 || cy >= TY) continue;
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const int cx = lx + k;
                    if (cx < 0 || cx >= TX) continue;
#pragma unroll
                    for (int i = 0; i < 2; ++i)     // trilinear_interpolation.py:40-41: (gz_i*gy_j)*gx_k
                        atomicAdd(&sm[(fz + i) * ncol + cy * TX + cx], mul(mul(wz[i], wy[j]), wx[k]));
                }
            }
        }
    }
    __syncthreads();

    const bool has_scale = scale != nullptr;
    const float sc = has_scale ? scale[b] : 1.f;
    const float c0 = (mode == B3D_MODE_REFERENCE) ? expf(TERM_EPS) : 1.f;   // D10 pad row
    for (int col = tid; col < ncol; col += NTHREADS) {
        const int y = ty0 + col / TX, x = tx0 + col % TX;
        if (y >= V || x >= V) continue;
        const float* colp = sm + col;
        float T = 1.f, acc = 0.f;
        for (int zb = 0; zb < V; zb += ZB) {
            float S[ZB];
            blur_window<KT, true, false>(colp, ncol, V, zb, taps, S);
#pragma unroll
            for (int j = 0; j < ZB; ++j) {
                const int z = zb + j;
                if (z >= V) break;
                const float o = clamp_nan(scaled(S[j], has_scale, sc), TERM_EPS, 1.f - TERM_EPS);
                float term = o * T;                       // o_k * prod_{j<k}(1-o_j)
                if (z == 0) term *= c0;
                acc += term;
                T *= (1.f - o);
            }
        }
        sil[((size_t)b * V + (V - 1 - y)) * V + x] = acc;   // flip(1): effective_loss_function.py:81
    }
}

// ----------------------------------------------------------------------------------------------
// backward
// ----------------------------------------------------------------------------------------------
template <int KT>
__global__ void __launch_bounds__(NTHREADS)
pc_sil_bwd_kernel(const float4* __restrict__ pg, const Taps taps, const float* __restrict__ scale,
                  const float* __restrict__ dsil, int N, int V, int TY, int mode,
                  float4* __restrict__ dpg, float* __restrict__ dscale) {
    extern __shared__ float sm[];
    __shared__ float red[32];
    const int b = blockIdx.z, ty0 = blockIdx.y * TY, tx0 = blockIdx.x * TX;
    const int EY = TY + 1, EX = TX + 1, ncol = EY * EX;
    float* A1 = sm;                 // raw sums G -> (+-)T_z -> dG
    float* A2 = sm + V * ncol;      // blurred S -> dS
    const int tid = threadIdx.x;
    for (int i = tid; i < V * ncol; i += NTHREADS) A1[i] = 0.f;
    __syncthreads();

    // splat every point touching the extended patch [ty0, ty0+TY] x [tx0, tx0+TX]
    const float4* p = pg + (size_t)b * N;
    for (int n0 = 0; n0 < N; n0 += 4 * NTHREADS) {
        float4 g[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int n = n0 + u * NTHREADS + tid;
            g[u] = n < N ? __ldg(p + n) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (g[u].w == 0.f) continue;
            const float fzf = floorf(g[u].x), fyf = floorf(g[u].y), fxf = floorf(g[u].z);
            const int ly = (int)fyf - ty0, lx = (int)fxf - tx0;
            if (ly < -1 || ly > TY || lx < -1 || lx > TX) continue;
            const int fz = (int)fzf;
            float wz[2], wy[2], wx[2];
            axis_weights(g[u].x, fzf, mode, wz[0], wz[1]);
            axis_weights(g[u].y, fyf, mode, wy[0], wy[1]);
            axis_weights(g[u].z, fxf, mode, wx[0], wx[1]);
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int cy = ly + j;
                if (cy < 0 || cy > TY) continue;
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const int cx = lx + k;
                    if (cx < 0 || cx > TX) continue;
#pragma unroll
                    for (int i = 0; i < 2; ++i)
                        atomicAdd(&A1[(fz + i) * ncol + cy * EX + cx], mul(mul(wz[i], wy[j]), wx[k]));
                }
            }
        }
    }
    __syncthreads();

    const bool has_scale = scale != nullptr;
    const float sc = has_scale ? scale[b] : 1.f;
    const float c0 = (mode == B3D_MODE_REFERENCE) ? expf(TERM_EPS) : 1.f;
    float dsc = 0.f;
    for (int col = tid; col < ncol; col += NTHREADS) {
        const int cy = col / EX, cx = col % EX;
        const int y = ty0 + cy, x = tx0 + cx;
        if (y >= V || x >= V) continue;
        const bool owned = cy < TY && cx < TX;      // halo columns are owned by the neighbour patch
        float* c1 = A1 + col;
        float* c2 = A2 + col;
        // (a) S = blur_z(clamp(G, 0, 1))
        for (int zb = 0; zb < V; zb += ZB) {
            float S[ZB];
            blur_window<KT, true, false>(c1, ncol, V, zb, taps, S);
#pragma unroll
            for (int j = 0; j < ZB; ++j)
                if (zb + j < V) c2[(zb + j) * ncol] = S[j];
        }
        // (b) prefix transmittance T_z; its sign bit keeps the clamp mask 0 <= G <= 1
        float T = 1.f;
        for (int z = 0; z < V; ++z) {
            const float G = c1[z * ncol];
            const float o = clamp_nan(scaled(c2[z * ncol], has_scale, sc), TERM_EPS, 1.f - TERM_EPS);
            c1[z * ncol] = (G >= 0.f && G <= 1.f) ? T : -T;
            T *= (1.f - o);
        }
        // (c) suffix silhouette Q_{z+1}; d sil / d o_z = T_z (1 - Q_{z+1})   (z = 0: c0 - Q_1)
        const float go = dsil[((size_t)b * V + (V - 1 - y)) * V + x];
        float Q = 0.f;
        for (int z = V - 1; z >= 0; --z) {
            const float S = c2[z * ncol];
            const float t = has_scale ? S * sc : S;
            const float s2 = has_scale ? clamp_nan(t, 0.f, 1.f) : S;
            const float o = clamp_nan(s2, TERM_EPS, 1.f - TERM_EPS);
            const float Tz = fabsf(c1[z * ncol]);
            const float coef = (z == 0) ? (c0 - Q) : Tz * (1.f - Q);
            float d = go * coef;
            d = (s2 >= TERM_EPS && s2 <= 1.f - TERM_EPS) ? d : 0.f;      // clamp(eps, 1-eps) adjoint
            if (has_scale) {
                d = (t >= 0.f && t <= 1.f) ? d : 0.f;                    // clamp(0, 1) adjoint
                if (owned) dsc = fmaf(d, S, dsc);
                d *= sc;
            }
            c2[z * ncol] = d;
            Q = fmaf(1.f - o, Q, o);
        }
        // (d) dG = mask * blur_z^T(dS)
        for (int zb = 0; zb < V; zb += ZB) {
            float D[ZB];
            blur_window<KT, false, true>(c2, ncol, V, zb, taps, D);
#pragma unroll
            for (int j = 0; j < ZB; ++j) {
                const int z = zb + j;
                if (z < V) c1[z * ncol] = signbit(c1[z * ncol]) ? 0.f : D[j];
            }
        }
    }
    if (dscale) {
        const float tot = b3d::block_sum(dsc, red);
        if (tid == 0 && tot != 0.f) atomicAdd(dscale + b, tot);
    }
    __syncthreads();

    // gather: every in-bounds point is owned by the patch holding its base cell
    float4* dp = dpg + (size_t)b * N;
    for (int n = tid; n < N; n += NTHREADS) {
        const float4 g = __ldg(p + n);
        if (g.w == 0.f) continue;
        const float fzf = floorf(g.x), fyf = floorf(g.y), fxf = floorf(g.z);
        const int ly = (int)fyf - ty0, lx = (int)fxf - tx0;
        if (ly < 0 || ly >= TY || lx < 0 || lx >= TX) continue;
        const int fz = (int)fzf;
        float wz[2], wy[2], wx[2];
        axis_weights(g.x, fzf, mode, wz[0], wz[1]);
        axis_weights(g.y, fyf, mode, wy[0], wy[1]);
        axis_weights(g.z, fxf, mode, wx[0], wx[1]);
        float dz = 0.f, dy = 0.f, dx = 0.f;
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const float d = A1[(fz + i) * ncol + (ly + j) * EX + lx + k];
                    // d w0/dg = -1, d w1/dg = +1 in both modes (floor has zero gradient)
                    dz += d * (i ? 1.f : -1.f) * wy[j] * wx[k];
                    dy += d * wz[i] * (j ? 1.f : -1.f) * wx[k];
                    dx += d * wz[i] * wy[j] * (k ? 1.f : -1.f);
                }
        dp[n] = make_float4(dz, dy, dx, 0.f);
    }
}

// ----------------------------------------------------------------------------------------------
// adjoint of pc_project_kernel
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NTHREADS)
pc_project_bwd_kernel(const float* __restrict__ points, const float* __restrict__ quat,
                      const float4* __restrict__ pg, const float4* __restrict__ dpg, int N, int V,
                      float FOV, float CAM_DIST, float* __restrict__ dpoints, float* __restrict__ dquat) {
    __shared__ float red[32];
    const int b = blockIdx.y;
    const int n = blockIdx.x * NTHREADS + threadIdx.x;
    float nrm;
    const Quat q = normalize_quat(quat + 4 * b, &nrm);
    Quat dq{0.f, 0.f, 0.f, 0.f};
    if (n < N) {
        const size_t i = (size_t)b * N + n;
        float d0 = 0.f, d1 = 0.f, d2 = 0.f;
        if (pg[i].w != 0.f) {
            const float4 dg = dpg[i];
            const float vm1 = (float)(V - 1);
            const float dc0 = vm1 * dg.x, dc1 = vm1 * dg.y, dc2 = vm1 * dg.z;
            const Quat P{0.f, points[3 * i], points[3 * i + 1], points[3 * i + 2]};
            const Quat u = hamilton(q, P);
            const Quat r = hamilton(u, conj(q));
            const float inv = 1.f / (r.x + CAM_DIST);
            const float c1 = r.y * FOV * inv, c2 = r.z * FOV * inv;
            const Quat Gr{0.f, dc0 - (c1 * dc1 + c2 * dc2) * inv, dc1 * FOV * inv, dc2 * FOV * inv};
            const Quat gu = hamilton(Gr, q);               // d/d(q (x) P)
            const Quat gw = hamilton(conj(u), Gr);         // d/d(q*)
            const Quat a = hamilton(gu, conj(P));
            dq = Quat{a.w + gw.w, a.x - gw.x, a.y - gw.y, a.z - gw.z};
            const Quat dP = hamilton(conj(q), gu);
            d0 = dP.x;
            d1 = dP.y;
            d2 = dP.z;
        }
        dpoints[3 * i + 0] = d0;
        dpoints[3 * i + 1] = d1;
        dpoints[3 * i + 2] = d2;
    }
    const float sw = b3d::block_sum(dq.w, red);
    const float sx = b3d::block_sum(dq.x, red);
    const float sy = b3d::block_sum(dq.y, red);
    const float sz = b3d::block_sum(dq.z, red);
    if (threadIdx.x == 0) {
        // adjoint of q / max(||q||, eps): (I - q^ q^T) / n  (identity / eps in the degenerate branch)
        const float dot = (nrm > 1e-12f) ? (q.w * sw + q.x * sx + q.y * sy + q.z * sz) : 0.f;
        atomicAdd(dquat + 4 * b + 0, (sw - q.w * dot) / nrm);
        atomicAdd(dquat + 4 * b + 1, (sx - q.x * dot) / nrm);
        atomicAdd(dquat + 4 * b + 2, (sy - q.y * dot) / nrm);
        atomicAdd(dquat + 4 * b + 3, (sz - q.z * dot) / nrm);
    }
}

// ----------------------------------------------------------------------------------------------
// materialised grid (tests / mode P)
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NTHREADS)
pc_splat_grid_kernel(const float4* __restrict__ pg, int N, int V, int mode, float* __restrict__ grid) {
    const int b = blockIdx.y;
    const int n = blockIdx.x * NTHREADS + threadIdx.x;
    if (n >= N) return;
    const float4 g = pg[(size_t)b * N + n];
    if (g.w == 0.f) return;
    const float fzf = floorf(g.x), fyf = floorf(g.y), fxf = floorf(g.z);
    float wz[2], wy[2], wx[2];
    axis_weights(g.x, fzf, mode, wz[0], wz[1]);
    axis_weights(g.y, fyf, mode, wy[0], wy[1]);
    axis_weights(g.z, fxf, mode, wx[0], wx[1]);
    float* gb = grid + (size_t)b * V * V * V;
    const int fz = (int)fzf, fy = (int)fyf, fx = (int)fxf;
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int k = 0; k < 2; ++k)
                atomicAdd(gb + ((size_t)(fz + i) * V + fy + j) * V + fx + k, mul(mul(wz[i], wy[j]), wx[k]));
}

__global__ void __launch_bounds__(NTHREADS) clamp01_kernel(float* __restrict__ x, size_t n) {
    for (size_t i = (size_t)blockIdx.x * NTHREADS + threadIdx.x; i < n; i += (size_t)gridDim.x * NTHREADS)
        x[i] = clamp_nan(x[i], 0.f, 1.f);
}

// ----------------------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------------------
constexpr size_t SMEM_BUDGET = 200 * 1024;

int pick_ty(int V, bool bwd) {
    if (const char* e = getenv(bwd ? "B3D_PC_TY_BWD" : "B3D_PC_TY_FWD")) {
        const int v = atoi(e);
        if (v >= 1 && v <= 64) return v;
    }
    for (int ty = 16; ty >= 1; ty >>= 1) {
        const size_t bytes = bwd ? 2ull * V * (ty + 1) * (TX + 1) * 4 : 1ull * V * ty * TX * 4;
        if (bytes <= SMEM_BUDGET) return ty;
    }
    return 0;
}

size_t patch_bytes(int V, int ty, bool bwd) {
    return bwd ? 2ull * V * (ty + 1) * (TX + 1) * 4 : 1ull * V * ty * TX * 4;
}

int load_taps(const float* taps_dev, int ktaps, Taps& t, cudaStream_t st) {
    // taps live in device memory (the caller's buffer may be produced on the stream): stage through
    // a small pinned-free copy.  21 floats; the sync here is on the caller's stream only.
    B3D_CUDA_OK(cudaMemcpyAsync(t.w, taps_dev, sizeof(float) * ktaps, cudaMemcpyDeviceToHost, st));
    B3D_CUDA_OK(cudaStreamSynchronize(st));
    t.n = ktaps;
    return B3D_OK;
}

}  // namespace

extern "C" {

int b3d_pc_project(const float* points, const float* quat, int B, int N, int V, float fov, float cam_dist,
                   float* pg, float* coords, int32_t* base, uint8_t* inb, void* stream) {
    B3D_REQUIRE(B >= 0 && N >= 0 && V >= 2, B3D_EINVAL, "b3d_pc_project: bad sizes B=%d N=%d V=%d", B, N, V);
    if (B == 0 || N == 0) return B3D_OK;
    B3D_REQUIRE(points && quat && pg, B3D_EINVAL, "b3d_pc_project: null pointer");
    B3D_CHECK_ALIGNED(pg);
    dim3 grid(b3d::ceil_div(N, NTHREADS), B);
    pc_project_kernel<<<grid, NTHREADS, 0, (cudaStream_t)stream>>>(points, quat, N, V, fov, cam_dist, (float4*)pg,
                                                                  coords, base, inb);
    B3D_LAUNCH_OK();
    return B3D_OK;
}

size_t b3d_pc_silhouette_workspace_bytes(int B, int V, int mode) {
    if (mode == B3D_MODE_REFERENCE) return 0;
    return 2ull * (size_t)B * V * V * V * sizeof(float);
}

}  // extern "C"

namespace {

template <typename K>
int set_smem(K kernel, size_t bytes) {
    B3D_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return B3D_OK;
}

int sil_fwd_impl(const float* pg, const Taps& t, const float* scale, int B, int N, int V, int mode, float* sil,
                 cudaStream_t st) {
    const int TY = pick_ty(V, false);
    B3D_REQUIRE(TY > 0, B3D_EINVAL, "b3d_pc_silhouette_fwd: V=%d does not fit the shared-memory patch", V);
    const size_t smem = patch_bytes(V, TY, false);
    dim3 grid(b3d::ceil_div(V, TX), b3d::ceil_div(V, TY), B);
    if (t.n == 21) {
        if (int rc = set_smem(pc_sil_fwd_kernel<21>, smem)) return rc;
        pc_sil_fwd_kernel<21><<<grid, NTHREADS, smem, st>>>((const float4*)pg, t, scale, N, V, TY, mode, sil);
    } else {
        if (int rc = set_smem(pc_sil_fwd_kernel<0>, smem)) return rc;
        pc_sil_fwd_kernel<0><<<grid, NTHREADS, smem, st>>>((const float4*)pg, t, scale, N, V, TY, mode, sil);
    }
    B3D_LAUNCH_OK();
    return B3D_OK;
}

int sil_bwd_impl(const float* pg, const Taps& t, const float* scale, const float* dsil, int B, int N, int V,
                 int mode, float* dpg, float* dscale, cudaStream_t st) {
    const int TY = pick_ty(V, true);
    B3D_REQUIRE(TY > 0, B3D_EINVAL, "b3d_pc_silhouette_bwd: V=%d does not fit the shared-memory patch", V);
    const size_t smem = patch_bytes(V, TY, true);
    if (dscale) B3D_CUDA_OK(cudaMemsetAsync(dscale, 0, sizeof(float) * B, st));
    dim3 grid(b3d::ceil_div(V, TX), b3d::ceil_div(V, TY), B);
    if (t.n == 21) {
        if (int rc = set_smem(pc_sil_bwd_kernel<21>, smem)) return rc;
        pc_sil_bwd_kernel<21><<<grid, NTHREADS, smem, st>>>((const float4*)pg, t, scale, dsil, N, V, TY, mode,
                                                           (float4*)dpg, dscale);
    } else {
        if (int rc = set_smem(pc_sil_bwd_kernel<0>, smem)) return rc;
        pc_sil_bwd_kernel<0><<<grid, NTHREADS, smem, st>>>((const float4*)pg, t, scale, dsil, N, V, TY, mode,
                                                          (float4*)dpg, dscale);
    }
    B3D_LAUNCH_OK();
    return B3D_OK;
}

int check_sil_args(const char* who, const void* pg, const void* taps, int ktaps, int B, int N, int V, int mode) {
    B3D_REQUIRE(B >= 0 && N >= 0 && V >= 2, B3D_EINVAL, "%s: bad sizes B=%d N=%d V=%d", who, B, N, V);
    B3D_REQUIRE(ktaps >= 1 && ktaps <= MAX_TAPS && (ktaps & 1), B3D_EINVAL, "%s: ktaps=%d must be odd, <= %d", who,
                ktaps, MAX_TAPS);
    B3D_REQUIRE(mode == B3D_MODE_REFERENCE, B3D_EINVAL,
                "%s: mode %d not available in this build (only B3D_MODE_REFERENCE)", who, mode);
    B3D_REQUIRE(taps && (pg || N == 0 || B == 0), B3D_EINVAL, "%s: null pointer", who);
    return B3D_OK;
}

}  // namespace

extern "C" {

// taps given in HOST memory (the Python wrapper computes them with the reference's torch expression on
// the CPU, 21 floats): avoids the device->host sync of the device-taps entry points.
int b3d_pc_silhouette_fwd_hosttaps(const float* pg, const float* taps_host, int ktaps, const float* scale, int B,
                                   int N, int V, int mode, float* sil, void* workspace, size_t workspace_bytes,
                                   void* stream) {
    (void)workspace;
    (void)workspace_bytes;
    if (int rc = check_sil_args("b3d_pc_silhouette_fwd", pg, taps_host, ktaps, B, N, V, mode)) return rc;
    if (B == 0) return B3D_OK;
    B3D_REQUIRE(sil, B3D_EINVAL, "b3d_pc_silhouette_fwd: null output");
    Taps t;
    for (int i = 0; i < ktaps; ++i) t.w[i] = taps_host[i];
    t.n = ktaps;
    return sil_fwd_impl(pg, t, scale, B, N, V, mode, sil, (cudaStream_t)stream);
}

int b3d_pc_silhouette_fwd(const float* pg, const float* taps, int ktaps, const float* scale, int B, int N, int V,
                          int mode, float* sil, void* workspace, size_t workspace_bytes, void* stream) {
    (void)workspace;
    (void)workspace_bytes;
    if (int rc = check_sil_args("b3d_pc_silhouette_fwd", pg, taps, ktaps, B, N, V, mode)) return rc;
    if (B == 0) return B3D_OK;
    B3D_REQUIRE(sil, B3D_EINVAL, "b3d_pc_silhouette_fwd: null output");
    Taps t;
    if (int rc = load_taps(taps, ktaps, t, (cudaStream_t)stream)) return rc;
    return sil_fwd_impl(pg, t, scale, B, N, V, mode, sil, (cudaStream_t)stream);
}

int b3d_pc_silhouette_bwd_hosttaps(const float* pg, const float* taps_host, int ktaps, const float* scale,
                                   const float* dsil, int B, int N, int V, int mode, float* dpg, float* dscale,
                                   void* workspace, size_t workspace_bytes, void* stream) {
    (void)workspace;
    (void)workspace_bytes;
    if (int rc = check_sil_args("b3d_pc_silhouette_bwd", pg, taps_host, ktaps, B, N, V, mode)) return rc;
    if (B == 0) return B3D_OK;
    B3D_REQUIRE(dsil && (dpg || N == 0), B3D_EINVAL, "b3d_pc_silhouette_bwd: null pointer");
    B3D_REQUIRE((scale == nullptr) == (dscale == nullptr), B3D_EINVAL,
                "b3d_pc_silhouette_bwd: scale and dscale must both be given or both be NULL");
    Taps t;
    for (int i = 0; i < ktaps; ++i) t.w[i] = taps_host[i];
    t.n = ktaps;
    return sil_bwd_impl(pg, t, scale, dsil, B, N, V, mode, dpg, dscale, (cudaStream_t)stream);
}

int b3d_pc_silhouette_bwd(const float* pg, const float* taps, int ktaps, const float* scale, const float* dsil,
                          int B, int N, int V, int mode, float* dpg, float* dscale, void* workspace,
                          size_t workspace_bytes, void* stream) {
    (void)workspace;
    (void)workspace_bytes;
    if (int rc = check_sil_args("b3d_pc_silhouette_bwd", pg, taps, ktaps, B, N, V, mode)) return rc;
    if (B == 0) return B3D_OK;
    B3D_REQUIRE(dsil && (dpg || N == 0), B3D_EINVAL, "b3d_pc_silhouette_bwd: null pointer");
    B3D_REQUIRE((scale == nullptr) == (dscale == nullptr), B3D_EINVAL,
                "b3d_pc_silhouette_bwd: scale and dscale must both be given or both be NULL");
    Taps t;
    if (int rc = load_taps(taps, ktaps, t, (cudaStream_t)stream)) return rc;
    return sil_bwd_impl(pg, t, scale, dsil, B, N, V, mode, dpg, dscale, (cudaStream_t)stream);
}

int b3d_pc_project_bwd(const float* points, const float* quat, const float* pg, const float* dpg, int B, int N,
                       int V, float fov, float cam_dist, float* dpoints, float* dquat, void* stream) {
    B3D_REQUIRE(B >= 0 && N >= 0 && V >= 2, B3D_EINVAL, "b3d_pc_project_bwd: bad sizes");
    if (B == 0) return B3D_OK;
    B3D_REQUIRE(quat && dquat, B3D_EINVAL, "b3d_pc_project_bwd: null pointer");
    B3D_CUDA_OK(cudaMemsetAsync(dquat, 0, sizeof(float) * 4 * B, (cudaStream_t)stream));
    if (N == 0) return B3D_OK;
    B3D_REQUIRE(points && pg && dpg && dpoints, B3D_EINVAL, "b3d_pc_project_bwd: null pointer");
    dim3 grid(b3d::ceil_div(N, NTHREADS), B);
    pc_project_bwd_kernel<<<grid, NTHREADS, 0, (cudaStream_t)stream>>>(points, quat, (const float4*)pg,
                                                                      (const float4*)dpg, N, V, fov, cam_dist, dpoints,
                                                                      dquat);
    B3D_LAUNCH_OK();
    return B3D_OK;
}

int b3d_pc_splat_grid(const float* pg, int B, int N, int V, int mode, float* grid, void* stream) {
    B3D_REQUIRE(B >= 0 && N >= 0 && V >= 2, B3D_EINVAL, "b3d_pc_splat_grid: bad sizes");
    B3D_REQUIRE(mode == B3D_MODE_REFERENCE || mode == B3D_MODE_PAPER, B3D_EINVAL, "b3d_pc_splat_grid: bad mode");
    if (B == 0) return B3D_OK;
    B3D_REQUIRE(grid, B3D_EINVAL, "b3d_pc_splat_grid: null grid");
    const size_t cells = (size_t)B * V * V * V;
    cudaStream_t st = (cudaStream_t)stream;
    B3D_CUDA_OK(cudaMemsetAsync(grid, 0, cells * sizeof(float), st));
    if (N > 0) {
        B3D_REQUIRE(pg, B3D_EINVAL, "b3d_pc_splat_grid: null pg");
        dim3 g(b3d::ceil_div(N, NTHREADS), B);
        pc_splat_grid_kernel<<<g, NTHREADS, 0, st>>>((const float4*)pg, N, V, mode, grid);
        B3D_LAUNCH_OK();
    }
    const int blocks = (int)((cells + NTHREADS * 8 - 1) / (NTHREADS * 8));
    clamp01_kernel<<<blocks < 148 * 8 ? blocks : 148 * 8, NTHREADS, 0, st>>>(grid, cells);
    B3D_LAUNCH_OK();
    return B3D_OK;
}

}  // extern "C"
