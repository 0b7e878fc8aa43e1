// HBM-bound helper kernels between the GAN's convolutions ("elementwise / data-movement between
// convs ... each a full HBM round trip -> fuse"), NHWC fp32, 16-byte vector accesses, one pass each.
//   pad_x_*      replicate / circular padding along x (gan.py:329 F.pad replicate; rendering/utils.py:29-33 circpad)
//                forward = gather, backward = gather of the (up to three) output columns that read an input column.
//   leaky_bwd    gradient of the LeakyReLU that the conv epilogue fused (mask from the sign of the OUTPUT).
#include "b3d_common.cuh"

namespace {
constexpr int NT = 256;

// mode 0 = replicate, 1 = circular
__device__ __forceinline__ int src_col(int wo, int a, int W, int mode) {
    const int w = wo - a;
    if (mode == 0) return w < 0 ? 0 : (w >= W ? W - 1 : w);
    return w < 0 ? w + W : (w >= W ? w - W : w);
}

__global__ void __launch_bounds__(NT)
pad_x_fwd_kernel(const float4* __restrict__ x, float4* __restrict__ out, long long rows, int W, int C4, int a, int mode) {
    const int Wo = W + 2 * a;
    const long long total = rows * Wo * C4;
    for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < total; i += (long long)gridDim.x * NT) {
        const int c = (int)(i % C4);
        const long long t = i / C4;
        const int wo = (int)(t % Wo);
        const long long r = t / Wo;
        out[i] = __ldg(x + (r * W + src_col(wo, a, W, mode)) * C4 + c);
    }
}

__global__ void __launch_bounds__(NT)
pad_x_bwd_kernel(const float4* __restrict__ go, float4* __restrict__ gx, long long rows, int W, int C4, int a, int mode) {
    const int Wo = W + 2 * a;
    const long long total = rows * W * C4;
    for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < total; i += (long long)gridDim.x * NT) {
        const int c = (int)(i % C4);
        const long long t = i / C4;
        const int w = (int)(t % W);
        const long long r = t / W;
        const float4* g = go + (r * Wo) * C4 + c;
        float4 s = __ldg(g + (long long)(w + a) * C4);
        if (mode == 0) {
            if (w == 0)
                for (int k = 0; k < a; ++k) { const float4 v = __ldg(g + (long long)k * C4); s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w; }
            if (w == W - 1)
                for (int k = 0; k < a; ++k) { const float4 v = __ldg(g + (long long)(W + a + k) * C4); s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w; }
        } else {
            if (w < a) { const float4 v = __ldg(g + (long long)(w + a + W) * C4); s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w; }
            if (w >= W - a) { const float4 v = __ldg(g + (long long)(w + a - W) * C4); s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w; }
        }
        gx[i] = s;
    }
}

__global__ void __launch_bounds__(NT)
leaky_bwd_kernel(const float4* __restrict__ gy, const float4* __restrict__ y, float4* __restrict__ out, long long n4, float slope) {
    for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < n4; i += (long long)gridDim.x * NT) {
        const float4 g = __ldg(gy + i), v = __ldg(y + i);
        out[i] = make_float4(v.x >= 0.f ? g.x : g.x * slope, v.y >= 0.f ? g.y : g.y * slope, v.z >= 0.f ? g.z : g.z * slope,
                             v.w >= 0.f ? g.w : g.w * slope);
    }
}

int grid_for(long long n) {
    long long b = (n + NT - 1) / NT;
    const long long cap = 148LL * 16;
    return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}
}  // namespace

extern "C" {
int b3d_pad_x_fwd(const float* x, float* out, long long rows, int W, int C, int amount, int mode, void* stream) {
    B3D_REQUIRE(rows >= 0 && W > 0 && C > 0 && C % 4 == 0 && amount >= 0 && amount <= W && (mode == 0 || mode == 1), B3D_EINVAL,
                "b3d_pad_x_fwd: bad arguments (C=%d must be a multiple of 4, amount=%d <= W=%d)", C, amount, W);
    if (rows == 0) return B3D_OK;
    B3D_REQUIRE(x && out, B3D_EINVAL, "b3d_pad_x_fwd: null pointer");
    B3D_CHECK_ALIGNED(x);
    B3D_CHECK_ALIGNED(out);
    pad_x_fwd_kernel<<<grid_for(rows * (W + 2 * amount) * (C / 4)), NT, 0, (cudaStream_t)stream>>>(
        (const float4*)x, (float4*)out, rows, W, C / 4, amount, mode);
    B3D_LAUNCH_OK();
    return B3D_OK;
}
int b3d_pad_x_bwd(const float* gout, float* gx, long long rows, int W, int C, int amount, int mode, void* stream) {
    B3D_REQUIRE(rows >= 0 && W > 0 && C > 0 && C % 4 == 0 && amount >= 0 && amount <= W && (mode == 0 || mode == 1), B3D_EINVAL,
                "b3d_pad_x_bwd: bad arguments");
    if (rows == 0) return B3D_OK;
    B3D_REQUIRE(gout && gx, B3D_EINVAL, "b3d_pad_x_bwd: null pointer");
    B3D_CHECK_ALIGNED(gout);
    B3D_CHECK_ALIGNED(gx);
    pad_x_bwd_kernel<<<grid_for(rows * W * (C / 4)), NT, 0, (cudaStream_t)stream>>>((const float4*)gout, (float4*)gx, rows, W,
                                                                                C / 4, amount, mode);
    B3D_LAUNCH_OK();
    return B3D_OK;
}
int b3d_leaky_bwd(const float* gy, const float* y, float* out, long long n, float slope, void* stream) {
    B3D_REQUIRE(n >= 0 && n % 4 == 0, B3D_EINVAL, "b3d_leaky_bwd: element count must be a multiple of 4");
    if (n == 0) return B3D_OK;
    B3D_REQUIRE(gy && y && out, B3D_EINVAL, "b3d_leaky_bwd: null pointer");
    leaky_bwd_kernel<<<grid_for(n / 4), NT, 0, (cudaStream_t)stream>>>((const float4*)gy, (const float4*)y, (float4*)out, n / 4, slope);
    B3D_LAUNCH_OK();
    return B3D_OK;
}
}
