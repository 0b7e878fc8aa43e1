// Probe: what does TMA write to shared memory for a 4-D box whose INNER dimension is 32 bytes (8 floats) under the
// 128-byte swizzle modes?  (on-the-fly stem fold: box {c=8, r=4, x=16, n=1} over a [N,H,W,8] tensor, dims (c, row, x, n))
// build: nvcc -gencode arch=compute_100a,code=sm_100a -o tools/probe/tma_probe tools/probe/tma_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void probe(const __grid_constant__ CUtensorMap m, float* out, int nfloats, int expect_bytes) {
    extern __shared__ __align__(1024) unsigned char sm[];
    float* buf = reinterpret_cast<float*>(sm);
    __shared__ uint64_t bar;
    for (int i = threadIdx.x; i < nfloats; i += blockDim.x) buf[i] = -7777.f;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(expect_bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
                         smem_u32(buf)),
                     "l"(reinterpret_cast<uint64_t>(&m)), "r"(smem_u32(&bar)), "r"(0), "r"(1), "r"(2), "r"(0)
                     : "memory");
    }
    // wait (bounded)
    unsigned ok = 0;
    for (int it = 0; it < 2000000 && !ok; ++it)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
    __syncthreads();
    for (int i = threadIdx.x; i < nfloats; i += blockDim.x) out[i] = buf[i];
    if (threadIdx.x == 0) out[nfloats] = (float)ok;
}

int main() {
    const int N = 1, H = 16, W = 32, C = 8;
    std::vector<float> h((size_t)N * H * W * C);
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x)
            for (int c = 0; c < C; ++c) h[((size_t)y * W + x) * C + c] = (float)(c + 10 * y + 1000 * x);   // value = c + 10*row + 1000*x
    float* d;
    cudaMalloc(&d, h.size() * 4);
    cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    const int nfloats = 16 * 1024;     // 64 KB window
    float* out;
    cudaMalloc(&out, (nfloats + 1) * 4);
    CUtensorMapSwizzle modes[3] = {CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_SWIZZLE_32B};
    const char* names[3] = {"SWIZZLE_128B", "SWIZZLE_128B_ATOM_32B", "SWIZZLE_32B"};
    for (int mi = 0; mi < 3; ++mi) {
        CUtensorMap m;
        cuuint64_t gd[4] = {8, (cuuint64_t)H, (cuuint64_t)W, (cuuint64_t)N};          // (c, row, x, n)
        cuuint64_t gs[3] = {(cuuint64_t)W * 32, 32, (cuuint64_t)H * W * 32};
        cuuint32_t bx[4] = {8, 4, 16, 1}, es[4] = {1, 1, 1, 1};
        CUresult r = cuTensorMapEncodeTiled(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, modes[mi],
                                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("== %s: encode rc=%d\n", names[mi], (int)r);
        if (r != CUDA_SUCCESS) continue;
        cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, nfloats * 4);
        probe<<<1, 256, nfloats * 4>>>(m, out, nfloats, 8 * 4 * 16 * 4);
        cudaError_t e = cudaDeviceSynchronize();
        printf("   kernel: %s\n", cudaGetErrorString(e));
        if (e != cudaSuccess) return 1;
        std::vector<float> o(nfloats + 1);
        cudaMemcpy(o.data(), out, (nfloats + 1) * 4, cudaMemcpyDeviceToHost);
        int written = 0, last = -1;
        for (int i = 0; i < nfloats; ++i)
            if (o[i] != -7777.f) { ++written; last = i; }
        printf("   barrier completed=%d, floats written=%d (expected 512), last written float index=%d\n", (int)o[nfloats], written, last);
        // box origin (c=0, row=1, x=2): element (c, r, xx) has value c + 10*(1+r) + 1000*(2+xx); expected linear position (xx*4 + r)*8 + c
        for (int row128 = 0; row128 < 20 && row128 * 32 <= last; ++row128) {
            printf("   smem[%3d*128B]:", row128);
            for (int j = 0; j < 32; ++j) printf(" %6.0f", o[row128 * 32 + j]);
            printf("\n");
        }
    }
    return 0;
}
