// Stride-1 implicit-GEMM convolution, second generation: halo-staged input, weight-tile reuse across R
// accumulators (sm_100a, tcgen05 kind::tf32).  Same math and operand formats as tc_conv.cu:
//
//   Y[n, y, x, co] = sum_t sum_ci X[n, y + dy[t], x + dx[t], ci] * Wt[t, co, ci]          (NHWC, fp32 storage)
//
// tc_conv.cu reloads a 128-pixel input tile for every filter tap (9x / 25x the unique bytes) and every CTA streams
// its weight slice for just 128 pixels: at 16 MAC per byte moved L2 -> SM the tensor cores idle (ncu / per-layer
// timings in profiles/).  Here
//   * pixels are indexed in the FLATTENED input space q = y*P + x (P = padded input width = row pitch), so a tap is
//     a constant row shift off[t] = dy*P + dx: the CTA stages rows [q0 + min_off, q0 + R*128 + max_off) of one
//     32-channel slice ONCE (a few TMA boxes, zero fill outside the image = the y padding) and feeds every tap to
//     the tensor core as a shifted view of that staging buffer — the UMMA descriptor start address moves by whole
//     128-byte rows and its base_offset field carries the swizzle phase of the unaligned start;
//   * the CTA owns R stacked 128-pixel tiles (R accumulators of BN columns in TMEM, up to all 512 columns), so every
//     weight tile that arrives is used R times and the input halo is shared between the stacked tiles.
// Outputs whose x falls into the pad columns (x >= Wout) are computed and dropped (<= (kw-1)/P waste).
// Warp roles as in tc_conv.cu (TMA producer / MMA issuer / 4 epilogue warps); the staging buffer is double buffered
// across channel slices, weights stream through their own ring.
#include "tc_common.cuh"

namespace {

constexpr int BM = 128, BK = 32, UMMA_K = 8, MAX_TAPS = 25, NTHREADS = 192, NB = 3;

struct FlatParams {
    int N, Hout, Wout, Cout, P;
    int ntaps, kslices, R, tiles_per_img;
    int off[MAX_TAPS];        // off[t] - min_off  (>= 0): row of the staging buffer where tap t starts for tile 0
    int min_off;              // flattened offset of staged row 0 relative to q0
    int main_boxes, tail_rows, staged_rows;
    int OH, OW, OC;
    float leaky;
    int use_base_offset;
};

__device__ __forceinline__ uint64_t desc_k128_at(uint32_t addr, int use_bo) {
    uint64_t d = tc::umma_desc_k128(addr);
    if (use_bo) d |= (uint64_t)((addr >> 7) & 7) << 49;      // swizzle phase of a start that is not 1024-B aligned
    return d;
}

template <int BN>
__global__ void __launch_bounds__(NTHREADS, 1)
conv_flat_tf32_kernel(const __grid_constant__ CUtensorMap tmap_main, const __grid_constant__ CUtensorMap tmap_tail,
                      const __grid_constant__ CUtensorMap tmap_w, const FlatParams p, const float* __restrict__ bias,
                      float* __restrict__ out, int a_stage_bytes, int tmem_cols_log2) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    unsigned char* a_buf = base;                                   // 2 staging buffers
    unsigned char* b_buf = base + 2 * a_stage_bytes;               // NB weight stages
    constexpr int B_BYTES = BN * BK * 4;
    uint64_t* a_full = reinterpret_cast<uint64_t*>(b_buf + NB * B_BYTES);
    uint64_t* a_empty = a_full + 2;
    uint64_t* b_full = a_empty + 2;
    uint64_t* b_empty = b_full + NB;
    uint64_t* acc_full = b_empty + NB;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = blockIdx.x / p.tiles_per_img;
    const int q0 = (blockIdx.x % p.tiles_per_img) * p.R * BM;
    const int c0 = blockIdx.y * BN;

    if (warp == 0 && lane == 0) {
        tc::tma_prefetch_desc(&tmap_main);
        tc::tma_prefetch_desc(&tmap_tail);
        tc::tma_prefetch_desc(&tmap_w);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < 2; ++i) {
            tc::mbar_init(a_full + i, 1);
            tc::mbar_init(a_empty + i, 1);
        }
        for (int i = 0; i < NB; ++i) {
            tc::mbar_init(b_full + i, 1);
            tc::mbar_init(b_empty + i, 1);
        }
        tc::mbar_init(acc_full, 1);
        tc::fence_barrier_init();
    }
    if (warp == 2) {
        if (tmem_cols_log2 == 9) tc::tmem_alloc<512>(tmem_slot);
        else if (tmem_cols_log2 == 8) tc::tmem_alloc<256>(tmem_slot);
        else if (tmem_cols_log2 == 7) tc::tmem_alloc<128>(tmem_slot);
        else tc::tmem_alloc<64>(tmem_slot);
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_acc = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            auto load_a = [&](int ks) {
                const int s = ks & 1, ph = (ks >> 1) & 1;
                tc::mbar_wait(a_empty + s, ph ^ 1);
                tc::mbar_arrive_expect_tx(a_full + s, (uint32_t)p.staged_rows * 128u);
                unsigned char* dst = a_buf + s * a_stage_bytes;
                const int row0 = q0 + p.min_off;
                for (int bx = 0; bx < p.main_boxes; ++bx)
                    tc::tma_load_3d(dst + bx * 256 * 128, &tmap_main, a_full + s, ks * BK, row0 + bx * 256, n);
                if (p.tail_rows)
                    tc::tma_load_3d(dst + p.main_boxes * 256 * 128, &tmap_tail, a_full + s, ks * BK,
                                    row0 + p.main_boxes * 256, n);
            };
            load_a(0);
            int it = 0;
            for (int ks = 0; ks < p.kslices; ++ks) {
                if (ks + 1 < p.kslices) load_a(ks + 1);
                for (int t = 0; t < p.ntaps; ++t, ++it) {
                    const int s = it % NB, ph = (it / NB) & 1;
                    tc::mbar_wait(b_empty + s, ph ^ 1);
                    tc::mbar_arrive_expect_tx(b_full + s, B_BYTES);
                    tc::tma_load_3d(b_buf + s * B_BYTES, &tmap_w, b_full + s, ks * BK, c0, t);
                }
            }
        }
    } else if (warp == 1) {
        // convergent issue loop (tc_common.cuh "MMA issue from a CONVERGENT warp"): all lanes walk it, one elected lane issues
        const uint32_t leader = tc::elect_one();
        const uint32_t tmem_u = tc::warp_uniform(tmem_acc);
        constexpr uint32_t idesc = tc::umma_idesc_tf32(BM, BN);
        int it = 0;
        for (int ks = 0; ks < p.kslices; ++ks) {
            const int sa = ks & 1;
            tc::mbar_wait(a_full + sa, (ks >> 1) & 1);
            const uint32_t a0 = tc::smem_u32(a_buf + sa * a_stage_bytes);
            for (int t = 0; t < p.ntaps; ++t, ++it) {
                const int s = it % NB, ph = (it / NB) & 1;
                tc::mbar_wait(b_full + s, ph);
                tc::tc_fence_after();
                const uint64_t db0 = tc::umma_desc_k128(tc::smem_u32(b_buf + s * B_BYTES));
                for (int j = 0; j < p.R; ++j) {
                    const uint32_t arow = a0 + (uint32_t)(p.off[t] + j * BM) * 128u;
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        const uint64_t da = desc_k128_at(arow + k * UMMA_K * 4, p.use_base_offset);
                        tc::umma_tf32_words_if(leader, tmem_u + j * BN, tc::desc_lo(da), tc::desc_hi(da),
                                               tc::desc_lo(db0) + ((k * UMMA_K * 4) >> 4), tc::desc_hi(db0), idesc, (ks | t | k) ? 1u : 0u);
                    }
                }
                tc::umma_commit_if(leader, b_empty + s);
            }
            tc::umma_commit_if(leader, a_empty + sa);
        }
        tc::umma_commit_if(leader, acc_full);
    } else {
        const int q = warp & 3;
        tc::mbar_wait(acc_full, 0);
        tc::tc_fence_after();
        for (int j = 0; j < p.R; ++j) {
            const int vq = q0 + j * BM + q * 32 + lane;          // virtual output pixel
            const int y = vq / p.P, x = vq % p.P;
            const bool valid = y < p.Hout && x < p.Wout;
            float* dst = out + (((size_t)n * p.OH + y) * p.OW + x) * p.OC;
#pragma unroll 1
            for (int c = 0; c < BN; c += 32) {
                float v[32];
                tc::tmem_ld32(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(j * BN + c), v);
                if (valid) {
                    const int cb = c0 + c;
                    if (cb + 32 <= p.Cout && (p.OC & 3) == 0) {
#pragma unroll
                        for (int i = 0; i < 32; i += 4) {
                            float4 o;
                            o.x = v[i] + (bias ? __ldg(bias + cb + i) : 0.f);
                            o.y = v[i + 1] + (bias ? __ldg(bias + cb + i + 1) : 0.f);
                            o.z = v[i + 2] + (bias ? __ldg(bias + cb + i + 2) : 0.f);
                            o.w = v[i + 3] + (bias ? __ldg(bias + cb + i + 3) : 0.f);
                            o.x = o.x >= 0.f ? o.x : o.x * p.leaky;
                            o.y = o.y >= 0.f ? o.y : o.y * p.leaky;
                            o.z = o.z >= 0.f ? o.z : o.z * p.leaky;
                            o.w = o.w >= 0.f ? o.w : o.w * p.leaky;
                            *reinterpret_cast<float4*>(dst + cb + i) = o;
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            const int co = cb + i;
                            if (co < p.Cout) {
                                float o = v[i] + (bias ? __ldg(bias + co) : 0.f);
                                dst[co] = o >= 0.f ? o : o * p.leaky;
                            }
                        }
                    }
                }
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        if (tmem_cols_log2 == 9) tc::tmem_dealloc<512>(tmem_acc);
        else if (tmem_cols_log2 == 8) tc::tmem_dealloc<256>(tmem_acc);
        else if (tmem_cols_log2 == 7) tc::tmem_dealloc<128>(tmem_acc);
        else tc::tmem_dealloc<64>(tmem_acc);
    }
}

constexpr int SMEM_LIMIT = 225 * 1024;

}  // namespace

extern "C" {

// Stride-1 convolution over the flattened pixel space.  x [N,H,P,Cin] NHWC (P = padded width, Cin % 32 == 0),
// wt [ntaps,Cout,Cin], out [N,OH,OW,OC]: out[n,y,x,co] for y < Hout, x < Wout.  Taps (dy, dx) with dx >= 0 reaching at
// most column x + dx <= P - 1 for valid outputs (the caller's x padding guarantees it).
// Returns B3D_EINVAL with "does not fit" when the staged halo exceeds shared memory (caller falls back to b3d_conv2d_tf32).
int b3d_conv2d_flat_tf32(const float* x, const float* wt, const float* bias, float* out, int N, int H, int P, int Cin,
                         int Hout, int Wout, int Cout, int ntaps, const int* dy, const int* dx, int OH, int OW, int OC,
                         float leaky, void* stream) {
    B3D_REQUIRE(N > 0 && H > 0 && P > 0 && Hout > 0 && Wout > 0 && Cout > 0, B3D_EINVAL, "b3d_conv2d_flat_tf32: bad sizes");
    B3D_REQUIRE(Cin > 0 && Cin % BK == 0, B3D_EINVAL, "b3d_conv2d_flat_tf32: Cin=%d must be a multiple of %d", Cin, BK);
    B3D_REQUIRE(ntaps >= 1 && ntaps <= MAX_TAPS && dy && dx && x && wt && out, B3D_EINVAL, "b3d_conv2d_flat_tf32: bad arguments");
    B3D_REQUIRE(Wout <= P, B3D_EINVAL, "b3d_conv2d_flat_tf32: Wout=%d exceeds the input pitch %d", Wout, P);
    B3D_CHECK_ALIGNED(x);
    B3D_CHECK_ALIGNED(wt);
    const int BN = Cout > 64 ? 128 : 64;
    FlatParams p{};
    int mn = 1 << 30, mx = -(1 << 30);
    for (int t = 0; t < ntaps; ++t) {
        const int o = dy[t] * P + dx[t];
        mn = o < mn ? o : mn;
        mx = o > mx ? o : mx;
    }
    const int span = mx - mn;
    const int b_bytes = BN * BK * 4;
    int R = 0, a_stage = 0;
    for (int r = (512 / BN < 4 ? 512 / BN : 4); r >= 1; --r) {
        const int rows = r * BM + span;
        const int a = ((rows * 128) + 1023) & ~1023;
        if (2 * a + NB * b_bytes + 2048 <= SMEM_LIMIT) {
            R = r;
            a_stage = a;
            break;
        }
    }
    B3D_REQUIRE(R > 0, B3D_EINVAL, "b3d_conv2d_flat_tf32: does not fit (halo of %d rows)", span);
    const long long vpix = (long long)Hout * P;
    // do not stack more tiles than the image has
    while (R > 1 && (long long)(R - 1) * BM >= vpix) --R;
    p.N = N; p.Hout = Hout; p.Wout = Wout; p.Cout = Cout; p.P = P;
    p.ntaps = ntaps; p.kslices = Cin / BK; p.R = R;
    p.tiles_per_img = (int)((vpix + R * BM - 1) / (R * BM));
    for (int t = 0; t < ntaps; ++t) p.off[t] = dy[t] * P + dx[t] - mn;
    p.min_off = mn;
    p.staged_rows = R * BM + span;
    p.main_boxes = p.staged_rows / 256;
    p.tail_rows = p.staged_rows % 256;
    p.OH = OH; p.OW = OW; p.OC = OC; p.leaky = leaky;
    p.use_base_offset = 0;      // measured: the swizzle is keyed on absolute smem address bits; a non-zero base_offset breaks it

    CUtensorMap m_main, m_tail, m_w;
    const uint64_t dims[3] = {(uint64_t)Cin, (uint64_t)H * P, (uint64_t)N};
    const uint64_t strides[2] = {(uint64_t)Cin * 4, (uint64_t)H * P * Cin * 4};
    {
        const uint32_t box[3] = {(uint32_t)BK, 256, 1};
        if (int rc = tc::make_tmap_f32(&m_main, x, 3, dims, strides, box)) return rc;
        const uint32_t boxt[3] = {(uint32_t)BK, (uint32_t)(p.tail_rows ? p.tail_rows : 8), 1};
        if (int rc = tc::make_tmap_f32(&m_tail, x, 3, dims, strides, boxt)) return rc;
    }
    {
        const uint64_t wd[3] = {(uint64_t)Cin, (uint64_t)Cout, (uint64_t)ntaps};
        const uint64_t ws[2] = {(uint64_t)Cin * 4, (uint64_t)Cout * Cin * 4};
        const uint32_t wb[3] = {(uint32_t)BK, (uint32_t)BN, 1};
        if (int rc = tc::make_tmap_f32(&m_w, wt, 3, wd, ws, wb)) return rc;
    }
    const int cols = R * BN;
    const int log2c = cols > 256 ? 9 : cols > 128 ? 8 : cols > 64 ? 7 : 6;
    const int smem = 2 * a_stage + NB * b_bytes + 2048;
    dim3 grid(N * p.tiles_per_img, b3d::ceil_div(Cout, BN));
    cudaStream_t st = (cudaStream_t)stream;
    if (BN == 128) {
        B3D_CUDA_OK(cudaFuncSetAttribute(conv_flat_tf32_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        conv_flat_tf32_kernel<128><<<grid, NTHREADS, smem, st>>>(m_main, m_tail, m_w, p, bias, out, a_stage, log2c);
    } else {
        B3D_CUDA_OK(cudaFuncSetAttribute(conv_flat_tf32_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        conv_flat_tf32_kernel<64><<<grid, NTHREADS, smem, st>>>(m_main, m_tail, m_w, p, bias, out, a_stage, log2c);
    }
    B3D_LAUNCH_OK();
    b3d::clear_variant();
    b3d::add_variant("conv_flat_tf32<%d>R%d", BN, R);
    return B3D_OK;
}

}  // extern "C"
