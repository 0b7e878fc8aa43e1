// Row-window implicit-GEMM convolution on tcgen05 (sm_100a): third generation of the stride-1 forward / input-gradient
// kernel for layers whose 128-pixel tile is ONE output row segment (G.blk6.*, the folded discriminator stems,
// conv_final's input gradient; models/gan.py:57-65,294-302,359).
//
// tc_conv.cu's persistent kernel re-fetches a 128-pixel x 32-channel A tile for every filter tap: at 64 output channels
// that is 0.034 bytes per FLOP through the L2 -> SM fabric, and ncu shows those launches bounded by it (tensor pipe
// 27-38 %, L2 -> SM 10-13 TB/s; profiles/r1_c_conv_final_full.md).  Here the producer stages, per (filter row, 32-channel
// slice), ONE window of 128 + kw - 1 input pixels per stacked tile and ONE 3-D box with the kw weight tiles of that filter
// row; the MMA issuer feeds the kw taps as shifted views of the window — the UMMA descriptor start address moves by whole
// 128-byte rows (the 128 B swizzle is keyed on absolute shared-memory address bits, so TMA's write pattern and the shifted
// reads agree; the same property tc_conv2.cu and the row-of-taps wgrad rely on).  A-operand bytes / 3 (3x3) or / 5 (1x5,
// 5x5): 0.015-0.019 B/FLOP.  Structure otherwise as the persistent kernel: warp 0 TMA producer, warp 1 MMA issuer,
// warps 2-5 epilogue, two TMEM accumulator buffers of R x BN columns, mbarrier ring across work items.
#include "tc_common.cuh"
#include "tc_rowwin.cuh"
#include <stdlib.h>

namespace {

constexpr int BM = 128, BK = 32, UMMA_K = 8, NTHREADS = 192;

struct RowParams {
    int N, Hout, Cout, xlo, xhi;
    int tiles_x, tiles, groups, work;
    int kh, kslices, ncls;
    int dy[4][5], dx0, shift[5];
    int OH, OW, OC, ooy[4], oox[4], osy, osx;
    int wtap0[4][5];
    float leaky;
    double* stats;
    const float* mask;     // nullable: activation adjoint fused into the epilogue (see ConvParams in tc_conv.cu)
    float mslope;
    int stats_sum;
};

template <int BN, int KW, int R, int STAGES>
struct RSmem {
    static constexpr int WIN_ROWS = BM + KW - 1;
    static constexpr int WIN_BYTES = WIN_ROWS * 128;                       // what one TMA box delivers
    static constexpr int WIN_STRIDE = (WIN_BYTES + 1023) / 1024 * 1024;    // windows start on swizzle-atom boundaries
    static constexpr int B_BYTES = KW * BN * 128;
    static constexpr int STAGE_BYTES = R * WIN_STRIDE + B_BYTES;
    static constexpr int TX_BYTES = R * WIN_BYTES + B_BYTES;
    static constexpr int TOTAL = STAGES * STAGE_BYTES + 1024 + 320 + 2 * BN * 4;
};

template <int BN, int KW, int R, int STAGES>
__global__ void __launch_bounds__(NTHREADS, 1)
conv_rowwin_tf32_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w, const RowParams p,
                        const float* __restrict__ bias, float* __restrict__ out) {
    using S = RSmem<BN, KW, R, STAGES>;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full = reinterpret_cast<uint64_t*>(base + STAGES * S::STAGE_BYTES);
    uint64_t* empty = full + STAGES;
    uint64_t* acc_full = empty + STAGES;
    uint64_t* acc_empty = acc_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
    float* sm_stats = reinterpret_cast<float*>(base + STAGES * S::STAGE_BYTES + 320);
    // BN = 16 (thin heads): the epilogue reads 32-column groups, so the allocation keeps 32 spare columns behind the last one
    constexpr uint32_t TCOLS = BN < 32 ? 4 * R * BN : 2 * R * BN;
    static_assert(TCOLS <= 512 && (TCOLS & (TCOLS - 1)) == 0 && TCOLS >= 32, "TMEM: 2 buffers x R accumulators x BN columns");
    for (int i = threadIdx.x; i < 2 * BN; i += blockDim.x) sm_stats[i] = 0.f;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        tc::tma_prefetch_desc(&tmap_x);
        tc::tma_prefetch_desc(&tmap_w);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            tc::mbar_init(full + s, 1);
            tc::mbar_init(empty + s, 1);
        }
        for (int b = 0; b < 2; ++b) {
            tc::mbar_init(acc_full + b, 1);
            tc::mbar_init(acc_empty + b, 4);
        }
        tc::fence_barrier_init();
    }
    if (warp == 2) tc::tmem_alloc<TCOLS>(tmem_slot);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_acc = *tmem_slot;
    const int KI = p.kh * p.kslices;

    if (warp == 0) {
        {
            const uint32_t leader = tc::elect_one();          // convergent producer loop: one elected lane arrives / issues the copies
            uint32_t git = 0;
            for (int w = blockIdx.x; w < p.work; w += gridDim.x) {
                // classes are the fastest index: CTAs working on the same pixel tiles at the same time share them in L2
                const int cls = w % p.ncls, wq = w / p.ncls;
                const int g = wq % p.groups, c0 = (wq / p.groups) * BN;
                int x0[R], y0[R], n0[R];
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    int t = g * R + r;
                    t = t < p.tiles ? t : p.tiles - 1;          // a group past the end re-loads the last tile (result dropped)
                    x0[r] = p.xlo + (t % p.tiles_x) * BM;
                    t /= p.tiles_x;
                    y0[r] = t % p.Hout;
                    n0[r] = t / p.Hout;
                }
                for (int it = 0; it < KI; ++it, ++git) {
                    const int s = git % STAGES, ph = (git / STAGES) & 1;
                    tc::mbar_wait(empty + s, ph ^ 1);
                    const int fr = it / p.kslices, ks = it % p.kslices;
                    unsigned char* a = base + s * S::STAGE_BYTES;
                    tc::mbar_arrive_expect_tx_if(leader, full + s, S::TX_BYTES);
#pragma unroll
                    for (int r = 0; r < R; ++r)
                        tc::tma_load_4d_if(leader, a + r * S::WIN_STRIDE, &tmap_x, full + s, ks * BK, x0[r] + p.dx0, y0[r] + p.dy[cls][fr], n0[r]);
                    tc::tma_load_3d_if(leader, a + R * S::WIN_STRIDE, &tmap_w, full + s, ks * BK, c0, p.wtap0[cls][fr]);
                }
            }
        }
    } else if (warp == 1) {
        // convergent issue loop (tc_common.cuh "MMA issue from a CONVERGENT warp"): all lanes walk it, one elected lane issues
        const uint32_t leader = tc::elect_one();
        const uint32_t tmem_u = tc::warp_uniform(tmem_acc);
        constexpr uint32_t idesc = tc::umma_idesc_tf32(BM, BN, false, false);
        uint32_t git = 0, j = 0;
        for (int w = blockIdx.x; w < p.work; w += gridDim.x, ++j) {
            const uint32_t buf = j & 1;
            tc::mbar_wait(acc_empty + buf, ((j >> 1) & 1) ^ 1);
            tc::tc_fence_after();
            const uint32_t acc = tmem_u + buf * (R * BN);
            for (int it = 0; it < KI; ++it, ++git) {
                const int s = git % STAGES, ph = (git / STAGES) & 1;
                tc::mbar_wait(full + s, ph);
                tc::tc_fence_after();
                const uint32_t a = tc::smem_u32(base + s * S::STAGE_BYTES);
                const uint64_t da0 = tc::umma_desc_k128(a), db0 = tc::umma_desc_k128(a + R * S::WIN_STRIDE);
                const uint32_t hi = tc::desc_hi(da0);
#pragma unroll
                for (int t = 0; t < KW; ++t) {
                    const uint32_t lat = tc::desc_lo(da0) + (uint32_t)p.shift[t] * (128u >> 4);     // window row shift of tap t
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        const uint32_t lb = tc::desc_lo(db0) + ((t * (BN * 128) + k * UMMA_K * 4) >> 4);
#pragma unroll
                        for (int r = 0; r < R; ++r)
                            tc::umma_tf32_words_if(leader, acc + r * BN, lat + ((r * S::WIN_STRIDE + k * UMMA_K * 4) >> 4), hi, lb, hi, idesc,
                                                   (it | t | k) ? 1u : 0u);
                    }
                }
                tc::umma_commit_if(leader, empty + s);
            }
            tc::umma_commit_if(leader, acc_full + buf);
        }
    } else {
        const int q = warp & 3;
        const int row = q * 32 + lane;
        uint32_t j = 0;
        for (int w = blockIdx.x; w < p.work; w += gridDim.x, ++j) {
            const int cls = w % p.ncls, wq = w / p.ncls;
            const int g = wq % p.groups, c0 = (wq / p.groups) * BN;
            const int ooy = p.ooy[cls], oox = p.oox[cls];
            const uint32_t buf = j & 1;
            constexpr int CW = BN < 32 ? 1 : BN / 32, NW = R * CW;
            uint32_t mbits[NW];
            if (p.mask) {                                         // signs of the item's activation mask, fetched under its main loop
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    int t = g * R + r;
                    const bool tile_ok = t < p.tiles;
                    const int x = p.xlo + (t % p.tiles_x) * BM + row;
                    t /= p.tiles_x;
                    const int y = t % p.Hout, n = t / p.Hout;
                    const bool valid = tile_ok && x < p.xhi;
                    const size_t off = (((size_t)n * p.OH + (size_t)(p.osy * y + ooy)) * p.OW + (size_t)(p.osx * x + oox)) * p.OC;
#pragma unroll
                    for (int cc = 0; cc < CW; ++cc) {
                        const int cb = c0 + 32 * cc;
                        mbits[r * CW + cc] = (valid && cb < p.Cout)
                            ? tc::act_mask_bits32(p.mask + off + cb, cb + 32 <= p.Cout && (p.OC & 3) == 0, p.Cout - cb) : 0xffffffffu;
                    }
                }
            }
            tc::mbar_wait(acc_full + buf, (j >> 1) & 1);
            tc::tc_fence_after();
#pragma unroll 1
            for (int r = 0; r < R; ++r) {
                int t = g * R + r;
                const bool tile_ok = t < p.tiles;
                const int x = p.xlo + (t % p.tiles_x) * BM + row;
                t /= p.tiles_x;
                const int y = t % p.Hout, n = t / p.Hout;
                const bool valid = tile_ok && x < p.xhi;
                const size_t off = (((size_t)n * p.OH + (size_t)(p.osy * y + ooy)) * p.OW + (size_t)(p.osx * x + oox)) * p.OC;
                float* dst = out + off;
#pragma unroll 1
                for (int c = 0; c < BN; c += 32) {
                    float v[32];
                    tc::tmem_ld32(tmem_acc + ((uint32_t)(q * 32) << 16) + buf * (R * BN) + r * BN + (uint32_t)c, v);
                    if (r == R - 1 && c + 32 >= BN) {            // last read of this buffer: hand it back before the stores
                        tc::tc_fence_before();
                        __syncwarp();
                        if (lane == 0) tc::mbar_arrive(acc_empty + buf);
                    }
                    if (p.mask) tc::apply_act_bits32(v, tc::pick_word<NW>(mbits, r * CW + c / 32), p.mslope);
                    if (BN >= 32 && p.stats) {
                        if (p.stats_sum) tc::stats_accumulate_sum(v, valid, sm_stats, c);
                        else tc::stats_accumulate(v, valid, sm_stats, BN, c);
                    }
                    if (valid) {
                        const int cb = c0 + c;
                        if (cb + 32 <= p.Cout && (p.OC & 3) == 0) {
#pragma unroll
                            for (int jj = 0; jj < 32; jj += 4) {
                                float4 o;
                                o.x = v[jj] + (bias ? __ldg(bias + cb + jj) : 0.f);
                                o.y = v[jj + 1] + (bias ? __ldg(bias + cb + jj + 1) : 0.f);
                                o.z = v[jj + 2] + (bias ? __ldg(bias + cb + jj + 2) : 0.f);
                                o.w = v[jj + 3] + (bias ? __ldg(bias + cb + jj + 3) : 0.f);
                                o.x = o.x >= 0.f ? o.x : o.x * p.leaky;
                                o.y = o.y >= 0.f ? o.y : o.y * p.leaky;
                                o.z = o.z >= 0.f ? o.z : o.z * p.leaky;
                                o.w = o.w >= 0.f ? o.w : o.w * p.leaky;
                                *reinterpret_cast<float4*>(dst + cb + jj) = o;
                            }
                        } else {
#pragma unroll
                            for (int jj = 0; jj < 32; ++jj) {
                                const int co = cb + jj;
                                if (co < p.Cout) {
                                    const float o = v[jj] + (bias ? __ldg(bias + co) : 0.f);
                                    dst[co] = o >= 0.f ? o : o * p.leaky;
                                }
                            }
                        }
                    }
                }
            }
            if (BN >= 32 && p.stats) tc::stats_flush(sm_stats, BN, p.stats, p.Cout, c0, threadIdx.x - 64);
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 2) tc::tmem_dealloc<TCOLS>(tmem_acc);
}

template <int BN, int KW, int R, int STAGES>
int launch_rowwin(const b3d::RowWinArgs& a, cudaStream_t st) {
    using S = RSmem<BN, KW, R, STAGES>;
    static_assert(S::TOTAL <= 227 * 1024, "row-window pipeline does not fit shared memory");
    RowParams p{};
    p.N = a.N; p.Hout = a.Hout; p.Cout = a.Cout; p.xlo = a.xlo; p.xhi = a.xhi;
    p.tiles_x = b3d::ceil_div(a.xhi - a.xlo, BM);
    p.tiles = p.tiles_x * a.Hout * a.N;
    p.groups = b3d::ceil_div(p.tiles, R);
    p.ncls = a.ncls < 1 ? 1 : a.ncls;
    p.work = p.groups * b3d::ceil_div(a.Cout, BN) * p.ncls;
    p.kh = a.kh; p.kslices = a.Cin / BK;
    for (int i = 0; i < 5; ++i) p.shift[i] = a.shift[i];
    for (int c = 0; c < 4; ++c) {
        p.ooy[c] = a.ooy[c]; p.oox[c] = a.oox[c];
        for (int i = 0; i < 5; ++i) { p.dy[c][i] = a.dy[c][i]; p.wtap0[c][i] = a.wtap0[c][i]; }
    }
    p.dx0 = a.dx0;
    p.OH = a.OH; p.OW = a.OW; p.OC = a.OC; p.leaky = a.leaky; p.stats = a.stats;
    p.osy = a.osy; p.osx = a.osx;
    p.mask = a.mask; p.mslope = a.mslope; p.stats_sum = a.stats_sum;
    CUtensorMap mx, mw;
    {
        const uint64_t dims[4] = {(uint64_t)a.Cin, (uint64_t)a.W, (uint64_t)a.H, (uint64_t)a.N};
        const uint64_t pitch = a.xpitch ? a.xpitch : a.W;                  // pixels per image row in memory
        const uint64_t strides[3] = {(uint64_t)a.Cin * 4, pitch * a.Cin * 4, (uint64_t)a.H * pitch * a.Cin * 4};
        const uint32_t box[4] = {(uint32_t)BK, (uint32_t)S::WIN_ROWS, 1, 1};
        if (int rc = tc::make_tmap_f32(&mx, a.x, 4, dims, strides, box)) return rc;
    }
    {
        // one box = the KW weight tiles of a filter row: taps wtap0[r], wtap0[r] + step, ... (element stride on the tap dim)
        const uint64_t dims[3] = {(uint64_t)a.Cin, (uint64_t)a.Cout, (uint64_t)a.wtaps_total};
        const uint64_t strides[2] = {(uint64_t)a.Cin * 4, (uint64_t)a.Cout * a.Cin * 4};
        const uint32_t box[3] = {(uint32_t)BK, (uint32_t)BN, (uint32_t)((KW - 1) * a.wtap_step + 1)};
        const uint32_t es[3] = {1, 1, (uint32_t)a.wtap_step};
        if (int rc = tc::make_tmap_f32(&mw, a.wt, 3, dims, strides, box, es)) return rc;
    }
    B3D_CUDA_OK(cudaFuncSetAttribute(conv_rowwin_tf32_kernel<BN, KW, R, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
    const int grid = p.work < 148 ? p.work : 148;
    conv_rowwin_tf32_kernel<BN, KW, R, STAGES><<<grid, NTHREADS, S::TOTAL, st>>>(mx, mw, p, a.bias, a.out);
    B3D_LAUNCH_OK();
    b3d::add_variant("conv_rowwin_tf32<%d,%d,%d,%d>", BN, KW, R, STAGES);
    return B3D_OK;
}

}  // namespace

namespace b3d {
int conv_rowwin_launch(const RowWinArgs& a, cudaStream_t st) {
    if (a.xhi - a.xlo < BM || a.Cin % BK || a.kh < 1 || a.kh > 5) return 1;
    // short K loops (folded stems: 1 filter row x 2 channel slices) cannot hide the two-stage ring's window loads behind
    // MMA work: measured slower than the per-tap kernel (0.71 vs 0.62 ms, profiles/r2_e_conv_layers.md)
    if (a.kh * (a.Cin / BK) < 4) return 1;
    // one CTA per SM: worth it only with >= ~2 waves of stacked work items
    const long long tiles = (long long)ceil_div(a.xhi - a.xlo, BM) * a.Hout * a.N * (a.ncls < 1 ? 1 : a.ncls);
    if (a.Cout <= 16 && a.kw == 5 && a.stats == nullptr) {
        // 1-16 output channels (conv_final 64 -> 3): N = 16 tensor-core tiles — 13 of 16 columns are padding, still ~2.5x the
        // fp32 CUDA-core kernel (one window load feeds five taps)
        if (tiles / 4 < 2 * 148) return 1;
        return launch_rowwin<16, 5, 4, 2>(a, st);
    }
    if (a.Cout == 64) {
        if (tiles / 4 < 2 * 148) return 1;
        if (a.kw == 2) return launch_rowwin<64, 2, 4, 2>(a, st);        // stride-2 dgrad parity classes (2 x 2 taps)
        static const int r2 = getenv("B3D_ROWWIN_R2") ? atoi(getenv("B3D_ROWWIN_R2")) : 0;     // experiment: 2 tiles x 3 stages
        if (a.kw == 3) return r2 ? launch_rowwin<64, 3, 2, 3>(a, st) : launch_rowwin<64, 3, 4, 2>(a, st);
        if (a.kw == 5) return launch_rowwin<64, 5, 4, 2>(a, st);
        return 1;
    }
    if (a.Cout == 128) {
        if (tiles / 2 < 2 * 148) return 1;
        if (a.kw == 2) return launch_rowwin<128, 2, 2, 2>(a, st);
        if (a.kw == 3) return launch_rowwin<128, 3, 2, 2>(a, st);
        return 1;
    }
    return 1;
}
}  // namespace b3d
