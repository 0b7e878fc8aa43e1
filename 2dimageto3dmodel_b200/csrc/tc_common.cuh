// sm_100a tensor-core plumbing shared by the GEMM / implicit-GEMM conv kernels: mbarrier, TMA
// (cp.async.bulk.tensor), tcgen05 (alloc / mma kind::tf32 / commit / ld) wrappers in inline PTX, UMMA
// shared-memory and instruction descriptors, and the host-side tensor-map encoder (driver entry point
// resolved at run time through cudaGetDriverEntryPoint: the library does not link libcuda).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "b3d_common.cuh"

namespace tc {

// ---------------------------------------------------------------------------------------------- mbarrier
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// ---------------------------------------------------------------------------------------------- TMA
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(smem)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
            smem_u32(smem)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* smem, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
        "[%2];" ::"r"(smem_u32(smem)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

__device__ __forceinline__ void tma_load_5d(void* smem, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2,
                                            int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], "
        "[%2];" ::"r"(smem_u32(smem)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}

// The same, issued only where `on` is non-zero: the producer warps run their loops convergently too (see "MMA issue from a
// CONVERGENT warp" below: coordinates, shared-memory addresses and barrier addresses stay in uniform registers, no per-load
// ELECT / R2UR waterfall) and one elected lane performs the arrive and the bulk-tensor copies.
__device__ __forceinline__ void mbar_arrive_expect_tx_if(uint32_t on, uint64_t* bar, uint32_t bytes) {
    asm volatile(
        "{\n"
        ".reg .pred q;\n"
        "setp.ne.b32 q, %2, 0;\n"
        "@q mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(bytes), "r"(on)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d_if(uint32_t on, void* smem, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "{\n"
        ".reg .pred q;\n"
        "setp.ne.b32 q, %6, 0;\n"
        "@q cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n"
        "}\n" ::"r"(smem_u32(smem)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(on)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d_if(uint32_t on, void* smem, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2,
                                               int c3) {
    asm volatile(
        "{\n"
        ".reg .pred q;\n"
        "setp.ne.b32 q, %7, 0;\n"
        "@q cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];\n"
        "}\n" ::"r"(smem_u32(smem)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(on)
        : "memory");
}
__device__ __forceinline__ void tma_load_5d_if(uint32_t on, void* smem, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2,
                                               int c3, int c4) {
    asm volatile(
        "{\n"
        ".reg .pred q;\n"
        "setp.ne.b32 q, %8, 0;\n"
        "@q cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];\n"
        "}\n" ::"r"(smem_u32(smem)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(on)
        : "memory");
}

// ---------------------------------------------------------------------------------------------- tcgen05
template <uint32_t COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem) {     // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "n"(COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <uint32_t COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {        // the same warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem], tf32 inputs (fp32 words in shared memory), fp32 accumulate
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// ---- MMA issue from a CONVERGENT warp -------------------------------------------------------------------------------
// ncu (profiles/r2_h_issue_loop.md): with the issue loop inside `if (lane == 0)` every loop variable lives in vector
// registers, so ptxas wraps each tcgen05.mma in an ELECT / R2UR.BROADCAST "waterfall" loop to move the TMEM address into a
// uniform register and rebuilds both descriptors with shift / mask / or chains: ~13 SASS instructions and a branch per MMA.
// At N = 64 an MMA occupies the tensor pipe for ~53 cycles but the issuing thread needed ~95 — the 64-wide layers were bound
// by instruction issue, not by tensor, L2 or shared-memory bandwidth.  Fix: all 32 lanes of the MMA warp walk the loop
// (warp-uniform control flow: ptxas keeps the loop state, the descriptors and the TMEM addresses in UNIFORM registers), the
// TMEM base goes through a warp reduction (REDUX writes a uniform register), descriptors are advanced with one integer add
// on their low word, and only an elected lane executes the tcgen05 instructions: 3.4 instructions per MMA.
__device__ __forceinline__ uint32_t elect_one() {          // 1 in exactly one lane of the (fully active) warp
    uint32_t pred = 0;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "elect.sync _|p, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(pred));
    return pred;
}
__device__ __forceinline__ uint32_t warp_uniform(uint32_t v) { return __reduce_or_sync(0xffffffffu, v); }   // same v in all lanes
// Descriptor words: the high word (stride offset, version, layout type) does not depend on the address; the low word holds
// the start address in 16-byte units in bits [0,14) (+ the leading byte offset in [16,30)), so the descriptor of
// `addr + off` is `lo(addr) + (off >> 4)` — no carry out of the field for any shared-memory address.
__device__ __forceinline__ uint32_t desc_lo(uint64_t d) { return (uint32_t)d; }
__device__ __forceinline__ uint32_t desc_hi(uint64_t d) { return (uint32_t)(d >> 32); }
__device__ __forceinline__ void umma_tf32_words_if(uint32_t on, uint32_t tmem_d, uint32_t lo_a, uint32_t hi_a, uint32_t lo_b,
                                                   uint32_t hi_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p, q;\n"
        ".reg .b64 da, db;\n"
        "setp.ne.b32 p, %6, 0;\n"
        "setp.ne.b32 q, %7, 0;\n"
        "mov.b64 da, {%1, %2};\n"
        "mov.b64 db, {%3, %4};\n"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, p;\n"
        "}\n" ::"r"(tmem_d),
        "r"(lo_a), "r"(hi_a), "r"(lo_b), "r"(hi_b), "r"(idesc), "r"(accumulate), "r"(on)
        : "memory");
}
// arrive on an mbarrier when all previously issued MMAs of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void umma_commit_if(uint32_t on, uint64_t* bar) {
    asm volatile(
        "{\n"
        ".reg .pred q;\n"
        "setp.ne.b32 q, %1, 0;\n"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(on)
        : "memory");
}
// 32 lanes x 32 columns of fp32: thread t of the warp gets lane (lane_base + t), columns col .. col+31
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// UMMA shared-memory matrix descriptor, K-major operand, 128-byte swizzle: rows of 128 B (32 fp32), 8-row
// swizzle atoms of 1024 B stacked along M/N (stride byte offset 1024), base 1024-B aligned.
__device__ __forceinline__ uint64_t umma_desc_k128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);        // start address,      bits [0,14)
    d |= (uint64_t)1 << 16;                             // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                   // stride byte offset, bits [32,46)
    d |= (uint64_t)1 << 46;                             // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                             // layout type: SWIZZLE_128B
    return d;
}

// K-major operand, 32-byte swizzle: rows of 32 B (8 fp32 = one UMMA_K step of tf32), 8-row atoms of 256 B stacked along
// M/N.  What TMA writes for a box whose inner dimension is 8 floats under CU_TENSOR_MAP_SWIZZLE_32B (tools/probe/tma_probe.cu:
// under the 128-byte modes such a box is padded to one 128-byte line per inner row instead).
__device__ __forceinline__ uint64_t umma_desc_k32(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;                             // leading byte offset (unused: K extent = one atom)
    d |= (uint64_t)(256 >> 4) << 32;                    // stride byte offset: 8 rows x 32 B
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)6 << 61;                             // layout type: SWIZZLE_32B
    return d;
}

// MN-major operand, 128-byte swizzle: the tile is stored as [mn block of 32][k rows][32 fp32 along M/N]; one
// swizzle atom = 8 k-rows x 128 B.  Leading byte offset = distance between consecutive 32-wide M/N blocks,
// stride byte offset = distance between consecutive groups of 8 k-rows.
__device__ __forceinline__ uint64_t umma_desc_mn128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                     uint32_t layout_type = 1) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)layout_type << 61;                   // 1 = SWIZZLE_128B with 32-byte atoms (32-bit MN-major)
    return d;
}

// instruction descriptor for kind::tf32: D = fp32, A/B = tf32, M x N tile; *_mn = operand is M/N-major (transposed)
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N, bool a_mn = false, bool b_mn = false) {
    return (1u << 4)                       // c_format = F32
           | (2u << 7)                     // a_format = TF32
           | (2u << 10)                    // b_format = TF32
           | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16)
           | ((uint32_t)(N >> 3) << 17)    // n_dim
           | ((uint32_t)(M >> 4) << 24);   // m_dim
}

// ---------------------------------------------------------------------------------------------- epilogue: BN statistics
// Per-channel sum / sum of squares of a conv OUTPUT accumulated by the epilogue that already holds the values in
// registers (SURVEY §5 / §7 step 7 "BN statistics in the conv epilogue"; replaces a separate pass over the tensor).
// v[32]: this lane's pixel, 32 consecutive channels (zeros for pixels outside the tensor).  A transposing butterfly
// (31 shuffles per quantity) leaves lane l with the warp's sum for channel l; the four epilogue warps fold into a CTA
// shared-memory pair sm_stats[2][BN] and the CTA flushes it with fp64 global atomics once per work item.
__device__ __forceinline__ void warp_channel_sums(const float (&v)[32], float& s, float& q) {
    float a[32], b[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) { a[i] = v[i]; b[i] = v[i] * v[i]; }
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int ofs = 16; ofs >= 1; ofs >>= 1) {
        const bool up = (lane & ofs) != 0;
#pragma unroll
        for (int i = 0; i < ofs; ++i) {
            const float sa = up ? a[i] : a[i + ofs], ka = up ? a[i + ofs] : a[i];
            const float sb = up ? b[i] : b[i + ofs], kb = up ? b[i + ofs] : b[i];
            a[i] = ka + __shfl_xor_sync(0xffffffffu, sa, ofs);
            b[i] = kb + __shfl_xor_sync(0xffffffffu, sb, ofs);
        }
    }
    s = a[0];
    q = b[0];
}
// epilogue warps: add this chunk's 32 channel sums into sm_stats[0][c..c+32) / sm_stats[1][...]  (BN = row length)
__device__ __forceinline__ void stats_accumulate(const float (&v)[32], bool valid, float* sm_stats, int BN_, int c) {
    float z[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) z[i] = valid ? v[i] : 0.f;
    float s, q;
    warp_channel_sums(z, s, q);
    const int lane = threadIdx.x & 31;
    atomicAdd(sm_stats + c + lane, s);
    atomicAdd(sm_stats + BN_ + c + lane, q);
}
// sums only (bias gradient of a fused activation adjoint): half the shuffles
__device__ __forceinline__ void stats_accumulate_sum(const float (&v)[32], bool valid, float* sm_stats, int c) {
    float a[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) a[i] = valid ? v[i] : 0.f;
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int ofs = 16; ofs >= 1; ofs >>= 1) {
        const bool up = (lane & ofs) != 0;
#pragma unroll
        for (int i = 0; i < ofs; ++i) {
            const float sa = up ? a[i] : a[i + ofs], ka = up ? a[i + ofs] : a[i];
            a[i] = ka + __shfl_xor_sync(0xffffffffu, sa, ofs);
        }
    }
    atomicAdd(sm_stats + c + lane, a[0]);
}
// Activation adjoint fused into an input-gradient epilogue: v[i] *= (m[i] >= 0 ? 1 : slope), m = the 32 activated forward
// values at this lane's output pixel (LeakyReLU keeps the sign, so the activated tensor is its own mask; the same rule as
// pad_leaky_bias_bwd_kernel in ew_kernels.cu).  The epilogue warps fetch the signs of a whole work item as bit words
// BEFORE they wait for the accumulator (the loads overlap the item's MMA main loop instead of sitting between tcgen05.ld
// and the stores).  `vec`: all 32 channels exist and are 16-byte aligned; channels >= nch read as "pass".
__device__ __forceinline__ uint32_t act_mask_bits32(const float* m, bool vec, int nch) {
    uint32_t bits = 0;
    if (vec) {
        float4 t[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) t[i] = __ldg(reinterpret_cast<const float4*>(m) + i);
#pragma unroll
        for (int i = 0; i < 8; ++i)
            bits |= ((t[i].x >= 0.f ? 1u : 0u) | (t[i].y >= 0.f ? 2u : 0u) | (t[i].z >= 0.f ? 4u : 0u) | (t[i].w >= 0.f ? 8u : 0u)) << (4 * i);
    } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) bits |= ((i >= nch || __ldg(m + (i < nch ? i : 0)) >= 0.f) ? 1u : 0u) << i;
    }
    return bits;
}
template <int NW>
__device__ __forceinline__ uint32_t pick_word(const uint32_t (&w)[NW], int idx) {      // register-resident dynamic index
    uint32_t r = w[0];
#pragma unroll
    for (int k = 1; k < NW; ++k) r = (k == idx) ? w[k] : r;
    return r;
}
__device__ __forceinline__ void apply_act_bits32(float (&v)[32], uint32_t bits, float slope) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = ((bits >> i) & 1u) ? v[i] : v[i] * slope;
}
// one epilogue warp after all four are done with the item (named barrier 1, 128 threads): flush + clear
__device__ __forceinline__ void stats_flush(float* sm_stats, int BN_, double* gstats, int C, int c0, int ep_tid) {
    asm volatile("bar.sync 1, 128;" ::: "memory");
    for (int i = ep_tid; i < BN_; i += 128) {
        if (c0 + i < C) {
            atomicAdd(gstats + c0 + i, (double)sm_stats[i]);
            atomicAdd(gstats + C + c0 + i, (double)sm_stats[BN_ + i]);
        }
        sm_stats[i] = 0.f;
        sm_stats[BN_ + i] = 0.f;
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
}

// ---------------------------------------------------------------------------------------------- host: tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// fp32 tensor, `rank` dims listed innermost first; strides in BYTES for dims 1..rank-1; 128-byte swizzle
inline int make_tmap_f32(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_b,
                         const uint32_t* box, const uint32_t* elem_strides = nullptr,
                         CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) {
        b3d::set_error("cuTensorMapEncodeTiled is not available from the driver");
        return B3D_ECUDA;
    }
    cuuint64_t gd[5], gs[5];
    cuuint32_t bx[5], es[5];
    for (int i = 0; i < rank; ++i) {
        gd[i] = dims[i];
        bx[i] = box[i];
        es[i] = elem_strides ? elem_strides[i] : 1;
        if (i > 0) gs[i - 1] = strides_b[i - 1];
    }
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        b3d::set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d, inner box %u)", (int)r, rank, box[0]);
        return B3D_ECUDA;
    }
    return B3D_OK;
}

}  // namespace tc
