// tcgen05 implicit-GEMM 3x3 convolution for sm_100a (the Block convs: 81.5 % of the step's MACs).
//
//   D[pixel][cout] (fp32, TMEM) += A[pixel][tap, cin] (smem, tf32|bf16) * W[cout][tap, cin] (smem, tf32|bf16)
//
// Mapping.  One CTA owns an output tile of ROWS=2 mel rows x 128 frames x NT output channels of one
// sample.  M = 128 consecutive frames of one row is one UMMA (M=128, N=NT, K=32 bytes); the two rows use two
// TMEM accumulators (2*NT columns) and share every weight stage.  K runs over (input-channel stage, 9 taps).
//
// A operand: the activations are NHWC fp32 in HBM and must be normalised (GroupNorm apply), activated (Mish),
// masked and biased by the time projection before the conv (diffusion.py:56-58,76) - so TMA cannot stage them.
// Eight producer warps load the (ROWS+2) x 130-pixel halo of a channel stage once, apply that prologue in
// registers, round to tf32/bf16 and store it to shared memory in the UMMA "interleaved" (no-swizzle) K-major
// layout: [16-byte channel chunk][halo row][pixel][16 B].  In that layout 8 consecutive pixels x 16 B form one
// core matrix, so EVERY one of the 9 taps is just a different start address into the same halo tile
// (start += (r*130 + s)*16 B): one load + one transform per input element, nine MMAs.
//
// B operand: weights are packed on the host into exactly the per-stage shared-memory image
// [tap][chunk][cout][16 B] and streamed with one cp.async.bulk per stage (mbarrier complete_tx).
//
// Pipeline: STAGES-deep ring of {A halo, B weights} with full_a/full_b/empty mbarriers; a single thread issues
// tcgen05.mma and releases stages with tcgen05.commit; the epilogue (same 8 warps) reads the accumulators with
// tcgen05.ld, adds the bias, writes NHWC fp32 and accumulates the GroupNorm {sum, sumsq} of the raw output.
#include "sbk_internal.h"

#include <cuda_bf16.h>
#include <math.h>
#include <stdint.h>

namespace sbk {

namespace tc {

constexpr int ROWS = 2;               // output mel rows per CTA
constexpr int HR = ROWS + 2;          // halo rows
constexpr int TPX = 128;              // output frames per CTA row (= UMMA M)
constexpr int PXP = TPX + 2;          // halo pixels per row
constexpr int KCH = 2;                // 16-byte K chunks per stage (= one UMMA K step of 32 bytes)
constexpr int STAGES = 3;
constexpr int NPROD = 256;            // producer / epilogue threads (8 warps)
constexpr int NTHREADS = NPROD + 64;  // + MMA warp + weight-loader warp
constexpr int A_STAGE_BYTES = KCH * HR * PXP * 16;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// bounded spin: a protocol bug must trap, not hang the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 24)) __trap();
    }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// shared-memory matrix descriptor, K-major, SWIZZLE_NONE ("interleaved"): element (row m, 16-byte K chunk c)
// lives at start + (m%8)*16 + (m/8)*SBO + c*LBO  (cute/arch/mma_sm100_desc.hpp SmemDescriptor; version_=1).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;   // version = 1 (Blackwell)
    return d;                 // base_offset = 0, lbo_mode = 0, layout_type = SWIZZLE_NONE (0)
}

// instruction descriptor (UMMA::InstrDescriptor): c=F32, a/b format, K-major both, N>>3 at [17,23), M>>4 at [24,29)
template <bool BF16>
__device__ __forceinline__ uint32_t make_idesc(int M, int N) {
    uint32_t d = 0;
    d |= 1u << 4;                           // c_format = F32
    d |= (BF16 ? 1u : 2u) << 7;             // a_format: BF16 = 1, TF32 = 2
    d |= (BF16 ? 1u : 2u) << 10;            // b_format
    d |= (uint32_t)(N >> 3) << 17;
    d |= (uint32_t)(M >> 4) << 24;
    return d;
}

template <bool BF16>
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    if (BF16) {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
    } else {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
    }
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ float mish_fast(float x) {
    // same closed form as mish_f (sbk_kernels.cu); exp via ex2.approx and an approximate reciprocal:
    // relative error ~1e-6, far below the tf32/bf16 operand rounding this path already applies.
    const float n = __expf(fminf(x, 20.f));
    const float a = n * (n + 2.f);
    return x > 20.f ? x : x * __fdividef(a, a + 2.f);
}

__device__ __forceinline__ uint32_t to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}

}  // namespace tc

using namespace tc;

// channels consumed per pipeline stage: KCH chunks x (4 tf32 | 8 bf16) elements
template <bool BF16> struct StageCh { static constexpr int value = KCH * (BF16 ? 8 : 4); };

template <bool BF16, int NT>
__global__ void __launch_bounds__(NTHREADS, 1) k_conv3x3_tc(const ConvTcParams p) {
    constexpr int CPS = StageCh<BF16>::value;              // input channels per stage
    constexpr int F4 = CPS / 4;                            // float4 loads per halo pixel per stage
    constexpr int B_STAGE_BYTES = 9 * KCH * NT * 16;
    constexpr uint32_t TMEM_COLS = ROWS * NT;              // 128 or 256: a power of two >= 32

    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sA = smem;                                            // [STAGES][KCH][HR][PXP][16]
    uint8_t* sB = sA + STAGES * A_STAGE_BYTES;                     // [STAGES][9][KCH][NT][16]
    float* s_tab = reinterpret_cast<float*>(sB + STAGES * B_STAGE_BYTES);   // mean|scale|beta|tb : 4*Cin
    const int Cin = p.c0 + p.c1;
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_tab + 4 * Cin);         // full_a[S], full_b[S], empty[S], acc
    float* s_st = reinterpret_cast<float*>(bars + 3 * STAGES + 1);         // [8 groups][2]
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_st + 16);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wtiles = (p.W + TPX - 1) / TPX;
    const int w0 = (blockIdx.x % wtiles) * TPX;
    const int h0 = (blockIdx.x / wtiles) * ROWS;
    const int n0 = blockIdx.y * NT;
    const int b = blockIdx.z;
    const int ksteps = Cin / CPS;

    const uint32_t bar0 = smem_u32(bars);
    auto full_a = [&](int s) { return bar0 + 8u * s; };
    auto full_b = [&](int s) { return bar0 + 8u * (STAGES + s); };
    auto empty = [&](int s) { return bar0 + 8u * (2 * STAGES + s); };
    const uint32_t acc_bar = bar0 + 8u * (3 * STAGES);

    // ---- one-time setup
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(full_a(s), NPROD / 32); mbar_init(full_b(s), 1); mbar_init(empty(s), 1); }
        mbar_init(acc_bar, 1);
        fence_barrier_init();
    }
    if (warp == NPROD / 32) tmem_alloc(smem_u32(s_tmem), TMEM_COLS);
    if (tid < 16) s_st[tid] = 0.f;
    // prologue tables (GroupNorm apply + time projection of the producing Block), all Cin channels
    {
        float* mean = s_tab; float* scale = mean + Cin; float* beta = scale + Cin; float* tbv = beta + Cin;
        if (p.pro == PRO_GN) {
            const int cpg = Cin / kGroups;
            const int row = p.tb_per_sample ? b : *p.step;
            const float* tb = p.tb + (long long)row * p.tb_stride;
            for (int c = tid; c < Cin; c += NTHREADS) {
                const int g = c / cpg;
                const double s = p.pgn.stats[(b * kGroups + g) * 2], ss = p.pgn.stats[(b * kGroups + g) * 2 + 1];
                const double m = s * (double)p.pgn.inv_count;
                double var = ss * (double)p.pgn.inv_count - m * m;
                var = var < 0.0 ? 0.0 : var;
                mean[c] = (float)m;
                scale[c] = (float)(1.0 / sqrt(var + 1e-5)) * p.pgn.gamma[c];
                beta[c] = p.pgn.beta[c];
                tbv[c] = tb[c];
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *s_tmem;

    if (warp < NPROD / 32) {
        // =============================== A producers ===============================
        // item it -> halo (row r, pixel q); each item = CPS channels = F4 float4 loads = KCH 16-byte smem chunks
        constexpr int ITEMS = HR * PXP;                       // 520
        constexpr int PER = (ITEMS + NPROD - 1) / NPROD;      // 3 (last round: 8 threads)
        const float* mean = s_tab; const float* scale = mean + Cin; const float* beta = scale + Cin; const float* tbv = beta + Cin;
        int it_r[PER], it_q[PER]; bool it_ok[PER]; float it_mask[PER]; long long it_off[PER];
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            const int it = tid + j * NPROD;
            const bool in = it < ITEMS;
            const int r = in ? it / PXP : 0, q = in ? it - r * PXP : 0;
            const int hi = h0 - 1 + r, wi = w0 - 1 + q;
            it_r[j] = r; it_q[j] = q;
            it_ok[j] = in && hi >= 0 && hi < p.H && wi >= 0 && wi < p.W;
            it_mask[j] = (it_ok[j] && p.pro != PRO_NONE) ? __ldg(p.mask + (long long)b * p.T + ((long long)wi << p.lvl)) : (it_ok[j] ? 1.f : 0.f);
            it_off[j] = ((long long)(b * p.H + hi) * p.W + wi);
            if (!in) it_r[j] = -1;
        }
        float4 cur[PER][F4], nxt[PER][F4];
        auto load = [&](int ks, float4 (&dst)[PER][F4]) {
            const int cc = ks * CPS;
            const bool second = cc >= p.c0;
            const float* src = second ? p.in1 : p.in0;
            const int cs = second ? p.c1 : p.c0;
            const int co = second ? cc - p.c0 : cc;
#pragma unroll
            for (int j = 0; j < PER; ++j) {
                if (it_ok[j] && it_mask[j] != 0.f) {
                    const float* g = src + it_off[j] * cs + co;
#pragma unroll
                    for (int f = 0; f < F4; ++f) dst[j][f] = __ldg(reinterpret_cast<const float4*>(g) + f);
                } else {
#pragma unroll
                    for (int f = 0; f < F4; ++f) dst[j][f] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
        };
        load(0, cur);
        for (int ks = 0; ks < ksteps; ++ks) {
            const int s = ks % STAGES;
            if (ks + 1 < ksteps) load(ks + 1, nxt);
            mbar_wait(empty(s), ((ks / STAGES) & 1) ^ 1);
            uint8_t* stage = sA + s * A_STAGE_BYTES;
            const int cc = ks * CPS;
#pragma unroll
            for (int j = 0; j < PER; ++j) {
                if (it_r[j] < 0) continue;
                float v[CPS];
#pragma unroll
                for (int f = 0; f < F4; ++f) { v[4 * f] = cur[j][f].x; v[4 * f + 1] = cur[j][f].y; v[4 * f + 2] = cur[j][f].z; v[4 * f + 3] = cur[j][f].w; }
                const bool live = it_ok[j] && it_mask[j] != 0.f;
                if (p.pro == PRO_GN) {
#pragma unroll
                    for (int e = 0; e < CPS; ++e) {
                        const int c = cc + e;
                        v[e] = live ? mish_fast((v[e] - mean[c]) * scale[c] + beta[c]) + tbv[c] : 0.f;
                    }
                }   // PRO_MASK / PRO_NONE: masked or out-of-range pixels were loaded as zeros, mask is {0,1}
                uint8_t* dst = stage + ((0 * HR + it_r[j]) * PXP + it_q[j]) * 16;
#pragma unroll
                for (int k = 0; k < KCH; ++k) {
                    uint4 w;
                    if (BF16) {
                        __nv_bfloat162 h0v = __floats2bfloat162_rn(v[8 * k + 0], v[8 * k + 1]);
                        __nv_bfloat162 h1v = __floats2bfloat162_rn(v[8 * k + 2], v[8 * k + 3]);
                        __nv_bfloat162 h2v = __floats2bfloat162_rn(v[8 * k + 4], v[8 * k + 5]);
                        __nv_bfloat162 h3v = __floats2bfloat162_rn(v[8 * k + 6], v[8 * k + 7]);
                        w.x = *reinterpret_cast<uint32_t*>(&h0v); w.y = *reinterpret_cast<uint32_t*>(&h1v);
                        w.z = *reinterpret_cast<uint32_t*>(&h2v); w.w = *reinterpret_cast<uint32_t*>(&h3v);
                    } else {
                        w.x = to_tf32(v[4 * k + 0]); w.y = to_tf32(v[4 * k + 1]); w.z = to_tf32(v[4 * k + 2]); w.w = to_tf32(v[4 * k + 3]);
                    }
                    *reinterpret_cast<uint4*>(dst + k * (HR * PXP * 16)) = w;
                }
            }
            fence_proxy_async();                 // generic-proxy smem writes -> visible to the tensor core (async proxy)
            __syncwarp();
            if (lane == 0) mbar_arrive(full_a(s));
#pragma unroll
            for (int j = 0; j < PER; ++j)
#pragma unroll
                for (int f = 0; f < F4; ++f) cur[j][f] = nxt[j][f];
        }

        // =============================== epilogue ===============================
        mbar_wait(acc_bar, 0);
        tc_fence_after();
        const int q4 = warp & 3, jrow = warp >> 2;            // TMEM lane quarter / accumulator (output row)
        const int px = q4 * 32 + lane;
        const int ho = h0 + jrow, wo = w0 + px;
        const bool valid = ho < p.H && wo < p.W;
        const int cpg = p.Cout / kGroups;
        float* op = p.out + ((long long)(b * p.H + ho) * p.W + wo) * p.Cout + n0;
#pragma unroll 1
        for (int cb = 0; cb < NT; cb += 32) {
            uint32_t r[32];
            tmem_ld32(tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(jrow * NT + cb), r);
            float v[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]) + __ldg(p.bias + n0 + cb + i);
            if (valid) {
#pragma unroll
                for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(op + cb + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
            }
            if (p.ostats) {
                // GroupNorm partials of this 32-column chunk: 8-channel sub-sums first (static register indexing),
                // then merged to the group width cpg (8 -> 4 groups, 16 -> 2 groups, >= 32 -> 1 group)
                float s8[4], q8[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    float s = 0.f, q = 0.f;
#pragma unroll
                    for (int i = 0; i < 8; ++i) { const float x = valid ? v[8 * k + i] : 0.f; s += x; q = fmaf(x, x, q); }
                    s8[k] = s; q8[k] = q;
                }
                const int ngrp = cpg == 8 ? 4 : (cpg == 16 ? 2 : 1);
                if (ngrp == 2) { s8[0] += s8[1]; q8[0] += q8[1]; s8[1] = s8[2] + s8[3]; q8[1] = q8[2] + q8[3]; }
                if (ngrp == 1) { s8[0] += s8[1] + s8[2] + s8[3]; q8[0] += q8[1] + q8[2] + q8[3]; }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (k < ngrp) {
                        float s = s8[k], q = q8[k];
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
                        if (lane == 0) {
                            const int gl = (n0 + cb + k * (32 / ngrp)) / cpg - n0 / cpg;
                            atomicAdd(&s_st[gl * 2], s);
                            atomicAdd(&s_st[gl * 2 + 1], q);
                        }
                    }
                }
            }
        }
        tc_fence_before();
    } else if (warp == NPROD / 32) {
        // =============================== MMA issuer ===============================
        if (lane == 0) {
            const uint32_t idesc = make_idesc<BF16>(TPX, NT);
            const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
            for (int ks = 0; ks < ksteps; ++ks) {
                const int s = ks % STAGES;
                const uint32_t ph = (ks / STAGES) & 1;
                mbar_wait(full_a(s), ph);
                mbar_wait(full_b(s), ph);
                tc_fence_after();
#pragma unroll
                for (int tap = 0; tap < 9; ++tap) {
                    const int r = tap / 3, sx = tap % 3;
                    const uint64_t bd = p.dbg_swap ? make_desc(b0 + s * B_STAGE_BYTES + tap * (KCH * NT * 16), 128, NT * 16)
                                                   : make_desc(b0 + s * B_STAGE_BYTES + tap * (KCH * NT * 16), NT * 16, 128);
#pragma unroll
                    for (int j = 0; j < ROWS; ++j) {
                        const uint32_t aaddr = a0 + s * A_STAGE_BYTES + ((r + j) * PXP + sx) * 16;
                        const uint64_t ad = p.dbg_swap ? make_desc(aaddr, 128, HR * PXP * 16) : make_desc(aaddr, HR * PXP * 16, 128);
                        umma<BF16>(tmem_base + j * NT, ad, bd, idesc, (ks | tap) != 0 ? 1u : 0u);
                    }
                }
                umma_commit(empty(s));                  // frees the stage when these MMAs have read it
            }
            umma_commit(acc_bar);                       // accumulators complete
        }
    } else {
        // =============================== weight loader ===============================
        if (lane == 0) {
            const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(p.wpk) + (size_t)blockIdx.y * ksteps * B_STAGE_BYTES;
            for (int ks = 0; ks < ksteps; ++ks) {
                const int s = ks % STAGES;
                mbar_wait(empty(s), ((ks / STAGES) & 1) ^ 1);
                mbar_arrive_expect_tx(full_b(s), B_STAGE_BYTES);
                bulk_g2s(smem_u32(sB + s * B_STAGE_BYTES), wsrc + (size_t)ks * B_STAGE_BYTES, B_STAGE_BYTES, full_b(s));
            }
        }
    }

    __syncthreads();
    if (p.ostats) {
        const int cpg = p.Cout / kGroups, gb = n0 / cpg, ng = (NT + cpg - 1) / cpg;
        if (tid < ng * 2) atomicAdd(&p.ostats[((long long)b * kGroups + gb + (tid >> 1)) * 2 + (tid & 1)], (double)s_st[tid]);
    }
    if (warp == NPROD / 32) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

template <bool BF16, int NT>
static size_t conv_tc_smem(int Cin) {
    return (size_t)STAGES * (A_STAGE_BYTES + 9 * KCH * NT * 16) + 4 * (size_t)Cin * sizeof(float) + (3 * STAGES + 1) * 8 + 16 * 4 + 16;
}

template <bool BF16, int NT>
static int launch_tc(const ConvTcParams& p, cudaStream_t s) {
    const size_t sm = conv_tc_smem<BF16, NT>(p.c0 + p.c1);
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(k_conv3x3_tc<BF16, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        attr_set = true;
    }
    const int wt = (p.W + TPX - 1) / TPX, ht = (p.H + ROWS - 1) / ROWS;
    dim3 grid(wt * ht, p.Cout / NT, p.B);
    k_conv3x3_tc<BF16, NT><<<grid, NTHREADS, sm, s>>>(p);
    return 1;
}

int conv_tc_ntile(int Cout) { return Cout % 128 == 0 ? 128 : 64; }
int conv_tc_stage_channels(int bf16) { return bf16 ? StageCh<true>::value : StageCh<false>::value; }

int launch_conv_tc(const ConvTcParams& p, cudaStream_t s) {
    const int nt = conv_tc_ntile(p.Cout);
    if (p.bf16) return nt == 128 ? launch_tc<true, 128>(p, s) : launch_tc<true, 64>(p, s);
    return nt == 128 ? launch_tc<false, 128>(p, s) : launch_tc<false, 64>(p, s);
}

}  // namespace sbk
