// tcgen05 projection kernel (K-B, tensor-core path):
//   C[M,N] = act(A[M,K] * W[N,K]^T + bias[N])      (nn.Linear as built by build_mlp,
//                                                   pyrutils/torch/models.py:31-33)
// fp32 in, fp32 out, fp32-class accuracy through the 3xTF32 split on the 5th-generation tensor cores:
//   x = hi + lo (hi: top 19 bits, exactly TF32; lo = x - hi), D += A_lo W_hi + A_hi W_lo + A_hi W_hi,
// accumulated in fp32 in TMEM.  Relative error per product ~2^-21, i.e. the same class as fp32 FFMA with a
// different summation order, which the discrete segmentation gates downstream require.
//
// CTA = one 128 x 128 output tile, 288 threads:
//   warps 0-7  producers, later epilogue.  Per k-block (32 floats = one 128-byte swizzle row) they load
//              A/W rows from global (coalesced LDG.128, prefetched two k-blocks ahead in registers), split into
//              hi/lo and store both tiles in the canonical K-major SWIZZLE_128B layout the UMMA descriptors
//              expect; fence.proxy.async + mbarrier arrive hands the stage to the tensor core.
//   warp 8     allocates TMEM (128 fp32 columns) and one elected lane issues 12 tcgen05.mma (kind::tf32,
//              M=128,N=128,K=8: 4 k-steps x 3 split products) per k-block; tcgen05.commit frees the stage.
//   epilogue   tcgen05.ld 32x32b.x32 (lane = output row), bias + ReLU, float4 stores.
// Several independent problems are batched into one launch like the SIMT path.
#include <stdlib.h>
#include "common.cuh"
#include "gemm.h"
#include "tcgen05.cuh"
#include <cuda_bf16.h>

namespace tg {

constexpr int TC_BM = 128, TC_BK = 32;
constexpr int TC_PRODUCER_WARPS = 8;
constexpr int TC_THREADS = (TC_PRODUCER_WARPS + 1) * 32;
constexpr int TC_TILE_BYTES = TC_BM * TC_BK * 4;                  // 16 KB, one 128-row operand tile (hi or lo)
// Output tile 128 x BN.  BN = 256 (large problems): a W tile is two 128-row tiles, the stage grows to 96 KB (two stages) and
// the fp32 operand bytes read from L2 per FLOP fall by a quarter — the kernel sits at the L2 -> SM bandwidth, not the MMA rate.
template <int BN> struct TcCfg {
    static constexpr int W_TILE_BYTES = BN * TC_BK * 4;
    static constexpr int STAGE_BYTES = 2 * TC_TILE_BYTES + 2 * W_TILE_BYTES;     // A_hi, A_lo, W_hi, W_lo
    static constexpr int STAGES = BN == 128 ? 3 : 2;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024;              // + alignment slack
    static constexpr int WP = BN / 32;                                           // W rows per producer thread
};

// kind::f16 instruction descriptor with bf16 operands: D = F32, A = B = BF16 (format 1), both K-major
__device__ __forceinline__ uint32_t umma_idesc_bf16(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ uint2 pack_bf16x4(const float4& x) {
    const __nv_bfloat162 lo = __floats2bfloat162_rn(x.x, x.y), hi = __floats2bfloat162_rn(x.z, x.w);
    uint2 r;
    r.x = *reinterpret_cast<const uint32_t*>(&lo);
    r.y = *reinterpret_cast<const uint32_t*>(&hi);
    return r;
}

// MASK: some problem of the group reads A through a ReLU mask (backward use); compiled out of the forward instantiation.
// PREC 0: 3xTF32 split (hi and lo tiles, 12 MMAs per k-block).  PREC 1 (dims.precision = 1): the producers round the fp32
// operands to bf16 and fill the FIRST 64 bytes of each 128-byte swizzle row of the hi tiles (32 K-elements per k-block as before),
// the issuer runs two kind::f16 MMAs (K = 16) per k-block: one sixth of the tensor-pipe work, half the shared-memory bytes.
template <bool MASK, int BN, int PREC>
__global__ void __launch_bounds__(TC_THREADS, 1) gemm_tc_kernel(const GemmGroup grp) {
    constexpr int TC_BN = BN, TC_STAGES = TcCfg<BN>::STAGES, TC_STAGE_BYTES = TcCfg<BN>::STAGE_BYTES, WP = TcCfg<BN>::WP;
    constexpr int W_TILE_BYTES = TcCfg<BN>::W_TILE_BYTES;
    constexpr uint32_t TC_TMEM_COLS = BN;
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[2 * TC_STAGES + 1];
    __shared__ uint32_t tmem_base_smem;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t tiles_u32 = smem_u32(tiles);
    const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[TC_STAGES]), tfull = smem_u32(&bars[2 * TC_STAGES]);

    // locate this CTA's problem and tile
    int pi = 0;
#pragma unroll 1
    for (int i = 1; i < grp.count; ++i)
        if ((int)blockIdx.x >= grp.p[i].tile_begin) pi = i;
    const GemmProblem& P = grp.p[pi];
    int tile = blockIdx.x - P.tile_begin;
    const int ks = tile % P.ksplit;                     // split-K slice of this CTA
    tile /= P.ksplit;
    const int tiles_n = (P.N + TC_BN - 1) / TC_BN;
    const int m0 = (tile / tiles_n) * TC_BM, n0 = (tile % tiles_n) * TC_BN;
    const int nkb_all = P.K / TC_BK;
    const int kb0 = (int)((long long)nkb_all * ks / P.ksplit);
    const int nkb = (int)((long long)nkb_all * (ks + 1) / P.ksplit) - kb0;

    if (tid == 0) {
        for (int s = 0; s < TC_STAGES; ++s) {
            mbar_init(full0 + 8 * s, TC_PRODUCER_WARPS);
            mbar_init(empty0 + 8 * s, 1);
        }
        mbar_init(tfull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == TC_PRODUCER_WARPS) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_smem)), "r"(TC_TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;

    if (warp < TC_PRODUCER_WARPS) {
        // ------------------------------ producers ------------------------------
        // thread -> 16-byte chunk c of rows r0 + 32*i (i < 4) of the A tile and of the W tile
        const int c = tid & 7, r0 = tid >> 3;                 // 256 threads: r0 in [0,32)
        const float* aptr[4];
        const float* wptr[WP];
        const float* mptr[4];
        bool aok[4], wok[WP];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = r0 + 32 * i;
            aok[i] = (m0 + r) < P.M;
            aptr[i] = P.A + (size_t)(aok[i] ? m0 + r : 0) * P.lda + c * 4 + (size_t)kb0 * TC_BK;
            mptr[i] = (MASK && P.amask != nullptr) ? P.amask + (size_t)(aok[i] ? m0 + r : 0) * P.ldm + c * 4 + (size_t)kb0 * TC_BK : nullptr;
        }
#pragma unroll
        for (int i = 0; i < WP; ++i) {
            const int r = r0 + 32 * i;
            wok[i] = (n0 + r) < P.N;
            wptr[i] = P.W + (size_t)(wok[i] ? n0 + r : 0) * P.ldw + c * 4 + (size_t)kb0 * TC_BK;
        }
        float4 pa0[4], pw0[WP], pa1[4], pw1[WP];              // register prefetch, two k-blocks deep (static slots)
        auto load = [&](int kb, float4 (&pa)[4], float4 (&pw)[WP]) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                pa[i] = aok[i] ? __ldg(reinterpret_cast<const float4*>(aptr[i] + (size_t)kb * TC_BK)) : make_float4(0.f, 0.f, 0.f, 0.f);
                if (MASK && mptr[i] != nullptr && aok[i]) {
                    const float4 mk = __ldg(reinterpret_cast<const float4*>(mptr[i] + (size_t)kb * TC_BK));
                    pa[i].x = mk.x > 0.f ? pa[i].x : 0.f; pa[i].y = mk.y > 0.f ? pa[i].y : 0.f;
                    pa[i].z = mk.z > 0.f ? pa[i].z : 0.f; pa[i].w = mk.w > 0.f ? pa[i].w : 0.f;
                }
            }
#pragma unroll
            for (int i = 0; i < WP; ++i)
                pw[i] = wok[i] ? __ldg(reinterpret_cast<const float4*>(wptr[i] + (size_t)kb * TC_BK)) : make_float4(0.f, 0.f, 0.f, 0.f);
        };
        auto split4 = [](const float4& x, float4& hi, float4& lo) {
            hi.x = __uint_as_float(__float_as_uint(x.x) & 0xffffe000u); lo.x = x.x - hi.x;
            hi.y = __uint_as_float(__float_as_uint(x.y) & 0xffffe000u); lo.y = x.y - hi.y;
            hi.z = __uint_as_float(__float_as_uint(x.z) & 0xffffe000u); lo.z = x.z - hi.z;
            hi.w = __uint_as_float(__float_as_uint(x.w) & 0xffffe000u); lo.w = x.w - hi.w;
        };
        auto produce = [&](int kb, float4 (&pa)[4], float4 (&pw)[WP]) {
            const int s = kb % TC_STAGES;
            mbar_wait(empty0 + 8 * s, ((kb / TC_STAGES) & 1) ^ 1);
            uint8_t* st = tiles + s * TC_STAGE_BYTES;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (PREC == 1) {     // float4 chunk c = bf16 elements 4c..4c+3 = bytes 8c..8c+7 of the row: 16-byte chunk c / 2
                    *reinterpret_cast<uint2*>(st + sw128_off(r0 + 32 * i, c >> 1) + (c & 1) * 8) = pack_bf16x4(pa[i]);
                    continue;
                }
                const uint32_t off = sw128_off(r0 + 32 * i, c);
                float4 hi, lo;
                split4(pa[i], hi, lo);
                *reinterpret_cast<float4*>(st + off) = hi;
                *reinterpret_cast<float4*>(st + TC_TILE_BYTES + off) = lo;
            }
#pragma unroll
            for (int i = 0; i < WP; ++i) {                       // W rows r0 + 32 i: the swizzle pattern repeats every 8 rows
                if (PREC == 1) {
                    *reinterpret_cast<uint2*>(st + 2 * TC_TILE_BYTES + sw128_off(r0 + 32 * i, c >> 1) + (c & 1) * 8) = pack_bf16x4(pw[i]);
                    continue;
                }
                const uint32_t off = sw128_off(r0 + 32 * i, c);
                float4 hi, lo;
                split4(pw[i], hi, lo);
                *reinterpret_cast<float4*>(st + 2 * TC_TILE_BYTES + off) = hi;
                *reinterpret_cast<float4*>(st + 2 * TC_TILE_BYTES + W_TILE_BYTES + off) = lo;
            }
            if (kb + 2 < nkb) load(kb + 2, pa, pw);
            asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // generic-proxy writes -> visible to the tensor core
            __syncwarp();
            if (lane == 0) mbar_arrive(full0 + 8 * s);
        };
        if (nkb > 0) load(0, pa0, pw0);
        if (nkb > 1) load(1, pa1, pw1);
#pragma unroll 1
        for (int kb = 0; kb < nkb; kb += 2) {
            produce(kb, pa0, pw0);
            if (kb + 1 < nkb) produce(kb + 1, pa1, pw1);
        }
        // ------------------------------ epilogue ------------------------------
        mbar_wait(tfull, 0);
        tc_fence_after();
        const int q = warp & 3;                                // TMEM lane quarter this warp may access
        const int row = m0 + q * 32 + lane;
        const int col_half = (warp >> 2) * (TC_BN / 2);        // warps 0-3: columns 0..63, warps 4-7: 64..127
#pragma unroll 1
        for (int cb = 0; cb < TC_BN / 2; cb += 32) {
            float v[32];
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(col_half + cb), v);
            const int nbase = n0 + col_half + cb;
            if (row < P.M && nbase < P.N) {
                float* dst = P.C + (size_t)row * P.ldc + nbase;
                if (P.ksplit > 1) {                       // partial sum of one K slice (no bias / activation on this path)
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (nbase + j < P.N) atomicAdd(dst + j, v[j]);
                    continue;
                }
                const bool vec = ((P.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) && (nbase + 32 <= P.N);
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    float x = v[j];
                    if (P.bias != nullptr && nbase + j < P.N) x += __ldg(P.bias + nbase + j);
                    if (P.relu) x = fmaxf(x, 0.0f);
                    if (P.beta && nbase + j < P.N) x += dst[j];
                    v[j] = x;
                }
                if (vec) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (nbase + j < P.N) dst[j] = v[j];
                }
            }
        }
    } else {
        // ------------------------------ MMA issuer ------------------------------
        const uint32_t idesc = PREC == 1 ? umma_idesc_bf16(TC_BM, TC_BN) : umma_idesc_tf32(TC_BM, TC_BN);
#pragma unroll 1
        for (int kb = 0; kb < nkb; ++kb) {
            const int s = kb % TC_STAGES;
            mbar_wait(full0 + 8 * s, (kb / TC_STAGES) & 1);
            tc_fence_after();
            if (lane == 0) {
                const uint32_t a_hi = tiles_u32 + s * TC_STAGE_BYTES, a_lo = a_hi + TC_TILE_BYTES;
                const uint32_t w_hi = a_hi + 2 * TC_TILE_BYTES, w_lo = w_hi + W_TILE_BYTES;
                if (PREC == 1) {
#pragma unroll
                    for (int kk = 0; kk < TC_BK / 16; ++kk)        // 16 bf16 = 32 bytes along the swizzled row
                        umma_bf16(tmem_base, umma_desc(a_hi + kk * 32), umma_desc(w_hi + kk * 32), idesc, (kb | kk) != 0);
                }
#pragma unroll
                for (int kk = 0; kk < (PREC == 1 ? 0 : TC_BK / 8); ++kk) {
                    const uint32_t ko = kk * 32;               // 8 tf32 = 32 bytes along the swizzled row
                    umma_tf32(tmem_base, umma_desc(a_lo + ko), umma_desc(w_hi + ko), idesc, (kb | kk) != 0);
                    umma_tf32(tmem_base, umma_desc(a_hi + ko), umma_desc(w_lo + ko), idesc, 1);
                    umma_tf32(tmem_base, umma_desc(a_hi + ko), umma_desc(w_hi + ko), idesc, 1);
                }
                umma_commit(empty0 + 8 * s);                  // frees the stage when these MMAs have read it
                if (kb == nkb - 1) umma_commit(tfull);        // accumulator complete
            }
            __syncwarp();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == TC_PRODUCER_WARPS) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(TC_TMEM_COLS) : "memory");
    }
}

int launch_gemm_tc(GemmGroup& grp, cudaStream_t stream) {
    if (grp.count == 0) return 0;
    TG_REQUIRE(grp.count <= GEMM_MAX_PROBLEMS, "gemm_tc: too many problems in one group (%d)", grp.count);
    int begin = 0;
    for (int i = 0; i < grp.count; ++i) {
        const GemmProblem& p = grp.p[i];
        TG_REQUIRE(p.M > 0 && p.N > 0 && p.K > 0, "gemm_tc: empty problem");
        TG_REQUIRE(p.K % TC_BK == 0, "gemm_tc: K=%d must be a multiple of %d", p.K, TC_BK);
        TG_REQUIRE(p.lda % 4 == 0 && p.ldw % 4 == 0, "gemm_tc: lda/ldw must be multiples of 4");
        TG_REQUIRE((reinterpret_cast<uintptr_t>(p.A) & 15) == 0 && (reinterpret_cast<uintptr_t>(p.W) & 15) == 0,
                   "gemm_tc: A and W must be 16-byte aligned");
        TG_REQUIRE(p.amask == nullptr || (p.ldm % 4 == 0 && (reinterpret_cast<uintptr_t>(p.amask) & 15) == 0),
                   "gemm_tc: mask must be 16-byte aligned with ldm a multiple of 4");
    }
    // 128 x 256 tiles when they still fill the GPU twice over (TGGCN_GEMM_BN=128 disables them)
    static int bn_env = -1;
    if (bn_env < 0) {
        const char* e = getenv("TGGCN_GEMM_BN");
        bn_env = (e != nullptr && atoi(e) == 128) ? 128 : 256;
    }
    int tiles256 = 0;
    for (int i = 0; i < grp.count; ++i) tiles256 += cdiv(grp.p[i].M, TC_BM) * cdiv(grp.p[i].N, 256);
    const int bn = (bn_env == 256 && tiles256 >= 2 * num_sms()) ? 256 : 128;
    const int TC_BN = bn;
    int base_tiles = 0;
    for (int i = 0; i < grp.count; ++i) base_tiles += cdiv(grp.p[i].M, TC_BM) * cdiv(grp.p[i].N, TC_BN);
    for (int i = 0; i < grp.count; ++i) {
        GemmProblem& p = grp.p[i];
        // under-filled grid and a long reduction: split K over several CTAs per tile (weight gradients, M' x N' small, K' = rows)
        p.ksplit = 1;
        const int nkb = p.K / TC_BK;
        if (p.bias == nullptr && !p.relu && base_tiles * 2 <= num_sms() && nkb >= 16) {
            int ks = num_sms() / base_tiles;
            if (ks > nkb / 8) ks = nkb / 8;
            if (ks > 1) {
                p.ksplit = ks;
                if (!p.beta)
                    TG_CUDA_OK(cudaMemset2DAsync(p.C, sizeof(float) * (size_t)p.ldc, 0, sizeof(float) * (size_t)p.N, (size_t)p.M, stream));
            }
        }
        p.tile_begin = begin;
        begin += cdiv(p.M, TC_BM) * cdiv(p.N, TC_BN) * p.ksplit;
    }
    bool mask = false;
    for (int i = 0; i < grp.count; ++i) mask |= grp.p[i].amask != nullptr;
    auto launch = [&](auto kern, int smem) -> int {
        if (int rc = ensure_smem((const void*)kern, smem)) return rc;
        kern<<<begin, TC_THREADS, smem, stream>>>(grp);
        TG_LAUNCH_OK();
        return 0;
    };
    if (grp.precision == 1) {
        if (bn == 256) return mask ? launch(gemm_tc_kernel<true, 256, 1>, TcCfg<256>::SMEM_BYTES) : launch(gemm_tc_kernel<false, 256, 1>, TcCfg<256>::SMEM_BYTES);
        return mask ? launch(gemm_tc_kernel<true, 128, 1>, TcCfg<128>::SMEM_BYTES) : launch(gemm_tc_kernel<false, 128, 1>, TcCfg<128>::SMEM_BYTES);
    }
    if (bn == 256) return mask ? launch(gemm_tc_kernel<true, 256, 0>, TcCfg<256>::SMEM_BYTES) : launch(gemm_tc_kernel<false, 256, 0>, TcCfg<256>::SMEM_BYTES);
    return mask ? launch(gemm_tc_kernel<true, 128, 0>, TcCfg<128>::SMEM_BYTES) : launch(gemm_tc_kernel<false, 128, 0>, TcCfg<128>::SMEM_BYTES);
}

}  // namespace tg
