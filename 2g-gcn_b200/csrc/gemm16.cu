// Projection kernel, second design (K-B): every nn.Linear of the forward with M = B*T*entities rows
//     C[M,N] = act(A[M,K] * W[N,K]^T + bias[N])          (build_mlp, pyrutils/torch/models.py:31-33; vhoi/models.py:646-779)
// as a TMA-fed tcgen05 GEMM on 16-bit operand planes.
//
//   pack16x_kernel   fp32 operand (any row stride) -> dense 16-bit planes [plane][rows][K]:
//                    precision 0: x = hi + lo with hi = fp16(x), lo = fp16(x - hi) — 22 mantissa bits; weights are scaled by 2^8
//                    first so that their lo parts stay normal numbers (undone in the epilogue); precision 1: one bf16 plane.
//   gemm16_kernel    one 128 x BN output tile per CTA (BN = 256 or 128), 320 threads, warp-specialised:
//     warp 0      TMA producer: per 64-wide k-block ONE cp.async.bulk.tensor per operand (3-D maps K x rows x plane, box
//                 64 x 128|BN x planes, SWIZZLE_128B) lands both planes of the tile in the K-major layout the UMMA descriptors
//                 read; mbarrier expect_tx / complete_tx.  No register pass, no LDG latency in the pipeline: this is what
//                 bounded gemm_tc.cu's producers (ncu r01: long_scoreboard 5.4 per issue).
//     warp 1      MMA issuer: tcgen05.mma kind::f16, M = 128, N = BN, K = 16; fp32-class accuracy from the 3-term split
//                 a*w = lo*hi + hi*lo + hi*hi (small terms first) — half the tensor-pipe time of the 3xTF32 split — or a
//                 single bf16 product; fp32 accumulators in tensor memory; tcgen05.commit frees the stage.
//     warps 2-9   epilogue: tcgen05.ld (lane = output row), scale, bias, ReLU, float4 stores; optionally the 16-bit planes of
//                 the result as well, so that a following projection needs no pack pass.
// Several problems per launch (grouped), as gemm_tc.cu.  Accuracy class measured in tests/test_gpu_linear.py (path 4 / 5).
#include <stdlib.h>
#include <mutex>
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>

#include "common.cuh"
#include "gemm.h"
#include "tcgen05.cuh"

namespace tg {

namespace {

constexpr int G16_BM = 128;                  // output rows per tile (UMMA M)
constexpr int G16_BK = 64;                   // K granularity of the problems (K % 64 == 0)
constexpr int G16_EPI_WARPS = 8;
constexpr int G16_THREADS = (2 + G16_EPI_WARPS) * 32;
constexpr int G16_MAX_PROBLEMS = GEMM_MAX_PROBLEMS;
constexpr float G16_W_SCALE = 256.0f;        // fp16 split: weights are stored times 2^8

// BK: K elements per pipeline stage = one swizzle row of the operand tiles: 64 (SWIZZLE_128B) or 32 (SWIZZLE_64B).  The fp16-split
// 128 x 256 tile has 96 KB stages at BK = 64 — only two fit, and a two-stage ring exposes the TMA latency (measured 2900 cycles
// per 64 K against a 1536-cycle tensor-pipe floor); BK = 32 gives four 48 KB stages.
template <int PREC, int BN, int BK> struct G16Cfg {
    static constexpr int PLANES = PREC == 0 ? 2 : 1;
    static constexpr int ROW_BYTES = BK * 2;
    static constexpr int A_PLANE = G16_BM * ROW_BYTES;                 // bytes of one plane of the A tile
    static constexpr int B_PLANE = BN * ROW_BYTES;
    static constexpr int STAGE_BYTES = PLANES * (A_PLANE + B_PLANE);
    static constexpr int STAGES = (226 * 1024) / STAGE_BYTES > 8 ? 8 : (226 * 1024) / STAGE_BYTES;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024;
};

// K-major shared-memory matrix descriptor for rows of ROW_BYTES (128: SWIZZLE_128B, layout 2; 64: SWIZZLE_64B, layout 4); the
// stride between 8-row groups is 8 * ROW_BYTES
template <int ROW_BYTES> __device__ __forceinline__ uint64_t g16_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)((8u * ROW_BYTES) >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(ROW_BYTES == 128 ? 2 : 4) << 61;
    return d;
}

struct G16Problem {
    int M, N, K;            // K: reduction length of the operand planes (a multiple of 64; zero-padded by the pack kernels)
    int m_tiles, n_tiles, tile_begin;
    int relu, ldc;
    int beta;               // 1: add to C instead of overwriting it
    int ksplit;             // > 1: the K range is split over this many CTAs per tile, combined with atomics (C zeroed / beta)
    int atomic;             // 1: combine with atomics even without a K split (several problems of the launch add into this C)
    const float* bias;      // (N) or null
    const float* inv_a;     // device scalars: inverse of the dynamic (amax-derived, power-of-two) scale of an operand, or null
    const float* inv_b;
    float* C;               // fp32 result, row stride ldc
    void* out16;            // optional: 16-bit planes [plane][M][N] of the result (unscaled), or null
};

struct G16Launch {
    CUtensorMap amap[G16_MAX_PROBLEMS];
    CUtensorMap bmap[G16_MAX_PROBLEMS];
    G16Problem p[G16_MAX_PROBLEMS];
    int count;
    float acc_scale;
};

__device__ __forceinline__ void g16_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void g16_tma_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n" ::"r"(dst),
                 "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void g16_prefetch_map(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
// kind::f16 instruction descriptor: D = F32, A = B = F16 (0) or BF16 (1), both K-major, N >> 3 at bits 17-22, M >> 4 at 24-28
__device__ __forceinline__ uint32_t g16_idesc(int bf16, int M, int N) {
    return (1u << 4) | ((uint32_t)bf16 << 7) | ((uint32_t)bf16 << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void g16_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

template <int PREC, int BN, int BK>
__global__ void __launch_bounds__(G16_THREADS, 1) gemm16_kernel(const __grid_constant__ G16Launch L) {
    using Cfg = G16Cfg<PREC, BN, BK>;
    constexpr int STAGES = Cfg::STAGES, PLANES = Cfg::PLANES, RB = Cfg::ROW_BYTES;
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[2 * STAGES + 1];
    __shared__ uint32_t tmem_base_smem;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t tiles_u32 = smem_u32(tiles);
    const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[STAGES]), tfull = smem_u32(&bars[2 * STAGES]);

    int pi = 0;
#pragma unroll 1
    for (int i = 1; i < L.count; ++i)
        if ((int)blockIdx.x >= L.p[i].tile_begin) pi = i;
    const G16Problem& P = L.p[pi];
    int tile = blockIdx.x - P.tile_begin;
    const int ks = tile % P.ksplit;                                   // split-K slice of this CTA
    tile /= P.ksplit;
    const int mt = tile / P.n_tiles, nt = tile - mt * P.n_tiles;      // consecutive CTAs share the activation rows
    const int m0 = mt * G16_BM, n0 = nt * BN;
    const int nkb_all = P.K / BK;
    const int kb0 = (int)((long long)nkb_all * ks / P.ksplit);
    const int nkb = (int)((long long)nkb_all * (ks + 1) / P.ksplit) - kb0;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full0 + 8 * s, 1);               // the producer's expect_tx arrive
            mbar_init(empty0 + 8 * s, 1);              // one tcgen05.commit
        }
        mbar_init(tfull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 0 && lane == 0) {
        g16_prefetch_map(&L.amap[pi]);
        g16_prefetch_map(&L.bmap[pi]);
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_smem)), "r"((uint32_t)BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;

    if (warp == 0) {
        // ------------------------------ TMA producer ------------------------------
        if (lane == 0) {
#pragma unroll 1
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % STAGES;
                mbar_wait_backoff(empty0 + 8 * s, ((kb / STAGES) & 1) ^ 1);
                const uint32_t bar = full0 + 8 * s;
                g16_expect_tx(bar, (uint32_t)Cfg::STAGE_BYTES);
                const uint32_t st = tiles_u32 + s * Cfg::STAGE_BYTES;
                g16_tma_3d(st, &L.amap[pi], (kb0 + kb) * BK, m0, 0, bar);                               // box {BK k, 128 rows, planes}
                g16_tma_3d(st + PLANES * Cfg::A_PLANE, &L.bmap[pi], (kb0 + kb) * BK, n0, 0, bar);       // box {BK k, BN rows, planes}
            }
        }
    } else if (warp == 1) {
        // ------------------------------ MMA issuer ------------------------------
        const uint32_t idesc = g16_idesc(PREC, G16_BM, BN);
#pragma unroll 1
        for (int kb = 0; kb < nkb; ++kb) {
            const int s = kb % STAGES;
            mbar_wait(full0 + 8 * s, (kb / STAGES) & 1);
            tc_fence_after();
            if (lane == 0) {
                const uint32_t st = tiles_u32 + s * Cfg::STAGE_BYTES;
                const uint32_t a_hi = st, a_lo = st + Cfg::A_PLANE;
                const uint32_t b_hi = st + PLANES * Cfg::A_PLANE, b_lo = b_hi + Cfg::B_PLANE;
#pragma unroll
                for (int kk = 0; kk < BK / 16; ++kk) {
                    const uint32_t ko = kk * 32;               // 16 halves = 32 bytes along the swizzled row
                    if (PREC == 0) {
                        g16_mma(tmem_base, g16_desc<RB>(a_lo + ko), g16_desc<RB>(b_hi + ko), idesc, (kb | kk) != 0);
                        g16_mma(tmem_base, g16_desc<RB>(a_hi + ko), g16_desc<RB>(b_lo + ko), idesc, 1);
                        g16_mma(tmem_base, g16_desc<RB>(a_hi + ko), g16_desc<RB>(b_hi + ko), idesc, 1);
                    } else {
                        g16_mma(tmem_base, g16_desc<RB>(a_hi + ko), g16_desc<RB>(b_hi + ko), idesc, (kb | kk) != 0);
                    }
                }
                umma_commit(empty0 + 8 * s);                  // frees the stage when these MMAs have read it
                if (kb == nkb - 1) umma_commit(tfull);        // accumulator complete
            }
            __syncwarp();
        }
    } else {
        // ------------------------------ epilogue ------------------------------
        // TMEM lane quarter of a warp = warp id % 4; the two warps of a quarter take the two halves of the BN columns
        const int ew = warp - 2, q = warp & 3, half = ew >> 2;
        const int row = m0 + q * 32 + lane;
        const bool valid = row < P.M;
        const uint32_t tq = tmem_base + ((uint32_t)(q * 32) << 16);
        float sc = L.acc_scale;
        if (P.inv_a != nullptr) sc *= __ldg(P.inv_a);
        if (P.inv_b != nullptr) sc *= __ldg(P.inv_b);
        float* crow = P.C + (size_t)(valid ? row : 0) * P.ldc;
        const bool add_bias = P.bias != nullptr && ks == 0;
        mbar_wait_backoff(tfull, 0);
        tc_fence_after();
#pragma unroll 1
        for (int c = half * (BN / 32); c < (half + 1) * (BN / 32); ++c) {
            const int col = n0 + c * 16;
            if (col >= P.N) break;                            // (warp-uniform)
            float v[16];
            tmem_ld16(tq + (uint32_t)(c * 16), v);
            if (!valid) continue;
            if (add_bias) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float4 b = __ldg(reinterpret_cast<const float4*>(P.bias + col) + j);
                    v[4 * j] = fmaf(sc, v[4 * j], b.x); v[4 * j + 1] = fmaf(sc, v[4 * j + 1], b.y);
                    v[4 * j + 2] = fmaf(sc, v[4 * j + 2], b.z); v[4 * j + 3] = fmaf(sc, v[4 * j + 3], b.w);
                }
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] *= sc;
            }
            if (P.relu) {
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.0f);
            }
            if (P.ksplit > 1 || P.atomic) {                   // partial sums: C was zeroed (or holds the beta term)
#pragma unroll
                for (int j = 0; j < 4; ++j) atomicAdd(reinterpret_cast<float4*>(crow + col) + j, make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]));
                continue;
            }
            if (P.beta) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float4 o = reinterpret_cast<const float4*>(crow + col)[j];
                    v[4 * j] += o.x; v[4 * j + 1] += o.y; v[4 * j + 2] += o.z; v[4 * j + 3] += o.w;
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) reinterpret_cast<float4*>(crow + col)[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            if (P.out16 != nullptr) {
                const size_t off = (size_t)row * P.N + col;
                if (PREC == 0) {
                    __align__(16) __half hi[16];
                    __align__(16) __half lo[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        hi[j] = __float2half_rn(v[j]);
                        lo[j] = __float2half_rn(v[j] - __half2float(hi[j]));
                    }
                    __half* ph = reinterpret_cast<__half*>(P.out16) + off;
                    __half* pl = ph + (size_t)P.M * P.N;
                    reinterpret_cast<uint4*>(ph)[0] = reinterpret_cast<const uint4*>(hi)[0];
                    reinterpret_cast<uint4*>(ph)[1] = reinterpret_cast<const uint4*>(hi)[1];
                    reinterpret_cast<uint4*>(pl)[0] = reinterpret_cast<const uint4*>(lo)[0];
                    reinterpret_cast<uint4*>(pl)[1] = reinterpret_cast<const uint4*>(lo)[1];
                } else {
                    __align__(16) __nv_bfloat16 hi[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) hi[j] = __float2bfloat16_rn(v[j]);
                    __nv_bfloat16* ph = reinterpret_cast<__nv_bfloat16*>(P.out16) + off;
                    reinterpret_cast<uint4*>(ph)[0] = reinterpret_cast<const uint4*>(hi)[0];
                    reinterpret_cast<uint4*>(ph)[1] = reinterpret_cast<const uint4*>(hi)[1];
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"((uint32_t)BN) : "memory");
    }
}

// ---- operand preparation --------------------------------------------------------------------------------------------------------
// One job = one fp32 source matrix [rows][cols] (row stride ld) -> 16-bit planes.
//   transpose 0: planes[r][c] = f(src[r][c])                      out_ld >= cols                (activations, weights)
//   transpose 1: planes[c][m] = f(src[m + shift][c]) for m < rows, zero for rows <= m < out_ld   (weight-gradient operands:
//                the reduction index becomes contiguous; a row shift inside blocks of `period` rows pairs dG_t with h_{t-1})
// f: optional ReLU mask (value kept where mask[r][c] > 0), then a scale: fixed, or dynamic = the power of two that brings the
// operand's largest magnitude (amax word, filled by amax16_kernel) into [2^13, 2^14) — gradients are far below the fp16 range.
constexpr int P16_MAX_JOBS = 3 * G16_MAX_PROBLEMS;
struct P16Job {
    const float* src;
    int ld, rows, cols;
    const float* mask;
    int ldm;
    int transpose, shift, period;
    int out_ld;
    float scale;
    float* colsum;                  // transpose only: per-column sums of the masked, unscaled source (atomically accumulated), or null
    const unsigned int* amax;       // device word: bits of max |x| (dynamic scale), or null
    float* inv_scale_out;           // device scalar the dynamic inverse scale is published to, or null
    void* hi;                       // first element of the hi plane region this job writes
    size_t plane_elems;             // distance hi plane -> lo plane
};
struct P16Jobs {
    P16Job j[P16_MAX_JOBS];
    int count;
    unsigned int* err;          // bit 1 set when a value leaves the fp16 range (may be null)
};

__device__ __forceinline__ float p16_scale(const P16Job& J, bool publish) {
    if (J.amax == nullptr) return J.scale;
    const float amax = __uint_as_float(__ldg(J.amax));
    int e = 0;
    float sc = 1.0f, inv = 1.0f;
    if (amax > 0.0f && amax < INFINITY) {
        frexpf(amax, &e);                            // amax = m * 2^e, m in [0.5, 1)
        sc = ldexpf(1.0f, 14 - e);                   // amax * sc in [2^13, 2^14)
        inv = ldexpf(1.0f, e - 14);
    }
    if (publish && J.inv_scale_out != nullptr) *J.inv_scale_out = inv;
    return sc;
}

template <int PREC> __device__ __forceinline__ void p16_store8(const P16Job& J, size_t o, const float (&v)[8], bool& bad) {
    if (PREC == 0) {
        __align__(16) __half hi[8];
        __align__(16) __half lo[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            bad |= !(fabsf(v[k]) < 65504.0f);
            hi[k] = __float2half_rn(v[k]);
            lo[k] = __float2half_rn(v[k] - __half2float(hi[k]));
        }
        *reinterpret_cast<uint4*>(reinterpret_cast<__half*>(J.hi) + o) = *reinterpret_cast<const uint4*>(hi);
        *reinterpret_cast<uint4*>(reinterpret_cast<__half*>(J.hi) + J.plane_elems + o) = *reinterpret_cast<const uint4*>(lo);
    } else {
        __align__(16) __nv_bfloat16 hi[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) hi[k] = __float2bfloat16_rn(v[k]);
        *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(J.hi) + o) = *reinterpret_cast<const uint4*>(hi);
    }
}

// max |x| of every job with a dynamic scale (one word per job, zeroed by the host)
__global__ void __launch_bounds__(256) amax16_kernel(const P16Jobs jobs) {
    const P16Job& J = jobs.j[blockIdx.y];
    if (J.amax == nullptr) return;
    const int c4 = J.cols / 4;
    const size_t n4 = (size_t)J.rows * c4;
    float m = 0.0f;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const size_t r = i / c4;
        const int c = (int)(i - r * c4) * 4;
        const float4 x = __ldg(reinterpret_cast<const float4*>(J.src + r * J.ld + c));
        m = fmaxf(m, fmaxf(fmaxf(fabsf(x.x), fabsf(x.y)), fmaxf(fabsf(x.z), fabsf(x.w))));
    }
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0 && m > 0.0f) atomicMax(const_cast<unsigned int*>(J.amax), __float_as_uint(m));
}

template <int PREC> __global__ void __launch_bounds__(256) pack16x_kernel(const P16Jobs jobs) {
    const P16Job& J = jobs.j[blockIdx.y];
    bool bad = false;
    const float sc = p16_scale(J, blockIdx.x == 0 && threadIdx.x == 0);
    if (!J.transpose) {
        const int c8 = J.cols / 8;
        const size_t n8 = (size_t)J.rows * c8;
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
            const size_t r = i / c8;
            const int c = (int)(i - r * c8) * 8;
            const float4 x0 = __ldg(reinterpret_cast<const float4*>(J.src + r * J.ld + c));
            const float4 x1 = __ldg(reinterpret_cast<const float4*>(J.src + r * J.ld + c) + 1);
            float v[8] = {x0.x * sc, x0.y * sc, x0.z * sc, x0.w * sc, x1.x * sc, x1.y * sc, x1.z * sc, x1.w * sc};
            if (J.mask != nullptr) {
                const float4 m0 = __ldg(reinterpret_cast<const float4*>(J.mask + r * J.ldm + c));
                const float4 m1 = __ldg(reinterpret_cast<const float4*>(J.mask + r * J.ldm + c) + 1);
                const float mk[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
#pragma unroll
                for (int k = 0; k < 8; ++k) v[k] = mk[k] > 0.0f ? v[k] : 0.0f;
            }
            p16_store8<PREC>(J, r * J.out_ld + c, v, bad);
        }
    } else {
        // 64 source rows x 32 source columns per tile through shared memory; output rows = source columns
        __shared__ float t[32][65];
        __shared__ float csum[8][32];
        const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
        const int ctiles = (J.cols + 31) / 32, mtiles = J.out_ld / 64;
        for (int tile = blockIdx.x; tile < ctiles * mtiles; tile += gridDim.x) {
            const int ct = tile / mtiles, mtile = tile - ct * mtiles;
            const int c0 = ct * 32, m0 = mtile * 64;
            float part = 0.0f;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int ml = i * 8 + ty, m = m0 + ml, c = c0 + tx;
                float x = 0.0f;
                if (m < J.rows && c < J.cols) {
                    bool ok = true;
                    if (J.shift != 0) {
                        const int pos = m % J.period + J.shift;
                        ok = pos >= 0 && pos < J.period;
                    }
                    if (ok) {
                        x = __ldg(J.src + (size_t)(m + J.shift) * J.ld + c);
                        if (J.mask != nullptr && !(__ldg(J.mask + (size_t)m * J.ldm + c) > 0.0f)) x = 0.0f;
                    }
                }
                t[tx][ml] = x * sc;
                part += x;
            }
            if (J.colsum != nullptr) csum[ty][tx] = part;
            __syncthreads();
            if (J.colsum != nullptr && ty == 0 && c0 + tx < J.cols) {
                float v = 0.0f;
#pragma unroll
                for (int w = 0; w < 8; ++w) v += csum[w][tx];
                atomicAdd(J.colsum + c0 + tx, v);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int cl = ty + 8 * j, c = c0 + cl;
                if (c < J.cols) {
                    const float x0 = t[cl][2 * tx], x1 = t[cl][2 * tx + 1];
                    const size_t o = (size_t)c * J.out_ld + m0 + 2 * tx;
                    if (PREC == 0) {
                        bad |= !(fabsf(x0) < 65504.0f) || !(fabsf(x1) < 65504.0f);
                        const __half2 h = __floats2half2_rn(x0, x1);
                        const float2 f = __half22float2(h);
                        const __half2 l = __floats2half2_rn(x0 - f.x, x1 - f.y);
                        *reinterpret_cast<__half2*>(reinterpret_cast<__half*>(J.hi) + o) = h;
                        *reinterpret_cast<__half2*>(reinterpret_cast<__half*>(J.hi) + J.plane_elems + o) = l;
                    } else {
                        *reinterpret_cast<__nv_bfloat162*>(reinterpret_cast<__nv_bfloat16*>(J.hi) + o) = __floats2bfloat162_rn(x0, x1);
                    }
                }
            }
            __syncthreads();
        }
    }
    if (PREC == 0 && bad && jobs.err != nullptr) atomicOr(jobs.err, 2u);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn g16_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// 3-D map over `planes` dense row-major [rows][K] matrices of 16-bit elements: box = bk K-elements x box_rows rows x planes,
// 128- or 64-byte swizzle (the K-major layouts of the UMMA descriptors), out-of-range rows read as zeros.
int g16_make_map(CUtensorMap* m, const void* base, int precision, size_t K, size_t rows, int box_rows, int bk) {
    EncodeTiledFn enc = g16_encode_fn();
    TG_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from this driver");
    const int planes = precision ? 1 : 2;
    cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)rows, (cuuint64_t)planes};
    cuuint64_t strides[2] = {(cuuint64_t)K * 2, (cuuint64_t)K * 2 * rows};
    cuuint32_t box[3] = {(cuuint32_t)bk, (cuuint32_t)box_rows, (cuuint32_t)planes};
    cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = enc(m, precision ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(base), dims,
                           strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, bk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    TG_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d): K=%zu rows=%zu box rows %d", (int)r, K, rows, box_rows);
    return 0;
}

template <int PREC, int BN, int BK> int g16_launch_t(const G16Launch& L, int tiles, cudaStream_t stream) {
    using Cfg = G16Cfg<PREC, BN, BK>;
    if (int rc = ensure_smem((const void*)gemm16_kernel<PREC, BN, BK>, Cfg::SMEM_BYTES)) return rc;
    gemm16_kernel<PREC, BN, BK><<<tiles, G16_THREADS, Cfg::SMEM_BYTES, stream>>>(L);
    TG_LAUNCH_OK();
    return 0;
}

size_t up256(size_t x) { return (x + 255) / 256 * 256; }
int pad64(int x) { return (x + 63) / 64 * 64; }

// rows / reduction length of the operand planes a source produces
int op_rows(const Gemm16Operand& o) { return o.transpose ? o.cols : o.rows; }
int op_k(const Gemm16Operand& o) { return o.transpose ? pad64(o.rows) : o.cols; }
bool same_op(const Gemm16Operand& a, const Gemm16Operand& b) {
    return a.src == b.src && a.ld == b.ld && a.rows == b.rows && a.cols == b.cols && a.mask == b.mask && a.ldm == b.ldm &&
           a.transpose == b.transpose && a.shift == b.shift && a.period == b.period && a.dynamic == b.dynamic && a.scale == b.scale;
}
bool op_ok(const Gemm16Operand& o) {
    if (o.src == nullptr || o.rows <= 0 || o.cols <= 0) return false;
    if (o.transpose) return true;                                   // scalar loads; any shape (the reduction length is padded)
    if (o.cols % G16_BK != 0 || o.ld % 4 != 0 || (reinterpret_cast<uintptr_t>(o.src) & 15)) return false;
    if (o.mask != nullptr && (o.ldm % 4 != 0 || (reinterpret_cast<uintptr_t>(o.mask) & 15))) return false;
    return true;
}
// dynamic-scale operands are reduced with float4 loads
bool amax_ok(const Gemm16Operand& o) { return !o.dynamic || (o.cols % 4 == 0 && o.ld % 4 == 0 && (reinterpret_cast<uintptr_t>(o.src) & 15) == 0); }

}  // namespace

bool gemm16_enabled() {
    static int enabled = -1;
    if (enabled < 0) {
        const char* e = getenv("TGGCN_GEMM16");
        enabled = (e != nullptr && e[0] == '0') ? 0 : 1;
    }
    return enabled != 0;
}

bool gemm16_eligible(const Gemm16Problem* p, int count) {
    if (!gemm16_enabled() || count <= 0 || count > G16_MAX_PROBLEMS) return false;
    for (int i = 0; i < count; ++i) {
        const Gemm16Problem& q = p[i];
        if (!op_ok(q.a) || !op_ok(q.b) || !amax_ok(q.a) || !amax_ok(q.b)) return false;
        if (op_k(q.a) != op_k(q.b)) return false;
        if (op_rows(q.b) % 16 != 0 || q.ldc % 4 != 0 || (reinterpret_cast<uintptr_t>(q.C) & 15)) return false;
        if (q.bias != nullptr && (reinterpret_cast<uintptr_t>(q.bias) & 15)) return false;
    }
    return true;
}

size_t gemm16_scratch_bytes(const Gemm16Problem* p, int count) {
    size_t total = 256;                                                 // amax words + published inverse scales
    for (int i = 0; i < count; ++i) {
        bool dupa = false, dupb = false;
        for (int j = 0; j < i; ++j) {
            dupa |= same_op(p[j].a, p[i].a);
            dupb |= same_op(p[j].b, p[i].b);
        }
        if (!dupa) total += up256((size_t)op_rows(p[i].a) * op_k(p[i].a) * 4);
        if (!dupb) total += up256((size_t)op_rows(p[i].b) * op_k(p[i].b) * 4);
    }
    return total;
}

// C[M,N] (+)= act(A' B'^T + bias) for every problem, A' / B' = the 16-bit planes of the two operands (packed here).
// precision 0: fp16 (hi, lo) split, fp32-class accuracy; 1: bf16 operands.  scratch: gemm16_scratch_bytes() bytes, 256-byte aligned.
// err: status word (bit 1 = an operand left the fp16 range), may be null.
int launch_gemm16(const Gemm16Problem* prob, int count, int precision, void* scratch, size_t scratch_bytes, unsigned int* err,
                  cudaStream_t stream) {
    if (count == 0) return 0;
    TG_REQUIRE(gemm16_eligible(prob, count), "gemm16: the group does not qualify (K %% 64, N %% 16, alignment)");
    TG_REQUIRE(scratch != nullptr && scratch_bytes >= gemm16_scratch_bytes(prob, count), "gemm16: operand scratch too small");
    TG_REQUIRE((reinterpret_cast<uintptr_t>(scratch) & 255) == 0, "gemm16: scratch must be 256-byte aligned");
    // tile width: 256 columns move a quarter fewer operand bytes per FLOP; choose by waves x bytes per k-block
    auto tiles_of = [&](int bn) { int t = 0; for (int i = 0; i < count; ++i) t += cdiv(op_rows(prob[i].a), G16_BM) * cdiv(op_rows(prob[i].b), bn); return t; };
    static int bn_env = -1, bk_env = -1;
    if (bn_env < 0) {
        const char* e = getenv("TGGCN_GEMM16_BN");
        bn_env = e != nullptr ? atoi(e) : 0;
        e = getenv("TGGCN_GEMM16_BK");
        bk_env = e != nullptr ? atoi(e) : 0;
    }
    // Tile width and launch order by a makespan estimate.  A tile costs ~ (K / 64) x (128 + BN) (the main loop runs at the rate the
    // operand bytes arrive) plus a fixed prologue / epilogue share; tiles are issued longest first (the hardware hands blocks to
    // SMs in index order as they free up = greedy longest-processing-time scheduling), and the width with the shorter estimated
    // makespan over the SMs wins.  ncu (profiles/r02_ncu_gemm16_embed.txt): the embedding stage ran 160 tiles of 128 x 256 with
    // its 64 long tiles (K = 3328) LAST — 1.08 waves, tensor pipe 37 % against 61 % for the well-filled gs stage.
    int order[G16_MAX_PROBLEMS];
    for (int i = 0; i < count; ++i) order[i] = i;
    for (int i = 1; i < count; ++i)                                     // problems by reduction length, longest first
        for (int j = i; j > 0 && op_k(prob[order[j]].a) > op_k(prob[order[j - 1]].a); --j) { const int t = order[j]; order[j] = order[j - 1]; order[j - 1] = t; }
    auto makespan = [&](int bnx) -> double {
        const int S = num_sms();
        if (tiles_of(bnx) > 16 * S) return (double)tiles_of(bnx) / S * (128 + bnx);      // many waves: quantisation does not matter
        double load[256];
        const int SS = S < 256 ? S : 256;
        for (int i = 0; i < SS; ++i) load[i] = 0.0;
        for (int oi = 0; oi < count; ++oi) {
            const Gemm16Problem& q = prob[order[oi]];
            const double cost = (double)(op_k(q.a) / 64) * (128 + bnx) + 6.0 * 384;          // + ~6 k-blocks of prologue / epilogue
            const int nt = cdiv(op_rows(q.a), G16_BM) * cdiv(op_rows(q.b), bnx);
            for (int t = 0; t < nt; ++t) {
                int best = 0;
                for (int i = 1; i < SS; ++i)
                    if (load[i] < load[best]) best = i;
                load[best] += cost;
            }
        }
        double m = 0.0;
        for (int i = 0; i < SS; ++i) m = load[i] > m ? load[i] : m;
        return m;
    };
    int bn = 0;
    {   // the estimate costs O(tiles x SMs) host operations: remember it per problem signature (a forward repeats its six stages)
        static std::mutex mu;
        static unsigned long long keys[128];
        static int vals[128];
        unsigned long long h = 1469598103934665603ull ^ (unsigned long long)num_sms();
        for (int i = 0; i < count; ++i) {
            const unsigned long long v[3] = {(unsigned long long)op_rows(prob[order[i]].a), (unsigned long long)op_rows(prob[order[i]].b),
                                             (unsigned long long)op_k(prob[order[i]].a)};
            for (unsigned long long x : v) h = (h ^ x) * 1099511628211ull;
        }
        if (h == 0) h = 1;
        std::lock_guard<std::mutex> lock(mu);
        const int slot = (int)(h % 128);
        if (keys[slot] == h) bn = vals[slot];
        else {
            bn = makespan(128) < makespan(256) ? 128 : 256;
            keys[slot] = h; vals[slot] = bn;
        }
    }
    if (bn_env == 128 || bn_env == 256) bn = bn_env;
    // stage depth: K = 32 per stage (SWIZZLE_64B rows) for the fp16-split 256-wide tile, whose 64-deep stages are 96 KB
    int bk = (precision == 0 && bn == 256) ? 32 : 64;
    if (bk_env == 32 || bk_env == 64) bk = bk_env;

    G16Launch L;
    memset(&L, 0, sizeof(L));
    P16Jobs jobs;
    memset(&jobs, 0, sizeof(jobs));
    jobs.err = err;
    uint8_t* ws = reinterpret_cast<uint8_t*>(scratch);
    unsigned int* amax_words = reinterpret_cast<unsigned int*>(ws);          // [0, 32): amax bits; [32, 64): inverse scales (float)
    float* inv_words = reinterpret_cast<float*>(ws) + 32;
    size_t off = 256;
    const void* planes[2][G16_MAX_PROBLEMS];
    const float* invs[2][G16_MAX_PROBLEMS];
    int jobidx[2][G16_MAX_PROBLEMS];
    int n_dyn = 0, begin = 0, base_tiles = tiles_of(bn);
    size_t max_work = 0;
    Gemm16Problem sorted[G16_MAX_PROBLEMS];
    for (int i = 0; i < count; ++i) sorted[i] = prob[order[i]];
    prob = sorted;                                                      // longest reduction first (see above)
    for (int i = 0; i < count; ++i) {
        const Gemm16Problem& p = prob[i];
        for (int side = 0; side < 2; ++side) {
            const Gemm16Operand& o = side == 0 ? p.a : p.b;
            planes[side][i] = nullptr;
            invs[side][i] = nullptr;
            jobidx[side][i] = -1;
            for (int j = 0; j < i; ++j) {
                const Gemm16Operand& oj = side == 0 ? prob[j].a : prob[j].b;
                if (same_op(oj, o)) { planes[side][i] = planes[side][j]; invs[side][i] = invs[side][j]; jobidx[side][i] = jobidx[side][j]; }
            }
            if (o.colsum != nullptr) {
                TG_REQUIRE(o.transpose, "gemm16: column sums are fused into the transposed pack only");
                if (!o.colsum_beta) TG_CUDA_OK(cudaMemsetAsync(o.colsum, 0, sizeof(float) * (size_t)o.cols, stream));
            }
            if (planes[side][i] != nullptr) {                      // packed once; the job may still owe this problem's column sums
                if (o.colsum != nullptr) {
                    P16Job& Jd = jobs.j[jobidx[side][i]];
                    TG_REQUIRE(Jd.colsum == nullptr || Jd.colsum == o.colsum, "gemm16: two column-sum destinations for one operand");
                    Jd.colsum = o.colsum;
                }
                continue;
            }
            const int R = op_rows(o), K = op_k(o);
            planes[side][i] = ws + off;
            off += up256((size_t)R * K * 4);
            TG_REQUIRE(jobs.count < P16_MAX_JOBS, "gemm16: too many operands in one group");
            jobidx[side][i] = jobs.count;
            P16Job& J = jobs.j[jobs.count++];
            J.colsum = o.colsum;
            J.src = o.src; J.ld = o.ld; J.rows = o.rows; J.cols = o.cols; J.mask = o.mask; J.ldm = o.ldm;
            J.transpose = o.transpose; J.shift = o.shift; J.period = o.period > 0 ? o.period : 1;
            J.out_ld = K; J.scale = o.scale != 0.0f ? o.scale : 1.0f;
            J.hi = const_cast<void*>(planes[side][i]); J.plane_elems = (size_t)R * K;
            if (o.dynamic && precision == 0) {
                TG_REQUIRE(n_dyn < 32, "gemm16: too many dynamically scaled operands");
                J.amax = amax_words + n_dyn; J.inv_scale_out = inv_words + n_dyn;
                invs[side][i] = inv_words + n_dyn;
                ++n_dyn;
            }
            const size_t work = (size_t)(o.transpose ? pad64(o.rows) : o.rows) * o.cols;
            if (work > max_work) max_work = work;
        }
        const int M = op_rows(p.a), N = op_rows(p.b), K = op_k(p.a);
        if (int rc = g16_make_map(&L.amap[i], planes[0][i], precision, K, M, G16_BM, bk)) return rc;
        if (int rc = g16_make_map(&L.bmap[i], planes[1][i], precision, K, N, bn, bk)) return rc;
        G16Problem& q = L.p[i];
        q.M = M; q.N = N; q.K = K; q.relu = p.relu; q.ldc = p.ldc; q.bias = p.bias; q.C = p.C; q.out16 = nullptr; q.beta = p.beta;
        q.inv_a = invs[0][i]; q.inv_b = invs[1][i];
        q.m_tiles = cdiv(M, G16_BM); q.n_tiles = cdiv(N, bn); q.tile_begin = begin;
        // under-filled grid and a long reduction (weight gradients): split K over several CTAs per tile, combined with atomics
        q.ksplit = 1;
        q.atomic = 0;
        for (int j = 0; j < count; ++j)
            if (j != i && prob[j].C == p.C) q.atomic = 1;              // e.g. the two time directions of a shared message MLP
        const int nkb = K / 64;
        if (!p.relu && base_tiles * 2 <= num_sms() && nkb >= 8) {
            int ks = num_sms() / base_tiles;
            if (ks > nkb / 4) ks = nkb / 4;
            if (ks > 1) q.ksplit = ks;
        }
        if ((q.ksplit > 1 || q.atomic) && !p.beta)
            TG_CUDA_OK(cudaMemset2DAsync(p.C, sizeof(float) * (size_t)p.ldc, 0, sizeof(float) * (size_t)N, (size_t)M, stream));
        begin += q.m_tiles * q.n_tiles * q.ksplit;
    }
    L.count = count;
    L.acc_scale = 1.0f;
    for (int i = 0; i < count; ++i) {                                  // fixed scales are folded per problem into the bias-free factor
        // (all problems of a group share acc_scale: fixed scales must agree across the group)
        const float f = (prob[i].a.dynamic || precision ? 1.0f : (prob[i].a.scale != 0.0f ? prob[i].a.scale : 1.0f)) *
                        (prob[i].b.dynamic || precision ? 1.0f : (prob[i].b.scale != 0.0f ? prob[i].b.scale : 1.0f));
        if (i == 0) L.acc_scale = 1.0f / f;
        else TG_REQUIRE(L.acc_scale == 1.0f / f, "gemm16: the fixed operand scales of a group must agree");
    }
    {
        int gx = (int)((max_work / 8 + 255) / 256);
        const int cap = 4 * num_sms();
        if (gx > cap) gx = cap;
        if (gx < 1) gx = 1;
        dim3 grid(gx, jobs.count);
        if (n_dyn > 0) {
            TG_CUDA_OK(cudaMemsetAsync(amax_words, 0, 32 * sizeof(unsigned int), stream));
            amax16_kernel<<<grid, 256, 0, stream>>>(jobs);
            TG_LAUNCH_OK();
        }
        if (precision) {
            for (int i = 0; i < jobs.count; ++i) jobs.j[i].scale = 1.0f;          // bf16 has the fp32 exponent range: no scaling
            pack16x_kernel<1><<<grid, 256, 0, stream>>>(jobs);
        } else {
            pack16x_kernel<0><<<grid, 256, 0, stream>>>(jobs);
        }
        TG_LAUNCH_OK();
    }
    if (precision) {
        if (bk == 32) return bn == 256 ? g16_launch_t<1, 256, 32>(L, begin, stream) : g16_launch_t<1, 128, 32>(L, begin, stream);
        return bn == 256 ? g16_launch_t<1, 256, 64>(L, begin, stream) : g16_launch_t<1, 128, 64>(L, begin, stream);
    }
    if (bk == 32) return bn == 256 ? g16_launch_t<0, 256, 32>(L, begin, stream) : g16_launch_t<0, 128, 32>(L, begin, stream);
    return bn == 256 ? g16_launch_t<0, 256, 64>(L, begin, stream) : g16_launch_t<0, 128, 64>(L, begin, stream);
}

// The forward's grouped nn.Linear problems (gemm.h) on this kernel: activations unscaled, weights times 2^8.
static void g16_from_group(const GemmGroup& grp, Gemm16Problem* out) {
    for (int i = 0; i < grp.count; ++i) {
        const GemmProblem& p = grp.p[i];
        Gemm16Problem& q = out[i];
        memset(&q, 0, sizeof(q));
        q.a.src = p.A; q.a.ld = p.lda; q.a.rows = p.M; q.a.cols = p.K; q.a.mask = p.amask; q.a.ldm = p.ldm; q.a.scale = 1.0f;
        q.b.src = p.W; q.b.ld = p.ldw; q.b.rows = p.N; q.b.cols = p.K; q.b.scale = G16_W_SCALE;
        q.bias = p.bias; q.C = p.C; q.ldc = p.ldc; q.relu = p.relu; q.beta = p.beta;
    }
}
bool gemm16_eligible(const GemmGroup& grp) {
    if (grp.count <= 0 || grp.count > G16_MAX_PROBLEMS) return false;
    Gemm16Problem q[G16_MAX_PROBLEMS];
    g16_from_group(grp, q);
    return gemm16_eligible(q, grp.count);
}
size_t gemm16_scratch_bytes(const GemmGroup& grp) {
    Gemm16Problem q[G16_MAX_PROBLEMS];
    g16_from_group(grp, q);
    return gemm16_scratch_bytes(q, grp.count);
}
int launch_gemm16(GemmGroup& grp, int precision, void* scratch, size_t scratch_bytes, unsigned int* err, cudaStream_t stream) {
    if (grp.count == 0) return 0;
    TG_REQUIRE(grp.count <= G16_MAX_PROBLEMS, "gemm16: too many problems in one group (%d)", grp.count);
    Gemm16Problem q[G16_MAX_PROBLEMS];
    g16_from_group(grp, q);
    return launch_gemm16(q, grp.count, precision, scratch, scratch_bytes, err, stream);
}

}  // namespace tg
