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
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>

#include "common.cuh"
#include "gemm.h"
#include "tcgen05.cuh"

namespace tg {

namespace {

constexpr int G16_BM = 128;                  // output rows per tile (UMMA M)
constexpr int G16_BK = 64;                   // K elements per k-block: 128 bytes of 16-bit = one swizzle row
constexpr int G16_EPI_WARPS = 8;
constexpr int G16_THREADS = (2 + G16_EPI_WARPS) * 32;
constexpr int G16_MAX_PROBLEMS = GEMM_MAX_PROBLEMS;
constexpr float G16_W_SCALE = 256.0f;        // fp16 split: weights are stored times 2^8

template <int PREC, int BN> struct G16Cfg {
    static constexpr int PLANES = PREC == 0 ? 2 : 1;
    static constexpr int A_PLANE = G16_BM * 128;                       // bytes of one plane of the A tile
    static constexpr int B_PLANE = BN * 128;
    static constexpr int STAGE_BYTES = PLANES * (A_PLANE + B_PLANE);   // fp16 split: 96 KB (BN 256) / 64 KB (BN 128); bf16: half
    static constexpr int STAGES = (226 * 1024) / STAGE_BYTES > 6 ? 6 : (226 * 1024) / STAGE_BYTES;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024;
};

struct G16Problem {
    int M, N, K;
    int m_tiles, n_tiles, tile_begin;
    int relu, ldc;
    const float* bias;      // (N) or null
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

template <int PREC, int BN>
__global__ void __launch_bounds__(G16_THREADS, 1) gemm16_kernel(const __grid_constant__ G16Launch L) {
    using Cfg = G16Cfg<PREC, BN>;
    constexpr int STAGES = Cfg::STAGES, PLANES = Cfg::PLANES;
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
    const int tile = blockIdx.x - P.tile_begin;
    const int mt = tile / P.n_tiles, nt = tile - mt * P.n_tiles;      // consecutive CTAs share the activation rows
    const int m0 = mt * G16_BM, n0 = nt * BN;
    const int nkb = P.K / G16_BK;

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
                g16_tma_3d(st, &L.amap[pi], kb * G16_BK, m0, 0, bar);                               // box {64 k, 128 rows, planes}
                g16_tma_3d(st + PLANES * Cfg::A_PLANE, &L.bmap[pi], kb * G16_BK, n0, 0, bar);       // box {64 k, BN rows, planes}
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
                for (int kk = 0; kk < G16_BK / 16; ++kk) {
                    const uint32_t ko = kk * 32;               // 16 halves = 32 bytes along the swizzled row
                    if (PREC == 0) {
                        g16_mma(tmem_base, umma_desc(a_lo + ko), umma_desc(b_hi + ko), idesc, (kb | kk) != 0);
                        g16_mma(tmem_base, umma_desc(a_hi + ko), umma_desc(b_lo + ko), idesc, 1);
                        g16_mma(tmem_base, umma_desc(a_hi + ko), umma_desc(b_hi + ko), idesc, 1);
                    } else {
                        g16_mma(tmem_base, umma_desc(a_hi + ko), umma_desc(b_hi + ko), idesc, (kb | kk) != 0);
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
        const float sc = L.acc_scale;
        float* crow = P.C + (size_t)(valid ? row : 0) * P.ldc;
        mbar_wait_backoff(tfull, 0);
        tc_fence_after();
#pragma unroll 1
        for (int c = half * (BN / 32); c < (half + 1) * (BN / 32); ++c) {
            const int col = n0 + c * 16;
            if (col >= P.N) break;                            // (warp-uniform)
            float v[16];
            tmem_ld16(tq + (uint32_t)(c * 16), v);
            if (!valid) continue;
            if (P.bias != nullptr) {
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
constexpr int P16_MAX_JOBS = 2 * G16_MAX_PROBLEMS;
struct P16Job {
    const float* src;
    int ld, rows, cols;
    float scale;
    void* hi;                   // [rows][cols]; the lo plane follows at + rows*cols elements
};
struct P16Jobs {
    P16Job j[P16_MAX_JOBS];
    int count;
    unsigned int* err;          // bit 1 set when a value leaves the fp16 range (may be null)
};

template <int PREC> __global__ void __launch_bounds__(256) pack16x_kernel(const P16Jobs jobs) {
    const P16Job& J = jobs.j[blockIdx.y];
    const int c8 = J.cols / 8;
    const size_t n8 = (size_t)J.rows * c8, plane = (size_t)J.rows * J.cols;
    const float sc = J.scale;
    bool bad = false;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
        const size_t r = i / c8;
        const int c = (int)(i - r * c8) * 8;
        const float4 x0 = __ldg(reinterpret_cast<const float4*>(J.src + r * J.ld + c));
        const float4 x1 = __ldg(reinterpret_cast<const float4*>(J.src + r * J.ld + c) + 1);
        const float v[8] = {x0.x * sc, x0.y * sc, x0.z * sc, x0.w * sc, x1.x * sc, x1.y * sc, x1.z * sc, x1.w * sc};
        const size_t o = r * J.cols + c;
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
            *reinterpret_cast<uint4*>(reinterpret_cast<__half*>(J.hi) + plane + o) = *reinterpret_cast<const uint4*>(lo);
        } else {
            __align__(16) __nv_bfloat16 hi[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) hi[k] = __float2bfloat16_rn(v[k]);
            *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(J.hi) + o) = *reinterpret_cast<const uint4*>(hi);
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

// 3-D map over `planes` dense row-major [rows][K] matrices of 16-bit elements: box = 64 K-elements x box_rows rows x planes,
// 128-byte swizzle (the K-major layout of the UMMA descriptors), out-of-range rows read as zeros.
int g16_make_map(CUtensorMap* m, const void* base, int precision, size_t K, size_t rows, int box_rows) {
    EncodeTiledFn enc = g16_encode_fn();
    TG_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from this driver");
    const int planes = precision ? 1 : 2;
    cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)rows, (cuuint64_t)planes};
    cuuint64_t strides[2] = {(cuuint64_t)K * 2, (cuuint64_t)K * 2 * rows};
    cuuint32_t box[3] = {(cuuint32_t)G16_BK, (cuuint32_t)box_rows, (cuuint32_t)planes};
    cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = enc(m, precision ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(base), dims,
                           strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    TG_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d): K=%zu rows=%zu box rows %d", (int)r, K, rows, box_rows);
    return 0;
}

template <int PREC, int BN> int g16_launch_t(const G16Launch& L, int tiles, cudaStream_t stream) {
    using Cfg = G16Cfg<PREC, BN>;
    if (int rc = ensure_smem((const void*)gemm16_kernel<PREC, BN>, Cfg::SMEM_BYTES)) return rc;
    gemm16_kernel<PREC, BN><<<tiles, G16_THREADS, Cfg::SMEM_BYTES, stream>>>(L);
    TG_LAUNCH_OK();
    return 0;
}

size_t up256(size_t x) { return (x + 255) / 256 * 256; }

}  // namespace

bool gemm16_eligible(const GemmGroup& grp) {
    static int enabled = -1;
    if (enabled < 0) {
        const char* e = getenv("TGGCN_GEMM16");
        enabled = (e != nullptr && e[0] == '0') ? 0 : 1;
    }
    if (!enabled || grp.count == 0 || grp.count > G16_MAX_PROBLEMS) return false;
    for (int i = 0; i < grp.count; ++i) {
        const GemmProblem& p = grp.p[i];
        if (p.K % G16_BK != 0 || p.N % 16 != 0 || p.ldc % 4 != 0 || p.lda % 4 != 0 || p.ldw % 4 != 0) return false;
        if (p.amask != nullptr || p.beta != 0) return false;
        if ((reinterpret_cast<uintptr_t>(p.A) | reinterpret_cast<uintptr_t>(p.W) | reinterpret_cast<uintptr_t>(p.C)) & 15) return false;
        if (p.bias != nullptr && (reinterpret_cast<uintptr_t>(p.bias) & 15)) return false;
    }
    return true;
}

// bytes of operand-plane scratch launch_gemm16 needs for this group (distinct operands are packed once)
size_t gemm16_scratch_bytes(const GemmGroup& grp) {
    size_t total = 0;
    for (int i = 0; i < grp.count; ++i) {
        const GemmProblem& p = grp.p[i];
        bool dupa = false, dupw = false;
        for (int j = 0; j < i; ++j) {
            const GemmProblem& q = grp.p[j];
            dupa |= q.A == p.A && q.lda == p.lda && q.M == p.M && q.K == p.K;
            dupw |= q.W == p.W && q.ldw == p.ldw && q.N == p.N && q.K == p.K;
        }
        if (!dupa) total += up256((size_t)p.M * p.K * 4);
        if (!dupw) total += up256((size_t)p.N * p.K * 4);
    }
    return total;
}

// precision 0: fp16 (hi, lo) split, fp32-class accuracy; 1: bf16 operands.  scratch: gemm16_scratch_bytes(grp) bytes, 256-byte aligned.
// err: status word of the forward (bit 1 = an operand left the fp16 range), may be null.
int launch_gemm16(GemmGroup& grp, int precision, void* scratch, size_t scratch_bytes, unsigned int* err, cudaStream_t stream) {
    if (grp.count == 0) return 0;
    TG_REQUIRE(gemm16_eligible(grp), "gemm16: the group does not qualify (K %% 64, N %% 16, alignment, no mask / beta)");
    TG_REQUIRE(scratch != nullptr && scratch_bytes >= gemm16_scratch_bytes(grp), "gemm16: operand scratch too small");
    TG_REQUIRE((reinterpret_cast<uintptr_t>(scratch) & 255) == 0, "gemm16: scratch must be 256-byte aligned");
    // tile width: 256 columns move a quarter fewer operand bytes per FLOP; choose by waves x bytes per k-block
    auto tiles_of = [&](int bn) { int t = 0; for (int i = 0; i < grp.count; ++i) t += cdiv(grp.p[i].M, G16_BM) * cdiv(grp.p[i].N, bn); return t; };
    static int bn_env = -1;
    if (bn_env < 0) {
        const char* e = getenv("TGGCN_GEMM16_BN");
        bn_env = e != nullptr ? atoi(e) : 0;
    }
    int bn = 256;
    {
        const long long c256 = (long long)cdiv(tiles_of(256), num_sms()) * (128 + 256), c128 = (long long)cdiv(tiles_of(128), num_sms()) * (128 + 128);
        if (c128 < c256) bn = 128;
        if (bn_env == 128 || bn_env == 256) bn = bn_env;
    }
    G16Launch L;
    memset(&L, 0, sizeof(L));
    P16Jobs jobs;
    memset(&jobs, 0, sizeof(jobs));
    jobs.err = err;
    uint8_t* ws = reinterpret_cast<uint8_t*>(scratch);
    size_t off = 0;
    const void* aplane[G16_MAX_PROBLEMS];
    const void* wplane[G16_MAX_PROBLEMS];
    int begin = 0;
    size_t max_elems = 0;
    for (int i = 0; i < grp.count; ++i) {
        const GemmProblem& p = grp.p[i];
        aplane[i] = wplane[i] = nullptr;
        for (int j = 0; j < i; ++j) {
            const GemmProblem& q = grp.p[j];
            if (q.A == p.A && q.lda == p.lda && q.M == p.M && q.K == p.K) aplane[i] = aplane[j];
            if (q.W == p.W && q.ldw == p.ldw && q.N == p.N && q.K == p.K) wplane[i] = wplane[j];
        }
        if (aplane[i] == nullptr) {
            aplane[i] = ws + off;
            off += up256((size_t)p.M * p.K * 4);
            P16Job& J = jobs.j[jobs.count++];
            J.src = p.A; J.ld = p.lda; J.rows = p.M; J.cols = p.K; J.scale = 1.0f; J.hi = const_cast<void*>(aplane[i]);
            if ((size_t)p.M * p.K > max_elems) max_elems = (size_t)p.M * p.K;
        }
        if (wplane[i] == nullptr) {
            wplane[i] = ws + off;
            off += up256((size_t)p.N * p.K * 4);
            P16Job& J = jobs.j[jobs.count++];
            J.src = p.W; J.ld = p.ldw; J.rows = p.N; J.cols = p.K; J.scale = precision ? 1.0f : G16_W_SCALE; J.hi = const_cast<void*>(wplane[i]);
            if ((size_t)p.N * p.K > max_elems) max_elems = (size_t)p.N * p.K;
        }
        if (int rc = g16_make_map(&L.amap[i], aplane[i], precision, p.K, p.M, G16_BM)) return rc;
        if (int rc = g16_make_map(&L.bmap[i], wplane[i], precision, p.K, p.N, bn)) return rc;
        G16Problem& q = L.p[i];
        q.M = p.M; q.N = p.N; q.K = p.K; q.relu = p.relu; q.ldc = p.ldc; q.bias = p.bias; q.C = p.C; q.out16 = nullptr;
        q.m_tiles = cdiv(p.M, G16_BM); q.n_tiles = cdiv(p.N, bn); q.tile_begin = begin;
        begin += q.m_tiles * q.n_tiles;
    }
    L.count = grp.count;
    L.acc_scale = precision ? 1.0f : 1.0f / G16_W_SCALE;
    {
        int gx = (int)((max_elems / 8 + 255) / 256);
        const int cap = 4 * num_sms();
        if (gx > cap) gx = cap;
        if (gx < 1) gx = 1;
        dim3 grid(gx, jobs.count);
        if (precision) pack16x_kernel<1><<<grid, 256, 0, stream>>>(jobs);
        else           pack16x_kernel<0><<<grid, 256, 0, stream>>>(jobs);
        TG_LAUNCH_OK();
    }
    if (precision) return bn == 256 ? g16_launch_t<1, 256>(L, begin, stream) : g16_launch_t<1, 128>(L, begin, stream);
    return bn == 256 ? g16_launch_t<0, 256>(L, begin, stream) : g16_launch_t<0, 128>(L, begin, stream);
}

}  // namespace tg
