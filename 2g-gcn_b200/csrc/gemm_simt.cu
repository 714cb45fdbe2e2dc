// Grouped fp32 SIMT projection kernel (K-B, exact-fp32 path):
//   C[M,N] = act(A[M,K] * W[N,K]^T + bias[N])      (nn.Linear as built by build_mlp,
//                                                   pyrutils/torch/models.py:31-33)
// Several independent problems are batched into one launch (one CTA per output tile across all
// problems) so that the many small projections of a forward fill the 148 SMs together.
// fp32 FMA throughout: this is the path the "fp32-tolerance" parity tests run on; the tensor-core
// path (gemm_tc.cu) reproduces the same contract with tcgen05 3xTF32.
#include "common.cuh"
#include "gemm.h"

namespace tg {

template <int TM, int TN>
__global__ void __launch_bounds__(256) gemm_simt_kernel(const GemmGroup grp) {
    constexpr int BM = 16 * TM, BN = 16 * TN, BK = 16;
    constexpr int GM = TM / 4, GN = TN / 4;
    constexpr int LDA_S = BM + 4, LDB_S = BN + 4;
    __shared__ __align__(16) float As[2][BK][LDA_S];
    __shared__ __align__(16) float Bs[2][BK][LDB_S];

    // locate this CTA's problem and tile
    int pi = 0;
#pragma unroll 1
    for (int i = 1; i < grp.count; ++i)
        if ((int)blockIdx.x >= grp.p[i].tile_begin) pi = i;
    const GemmProblem& P = grp.p[pi];
    const int tile = blockIdx.x - P.tile_begin;
    const int tiles_n = (P.N + BN - 1) / BN;
    const int m0 = (tile / tiles_n) * BM, n0 = (tile % tiles_n) * BN;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;

    // global -> register staging: each thread moves float4s along K
    constexpr int A_F4 = BM * BK / 4 / 256;   // float4 per thread for the A tile (2 or 1)
    constexpr int B_F4 = BN * BK / 4 / 256;
    float4 ra[A_F4], rb[B_F4];
    auto load_tiles = [&](int k0) {
#pragma unroll
        for (int i = 0; i < A_F4; ++i) {
            const int f = tid + i * 256;
            const int row = f >> 2, kq = (f & 3) * 4;
            const int m = m0 + row;
            ra[i] = m < P.M ? __ldg(reinterpret_cast<const float4*>(P.A + (size_t)m * P.lda + k0 + kq))
                            : make_float4(0.f, 0.f, 0.f, 0.f);
            if (P.amask != nullptr && m < P.M) {
                const float4 mk = __ldg(reinterpret_cast<const float4*>(P.amask + (size_t)m * P.ldm + k0 + kq));
                ra[i].x = mk.x > 0.f ? ra[i].x : 0.f; ra[i].y = mk.y > 0.f ? ra[i].y : 0.f;
                ra[i].z = mk.z > 0.f ? ra[i].z : 0.f; ra[i].w = mk.w > 0.f ? ra[i].w : 0.f;
            }
        }
#pragma unroll
        for (int i = 0; i < B_F4; ++i) {
            const int f = tid + i * 256;
            const int row = f >> 2, kq = (f & 3) * 4;
            const int n = n0 + row;
            rb[i] = n < P.N ? __ldg(reinterpret_cast<const float4*>(P.W + (size_t)n * P.ldw + k0 + kq))
                            : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    auto store_tiles = [&](int buf) {
#pragma unroll
        for (int i = 0; i < A_F4; ++i) {
            const int f = tid + i * 256;
            const int row = f >> 2, kq = (f & 3) * 4;
            As[buf][kq + 0][row] = ra[i].x; As[buf][kq + 1][row] = ra[i].y;
            As[buf][kq + 2][row] = ra[i].z; As[buf][kq + 3][row] = ra[i].w;
        }
#pragma unroll
        for (int i = 0; i < B_F4; ++i) {
            const int f = tid + i * 256;
            const int row = f >> 2, kq = (f & 3) * 4;
            Bs[buf][kq + 0][row] = rb[i].x; Bs[buf][kq + 1][row] = rb[i].y;
            Bs[buf][kq + 2][row] = rb[i].z; Bs[buf][kq + 3][row] = rb[i].w;
        }
    };

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.0f;

    const int nk = P.K / BK;
    load_tiles(0);
    store_tiles(0);
    __syncthreads();
    for (int kt = 0; kt < nk; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < nk) load_tiles((kt + 1) * BK);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float a[TM], b[TN];
#pragma unroll
            for (int g = 0; g < GM; ++g) {
                const float4 v = *reinterpret_cast<const float4*>(&As[buf][k][g * (BM / GM) + ty * 4]);
                a[g * 4 + 0] = v.x; a[g * 4 + 1] = v.y; a[g * 4 + 2] = v.z; a[g * 4 + 3] = v.w;
            }
#pragma unroll
            for (int g = 0; g < GN; ++g) {
                const float4 v = *reinterpret_cast<const float4*>(&Bs[buf][k][g * (BN / GN) + tx * 4]);
                b[g * 4 + 0] = v.x; b[g * 4 + 1] = v.y; b[g * 4 + 2] = v.z; b[g * 4 + 3] = v.w;
            }
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (kt + 1 < nk) {
            store_tiles(buf ^ 1);
            __syncthreads();
        }
    }

    // epilogue: bias, activation, store (rows/cols of the split register tile)
#pragma unroll
    for (int gi = 0; gi < GM; ++gi)
#pragma unroll
        for (int ii = 0; ii < 4; ++ii) {
            const int m = m0 + gi * (BM / GM) + ty * 4 + ii;
            if (m >= P.M) continue;
#pragma unroll
            for (int gj = 0; gj < GN; ++gj) {
                const int n = n0 + gj * (BN / GN) + tx * 4;
                float v[4];
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    float x = acc[gi * 4 + ii][gj * 4 + jj];
                    if (P.bias != nullptr && n + jj < P.N) x += __ldg(P.bias + n + jj);
                    if (P.relu) x = fmaxf(x, 0.0f);
                    v[jj] = x;
                }
                float* dst = P.C + (size_t)m * P.ldc + n;
                if (P.beta) {
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj)
                        if (n + jj < P.N) v[jj] += dst[jj];
                }
                if (n + 3 < P.N && ((P.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
                    *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
                } else {
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj)
                        if (n + jj < P.N) dst[jj] = v[jj];
                }
            }
        }
}

int validate_problem(const GemmProblem& p) {
    TG_REQUIRE(p.M > 0 && p.N > 0 && p.K > 0, "gemm: empty problem (%d,%d,%d)", p.M, p.N, p.K);
    TG_REQUIRE(p.K % 16 == 0, "gemm: K=%d must be a multiple of 16", p.K);
    TG_REQUIRE(p.lda % 4 == 0 && p.ldw % 4 == 0, "gemm: lda=%d / ldw=%d must be multiples of 4", p.lda, p.ldw);
    TG_REQUIRE((reinterpret_cast<uintptr_t>(p.A) & 15) == 0 && (reinterpret_cast<uintptr_t>(p.W) & 15) == 0,
               "gemm: A and W must be 16-byte aligned");
    TG_REQUIRE(p.amask == nullptr || (p.ldm % 4 == 0 && (reinterpret_cast<uintptr_t>(p.amask) & 15) == 0),
               "gemm: mask must be 16-byte aligned with ldm a multiple of 4");
    return 0;
}

int launch_gemm_simt(GemmGroup& grp, cudaStream_t stream) {
    if (grp.count == 0) return 0;
    TG_REQUIRE(grp.count <= GEMM_MAX_PROBLEMS, "gemm: too many problems in one group (%d)", grp.count);
    long tiles128 = 0, tiles64 = 0;
    for (int i = 0; i < grp.count; ++i) {
        if (int rc = validate_problem(grp.p[i])) return rc;
        tiles128 += (long)cdiv(grp.p[i].M, 128) * cdiv(grp.p[i].N, 128);
        tiles64 += (long)cdiv(grp.p[i].M, 64) * cdiv(grp.p[i].N, 64);
    }
    // big tiles only when they still give every SM at least ~2 CTAs
    const bool big = tiles128 >= 2L * num_sms();
    const int bm = big ? 128 : 64;
    int begin = 0;
    for (int i = 0; i < grp.count; ++i) {
        grp.p[i].tile_begin = begin;
        begin += cdiv(grp.p[i].M, bm) * cdiv(grp.p[i].N, bm);
    }
    if (big) gemm_simt_kernel<8, 8><<<begin, 256, 0, stream>>>(grp);
    else     gemm_simt_kernel<4, 4><<<begin, 256, 0, stream>>>(grp);
    TG_LAUNCH_OK();
    (void)tiles64;
    return 0;
}

}  // namespace tg
