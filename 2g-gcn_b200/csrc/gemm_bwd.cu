// Backward of nn.Linear (+ReLU) (pyrutils/torch/models.py:31-33, autograd of y = act(x W^T + b)):
//   Z  = dY (.) [Y > 0]                      (mask applied on load, never materialised)
//   dX = Z W            -> the forward NT kernels on (Z, W^T)          (W^T from transpose_kernel)
//   dW = Z^T X          -> gemm_tn_kernel  (reduction over the row index, both operands row-major)
//   db = column sums of Z -> colsum_kernel
// The TN kernel optionally shifts the rows of X by a fixed amount inside blocks of `period` rows: the weight
// gradient of a recurrence, dW_hh = sum_t dG_t^T h_{t-1}, is then one GEMM over all (video, t, entity) rows.
#include "common.cuh"
#include "gemm.h"

namespace tg {

struct TnProblem {
    const float* Z; int ldz;        // (M, N) upstream gradient
    const float* mask; int ldm;     // (M, N) forward output (ReLU mask) or null
    const float* A; int lda;        // (M, K) forward input
    float* C; int ldc;              // (N, K) weight gradient
    int M, N, K;
    int a_shift, period;            // row m of Z pairs with row m + a_shift of A when 0 <= m % period + a_shift < period
    int beta;
};

__global__ void __launch_bounds__(256) gemm_tn_kernel(const TnProblem P) {
    constexpr int BN = 128, BK = 128, BM = 16, LD = BN + 4;
    __shared__ __align__(16) float Zs[2][BM][LD];
    __shared__ __align__(16) float As[2][BM][LD];
    const int tiles_k = (P.K + BK - 1) / BK;
    const int n0 = (blockIdx.x / tiles_k) * BN, k0 = (blockIdx.x % tiles_k) * BK;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    float4 rz[2], ra[2];
    auto load = [&](int m0) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int f = tid + i * 256;
            const int row = f >> 5, c4 = (f & 31) * 4;
            const int m = m0 + row;
            float4 z = make_float4(0.f, 0.f, 0.f, 0.f), a = make_float4(0.f, 0.f, 0.f, 0.f);
            if (m < P.M) {
                if (n0 + c4 < P.N) {      // N % 4 == 0
                    z = __ldg(reinterpret_cast<const float4*>(P.Z + (size_t)m * P.ldz + n0 + c4));
                    if (P.mask != nullptr) {
                        const float4 mk = __ldg(reinterpret_cast<const float4*>(P.mask + (size_t)m * P.ldm + n0 + c4));
                        z.x = mk.x > 0.f ? z.x : 0.f; z.y = mk.y > 0.f ? z.y : 0.f;
                        z.z = mk.z > 0.f ? z.z : 0.f; z.w = mk.w > 0.f ? z.w : 0.f;
                    }
                }
                int ma = m;
                bool ok = true;
                if (P.a_shift != 0) {
                    const int pos = m % P.period + P.a_shift;
                    ok = pos >= 0 && pos < P.period;
                    ma = m + P.a_shift;
                }
                if (ok && k0 + c4 < P.K) a = __ldg(reinterpret_cast<const float4*>(P.A + (size_t)ma * P.lda + k0 + c4));
            }
            rz[i] = z; ra[i] = a;
        }
    };
    auto store = [&](int buf) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int f = tid + i * 256;
            const int row = f >> 5, c4 = (f & 31) * 4;
            *reinterpret_cast<float4*>(&Zs[buf][row][c4]) = rz[i];
            *reinterpret_cast<float4*>(&As[buf][row][c4]) = ra[i];
        }
    };
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;
    const int nm = (P.M + BM - 1) / BM;
    load(0);
    store(0);
    __syncthreads();
    for (int mt = 0; mt < nm; ++mt) {
        const int buf = mt & 1;
        if (mt + 1 < nm) load((mt + 1) * BM);
#pragma unroll
        for (int r = 0; r < BM; ++r) {
            float a[8], b[8];
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                const float4 v = *reinterpret_cast<const float4*>(&Zs[buf][r][g * 64 + ty * 4]);
                a[g * 4 + 0] = v.x; a[g * 4 + 1] = v.y; a[g * 4 + 2] = v.z; a[g * 4 + 3] = v.w;
                const float4 w = *reinterpret_cast<const float4*>(&As[buf][r][g * 64 + tx * 4]);
                b[g * 4 + 0] = w.x; b[g * 4 + 1] = w.y; b[g * 4 + 2] = w.z; b[g * 4 + 3] = w.w;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (mt + 1 < nm) {
            store(buf ^ 1);
            __syncthreads();
        }
    }
#pragma unroll
    for (int gi = 0; gi < 2; ++gi)
#pragma unroll
        for (int ii = 0; ii < 4; ++ii) {
            const int n = n0 + gi * 64 + ty * 4 + ii;
            if (n >= P.N) continue;
#pragma unroll
            for (int gj = 0; gj < 2; ++gj)
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const int k = k0 + gj * 64 + tx * 4 + jj;
                    if (k >= P.K) continue;
                    float* dst = P.C + (size_t)n * P.ldc + k;
                    const float v = acc[gi * 4 + ii][gj * 4 + jj];
                    *dst = P.beta ? *dst + v : v;
                }
        }
}

// out[n] (+)= sum_m Z[m][n] * [mask[m][n] > 0]; rows are split over blockIdx.y and combined with atomics
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ Z, int ldz, const float* __restrict__ mask, int ldm,
                                                     float* __restrict__ out, int M, int N) {
    __shared__ float part[8][33];
    const int n = blockIdx.x * 32 + (threadIdx.x & 31), ry = threadIdx.x >> 5;
    float s = 0.0f;
    if (n < N)
        for (int m = blockIdx.y * 8 + ry; m < M; m += 8 * gridDim.y) {
            float z = __ldg(Z + (size_t)m * ldz + n);
            if (mask != nullptr && !(__ldg(mask + (size_t)m * ldm + n) > 0.f)) z = 0.0f;
            s += z;
        }
    part[ry][threadIdx.x & 31] = s;
    __syncthreads();
    if (ry == 0 && n < N) {
        float t = 0.0f;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += part[i][threadIdx.x];
        atomicAdd(out + n, t);
    }
}

// out[c][r] = in[r][c]
__global__ void __launch_bounds__(256) transpose_kernel(const float* __restrict__ in, int ldi, float* __restrict__ out, int ldo, int R,
                                                        int Cc) {
    __shared__ float t[32][33];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    const int x = threadIdx.x & 31, y = threadIdx.x >> 5;
    for (int j = y; j < 32; j += 8)
        if (r0 + j < R && c0 + x < Cc) t[j][x] = __ldg(in + (size_t)(r0 + j) * ldi + c0 + x);
    __syncthreads();
    for (int j = y; j < 32; j += 8)
        if (c0 + j < Cc && r0 + x < R) out[(size_t)(c0 + j) * ldo + r0 + x] = t[x][j];
}

// out[c][m] = in[m + shift][c] * [mask[m + shift][c] > 0]   for m < M (0 where the shifted row leaves its block of `period`
// rows), and 0 for M <= m < Mp: the reduction index of a weight gradient becomes the contiguous (K-major) axis, so that
// dW = Z^T X runs on the tcgen05 NT projection kernel as  dW[N,K] = Zt[N,Mp] Xt[K,Mp]^T.  Optionally accumulates the column
// sums of the (masked) input with atomics: the bias gradient, for free.
__global__ void __launch_bounds__(256) transpose_prep_kernel(const float* __restrict__ in, int ldi, const float* __restrict__ mask,
                                                             int ldm, float* __restrict__ out, int ldo, int M, int Mp, int Cc, int shift,
                                                             int period, float* __restrict__ colsum) {
    __shared__ float t[32][33];
    const int c0 = blockIdx.x * 32, m0 = blockIdx.y * 32;
    const int x = threadIdx.x & 31, y = threadIdx.x >> 5;
    for (int j = y; j < 32; j += 8) {
        const int m = m0 + j, c = c0 + x;
        float v = 0.0f;
        if (m < M && c < Cc) {
            int ms = m;
            bool ok = true;
            if (shift != 0) {
                const int pos = m % period + shift;
                ok = pos >= 0 && pos < period;
                ms = m + shift;
            }
            if (ok) {
                v = __ldg(in + (size_t)ms * ldi + c);
                if (mask != nullptr && !(__ldg(mask + (size_t)ms * ldm + c) > 0.0f)) v = 0.0f;
            }
        }
        t[j][x] = v;
    }
    __syncthreads();
    for (int j = y; j < 32; j += 8)
        if (c0 + j < Cc && m0 + x < Mp) out[(size_t)(c0 + j) * ldo + m0 + x] = t[x][j];
    if (colsum != nullptr && y == 0 && c0 + x < Cc) {
        float sacc = 0.0f;
#pragma unroll
        for (int j = 0; j < 32; ++j) sacc += t[j][x];
        atomicAdd(colsum + c0 + x, sacc);
    }
}

int launch_transpose_prep(const float* in, int ldi, const float* mask, int ldm, float* out, int Mp, int M, int C, int shift, int period,
                          float* colsum, cudaStream_t stream) {
    TG_REQUIRE(Mp >= M && Mp % 32 == 0, "transpose_prep: padded length %d must be a multiple of 32 >= %d", Mp, M);
    TG_REQUIRE(shift == 0 || period > 0, "transpose_prep: shift needs a period");
    transpose_prep_kernel<<<dim3(cdiv(C, 32), Mp / 32), 256, 0, stream>>>(in, ldi, mask, ldm, out, Mp, M, Mp, C, shift, period, colsum);
    TG_LAUNCH_OK();
    return 0;
}

int launch_transpose(const float* in, int ldi, float* out, int ldo, int R, int C, cudaStream_t stream) {
    transpose_kernel<<<dim3(cdiv(C, 32), cdiv(R, 32)), 256, 0, stream>>>(in, ldi, out, ldo, R, C);
    TG_LAUNCH_OK();
    return 0;
}

int launch_gemm_tn(const float* Z, int ldz, const float* mask, int ldm, const float* A, int lda, float* C, int ldc, int M, int N,
                   int K, int a_shift, int period, int beta, cudaStream_t stream) {
    TG_REQUIRE(M > 0 && N > 0 && K > 0, "gemm_tn: empty problem");
    TG_REQUIRE(N % 4 == 0 && K % 4 == 0 && ldz % 4 == 0 && lda % 4 == 0, "gemm_tn: N, K, ldz, lda must be multiples of 4");
    TG_REQUIRE((reinterpret_cast<uintptr_t>(Z) & 15) == 0 && (reinterpret_cast<uintptr_t>(A) & 15) == 0, "gemm_tn: operands must be 16-byte aligned");
    TG_REQUIRE(mask == nullptr || ((reinterpret_cast<uintptr_t>(mask) & 15) == 0 && ldm % 4 == 0), "gemm_tn: mask alignment");
    TG_REQUIRE(a_shift == 0 || period > 0, "gemm_tn: a_shift needs a period");
    TnProblem P;
    P.Z = Z; P.ldz = ldz; P.mask = mask; P.ldm = ldm; P.A = A; P.lda = lda; P.C = C; P.ldc = ldc;
    P.M = M; P.N = N; P.K = K; P.a_shift = a_shift; P.period = period; P.beta = beta;
    gemm_tn_kernel<<<cdiv(N, 128) * cdiv(K, 128), 256, 0, stream>>>(P);
    TG_LAUNCH_OK();
    return 0;
}

int launch_colsum(const float* Z, int ldz, const float* mask, int ldm, float* out, int M, int N, int beta, cudaStream_t stream) {
    if (!beta) TG_CUDA_OK(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)N, stream));
    int ysplit = cdiv(M, 64);
    if (ysplit > 64) ysplit = 64;
    colsum_kernel<<<dim3(cdiv(N, 32), ysplit), 256, 0, stream>>>(Z, ldz, mask, ldm, out, M, N);
    TG_LAUNCH_OK();
    return 0;
}

}  // namespace tg

using namespace tg;

extern "C" int tggcn_linear_bwd(const float* dY, int ldy, const float* Y, int ldyf, const float* X, int ldx, const float* W,
                                          int ldw, float* dX, int lddx, int beta_dx, float* dW, int lddw, float* db,
                                          float* wt_scratch, int M, int N, int K, int gemm_path, void* stream_) {
    TG_REQUIRE(dY && X && W, "linear_bwd: null pointer");
    cudaStream_t stream = (cudaStream_t)stream_;
    if (dX != nullptr) {
        TG_REQUIRE(wt_scratch != nullptr, "linear_bwd: dX needs a K*N float scratch for W^T");
        TG_REQUIRE(N % 16 == 0, "linear_bwd: dX needs N %% 16 == 0 (got %d)", N);
        if (int rc = launch_transpose(W, ldw, wt_scratch, N, N, K, stream)) return rc;     // (N,K) -> (K,N)
        GemmGroup g;
        g.count = 0;
        gemm_add(g, dY, ldy, wt_scratch, N, nullptr, dX, lddx, M, K, N, 0);
        g.p[0].amask = Y; g.p[0].ldm = ldyf; g.p[0].beta = beta_dx;
        if (int rc = launch_gemm(g, gemm_path, stream)) return rc;
    }
    if (dW != nullptr)
        if (int rc = launch_gemm_tn(dY, ldy, Y, ldyf, X, ldx, dW, lddw, M, N, K, 0, 0, 0, stream)) return rc;
    if (db != nullptr)
        if (int rc = launch_colsum(dY, ldy, Y, ldyf, db, M, N, 0, stream)) return rc;
    return 0;
}
