// Problem descriptors shared by the SIMT and tcgen05 projection kernels.
#pragma once
#include <cuda_runtime.h>

namespace tg {

constexpr int GEMM_MAX_PROBLEMS = 8;

// C[M,N] = act(A[M,K] W[N,K]^T + bias[N]); A, W are K-major (row-major with leading dims lda, ldw).
struct GemmProblem {
    const float* A;
    const float* W;
    const float* bias;   // may be null
    float* C;
    int M, N, K;
    int lda, ldw, ldc;
    int relu;
    int tile_begin;      // filled by the launcher
    // backward use (gemm_bwd.cu): A is read as A[m][k] * (amask[m][k] > 0) and the result is added to C when beta != 0
    const float* amask;  // may be null
    int ldm;
    int beta;
    int ksplit;          // tcgen05 path: K range split over this many CTAs per tile, combined with atomics (filled by the launcher)
};

struct GemmGroup {
    GemmProblem p[GEMM_MAX_PROBLEMS];
    int count;
    int precision;       // tcgen05 path: 0 = 3xTF32 split (fp32-class), 1 = bf16 operands, fp32 accumulate (set by launch_gemm)
};

inline void gemm_add(GemmGroup& g, const float* A, int lda, const float* W, int ldw, const float* bias, float* C,
                     int ldc, int M, int N, int K, int relu) {
    GemmProblem& q = g.p[g.count++];
    q.A = A; q.W = W; q.bias = bias; q.C = C;
    q.M = M; q.N = N; q.K = K; q.lda = lda; q.ldw = ldw; q.ldc = ldc; q.relu = relu; q.tile_begin = 0;
    q.amask = nullptr; q.ldm = 0; q.beta = 0; q.ksplit = 1;
}

int launch_gemm_simt(GemmGroup& grp, cudaStream_t stream);
// TMA-fed tcgen05 kernel on 16-bit operand planes (gemm16.cu): precision 0 = fp16 (hi, lo) split (fp32-class), 1 = bf16
bool gemm16_eligible(const GemmGroup& grp);
size_t gemm16_scratch_bytes(const GemmGroup& grp);
int launch_gemm16(GemmGroup& grp, int precision, void* scratch, size_t scratch_bytes, unsigned int* err, cudaStream_t stream);

// General form (backward): C[M,N] (+)= act(A' B'^T + bias), A' / B' = 16-bit planes packed from fp32 sources.
struct Gemm16Operand {
    const float* src;       // fp32 source matrix [rows][cols], row stride ld
    int ld, rows, cols;
    const float* mask;      // optional ReLU mask of the same shape (value kept where mask > 0), row stride ldm
    int ldm;
    int transpose;          // 0: operand rows = source rows, reduction = source columns (cols % 64 == 0)
                            // 1: operand rows = source columns, reduction = source rows (zero-padded to a multiple of 64)
    int shift, period;      // transpose only: source row m + shift pairs with reduction index m when 0 <= m % period + shift < period
    int dynamic;            // 1: gradient operand — scaled by the power of two that brings its largest magnitude to [2^13, 2^14)
    float scale;            // fixed scale otherwise (0 = 1); the fixed scales of all problems of a group must agree
    float* colsum;          // transpose only, optional: colsum[c] (+)= sum_m src[m][c] * [mask > 0] (the bias gradient), fused into the pack
    int colsum_beta;        // 0: colsum is zeroed first, 1: accumulated into
};
struct Gemm16Problem {
    Gemm16Operand a, b;     // M = operand rows of a, N = operand rows of b (N % 16 == 0)
    const float* bias;
    float* C;
    int ldc, relu, beta;
};
bool gemm16_enabled();
bool gemm16_eligible(const Gemm16Problem* p, int count);
size_t gemm16_scratch_bytes(const Gemm16Problem* p, int count);
int launch_gemm16(const Gemm16Problem* p, int count, int precision, void* scratch, size_t scratch_bytes, unsigned int* err,
                  cudaStream_t stream);
int launch_gemm_tc(GemmGroup& grp, cudaStream_t stream);     // tcgen05 3xTF32 or bf16 (gemm_tc.cu)
// path 0: fp32 SIMT; 1: tcgen05 3xTF32 (K % 32 == 0 required); 2: tcgen05 where the shapes allow it, SIMT otherwise;
// 3: like 2 with bf16 operands on the tensor-core path (dims.precision = 1)
inline int launch_gemm(GemmGroup& grp, int path, cudaStream_t stream) {
    grp.precision = path == 3 ? 1 : 0;
    if (path == 2 || path == 3) {
        path = 1;
        for (int i = 0; i < grp.count; ++i)
            if (grp.p[i].K % 32 != 0) path = 0;
    }
    return path == 1 ? launch_gemm_tc(grp, stream) : launch_gemm_simt(grp, stream);
}

}  // namespace tg
