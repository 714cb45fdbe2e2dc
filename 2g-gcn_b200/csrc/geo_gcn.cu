// Geometry-level GCN (K-A): BatchNorm -> 1x1 conv 4->64 -> ReLU -> 1x1 conv 64->64 -> ReLU ->
// adaptive adjacency S = softmax(theta(e)^T phi(e)) -> (S e) W, stored as (B,128,V,T) so that the
// reference's un-permuted .view (vhoi/models.py:644-645) is a free reinterpretation.
// Reference: pyrutils/torch/models_gcn.py:30-100, input split vhoi/models.py:631-642.
//
// One CTA handles TT consecutive frames of one video entirely in shared memory; HBM traffic is the
// 16*V input bytes and 512*V output bytes per frame plus 112 KB of weights that stay L2-resident.
#include "common.cuh"
#include "recurrent_res.cuh"

namespace tg {

constexpr int GCN_TT = 4;        // frames per CTA
constexpr int GCN_THREADS = 256;
constexpr int GCN_LDE = 72;      // embedding row stride (64 + 8: 64-bit MMA fragment loads of a half-warp hit 16 distinct slots)
constexpr int GCN_LDT = 264;     // theta|phi row stride (256 + 8, same property)
constexpr int GCN_LDO = 129;     // output staging row stride
constexpr int GCN_LDS = 32;      // adjacency row stride (V <= 32)

// shared-memory rows: the tile's TT*V rows rounded up to whole 16-row MMA tiles, and far enough that the per-frame logit
// tiles (16-row / 8-column tiles starting at a frame's first row) stay inside
__host__ __device__ constexpr int gcn_rows_alloc(int V) {
    return ((GCN_TT * V + 15) / 16 * 16) > ((GCN_TT - 1) * V + 32) ? ((GCN_TT * V + 15) / 16 * 16) : ((GCN_TT - 1) * V + 32);
}

struct GcnParams {
    const float* xh;       // (B,T,H,Fh)
    const float* mean;     // (4V) running or batch mean
    const float* var;      // (4V) running or (biased) batch variance
    const float* gamma;    // (4V)
    const float* beta;     // (4V)
    const float* w1;       // (64,4)
    const float* b1;       // (64)
    const float* w3;       // (64,64)
    const float* b3;       // (64)
    const float* ws1;      // (128,64)
    const float* bs1;      // (128)
    const float* ws2;      // (128,64)
    const float* bs2;      // (128)
    const float* wg;       // (64,128)
    float* out;            // (B,128,V,T)
    int B, T, H, V, Fh;
};

// Per-channel batch statistics over (B,T) for training-mode BatchNorm1d(4V) (models_gcn.py:43-49),
// plus the running-stat update (momentum 0.1, unbiased variance).  One CTA per node v (4 channels).
__global__ void __launch_bounds__(256) geo_bn_stats_kernel(const float* __restrict__ xh, float* __restrict__ mean_out,
                                                          float* __restrict__ var_out, float* running_mean,
                                                          float* running_var, long long* num_batches, int N, int H,
                                                          int V, int Fh) {
    const int v = blockIdx.x;
    double s[4] = {0, 0, 0, 0}, q[4] = {0, 0, 0, 0};
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        const float4 x = *reinterpret_cast<const float4*>(xh + (size_t)n * H * Fh + 2048 + v * 4);
        s[0] += x.x; q[0] += (double)x.x * x.x;
        s[1] += x.y; q[1] += (double)x.y * x.y;
        s[2] += x.z; q[2] += (double)x.z * x.z;
        s[3] += x.w; q[3] += (double)x.w * x.w;
    }
    __shared__ double sh[8][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s[c] += __shfl_xor_sync(0xffffffffu, s[c], o);
            q[c] += __shfl_xor_sync(0xffffffffu, q[c], o);
        }
        if (lane == 0) { sh[warp][c] = s[c]; sh[warp][4 + c] = q[c]; }
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        const int c = threadIdx.x;
        double ss = 0, qq = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { ss += sh[w][c]; qq += sh[w][4 + c]; }
        const double m = ss / N;
        double var = qq / N - m * m;
        if (var < 0) var = 0;
        const int ch = c * V + v;              // channel index of x.view(bs, 4*V, step), models_gcn.py:46
        mean_out[ch] = (float)m;
        var_out[ch] = (float)var;
        const double unbiased = N > 1 ? var * N / (N - 1) : var;
        running_mean[ch] = (float)(0.9 * running_mean[ch] + 0.1 * m);
        running_var[ch] = (float)(0.9 * running_var[ch] + 0.1 * unbiased);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && num_batches) *num_batches += 1;
}

// ---- tensor-core building blocks (3xFP16 split on mma.sync.m16n8k16, fp32 accumulate: recurrent_res.cuh) -----------------------
// A fragments of one 16-row tile over K = 64 columns of an fp32 shared-memory matrix (row stride lda ≡ 8 mod 32 words: the
// 64-bit loads of a half-warp hit 16 distinct 8-byte slots), split in registers.
__device__ __forceinline__ void gcn_load_a(const float* A, int lda, int g8, int t4, uint32_t (&ah)[4][4], uint32_t (&al)[4][4]) {
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
        const float* ap = A + g8 * lda + ks * 16 + 2 * t4;
        const float2 x0 = *reinterpret_cast<const float2*>(ap);
        const float2 x1 = *reinterpret_cast<const float2*>(ap + 8 * lda);
        const float2 x2 = *reinterpret_cast<const float2*>(ap + 8);
        const float2 x3 = *reinterpret_cast<const float2*>(ap + 8 * lda + 8);
        split_f16x2(x0.x, x0.y, ah[ks][0], al[ks][0]);
        split_f16x2(x1.x, x1.y, ah[ks][1], al[ks][1]);
        split_f16x2(x2.x, x2.y, ah[ks][2], al[ks][2]);
        split_f16x2(x3.x, x3.y, ah[ks][3], al[ks][3]);
    }
}

// C[M x N] = act(A[M x 64] W^T + bias) with A in shared memory, W element (n, k) at W[n * sn + k * sk] in global memory
// (nn.Linear / 1x1-conv layout: sn = 64, sk = 1; the GCN weight (64,128) used as x W: sn = 1, sk = 128), C in shared memory.
// Warp w owns the n8 column tiles w, w + 8, ... (NTW of them): their weight fragments are loaded, scaled by 2^8 and split once
// and kept in registers while the warp walks over the 16-row tiles.
template <int NTW, bool RELU>
__device__ __forceinline__ void gcn_linear64(const float* A, int lda, int mtiles, int rows, const float* __restrict__ W, int sn, int sk,
                                             const float* __restrict__ bias, float* Cs, int ldc) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g8 = lane >> 2, t4 = lane & 3;
    uint32_t bh[NTW][4][2], bl[NTW][4][2];
    float bv[NTW][2];
#pragma unroll
    for (int j = 0; j < NTW; ++j) {
        const int n = (warp + 8 * j) * 8 + g8;                 // B fragment: column n = g8, k = 2 t4 + {0,1} (+8)
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            const float* wp = W + (size_t)n * sn + (size_t)(ks * 16 + 2 * t4) * sk;
            split_f16x2(__ldg(wp) * RES_WSCALE, __ldg(wp + sk) * RES_WSCALE, bh[j][ks][0], bl[j][ks][0]);
            split_f16x2(__ldg(wp + 8 * sk) * RES_WSCALE, __ldg(wp + 9 * sk) * RES_WSCALE, bh[j][ks][1], bl[j][ks][1]);
        }
        const int nc = (warp + 8 * j) * 8 + 2 * t4;            // C fragment columns 2 t4, 2 t4 + 1
        bv[j][0] = bias != nullptr ? __ldg(bias + nc) : 0.0f;
        bv[j][1] = bias != nullptr ? __ldg(bias + nc + 1) : 0.0f;
    }
    for (int mt = 0; mt < mtiles; ++mt) {
        uint32_t ah[4][4], al[4][4];
        gcn_load_a(A + mt * 16 * lda, lda, g8, t4, ah, al);
#pragma unroll
        for (int j = 0; j < NTW; ++j) {
            float c[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                mma_f16(c, al[ks], bh[j][ks]);
                mma_f16(c, ah[ks], bl[j][ks]);
                mma_f16(c, ah[ks], bh[j][ks]);
            }
            const int r0 = mt * 16 + g8, nc = (warp + 8 * j) * 8 + 2 * t4;
            float v0 = c[0] * (1.0f / RES_WSCALE) + bv[j][0], v1 = c[1] * (1.0f / RES_WSCALE) + bv[j][1];
            float v2 = c[2] * (1.0f / RES_WSCALE) + bv[j][0], v3 = c[3] * (1.0f / RES_WSCALE) + bv[j][1];
            if (RELU) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); v2 = fmaxf(v2, 0.f); v3 = fmaxf(v3, 0.f); }
            if (r0 < rows) { Cs[r0 * ldc + nc] = v0; Cs[r0 * ldc + nc + 1] = v1; }
            if (r0 + 8 < rows) { Cs[(r0 + 8) * ldc + nc] = v2; Cs[(r0 + 8) * ldc + nc + 1] = v3; }
        }
    }
}

__global__ void __launch_bounds__(GCN_THREADS, 1) geo_gcn_kernel(const GcnParams p) {
    extern __shared__ __align__(16) float smem[];
    const int V = p.V, T = p.T;
    const int tiles = (T + GCN_TT - 1) / GCN_TT;
    const int b = blockIdx.x / tiles, t0 = (blockIdx.x % tiles) * GCN_TT;
    const int nf = min(GCN_TT, T - t0);
    const int rows = nf * V;
    const int mtiles = (rows + 15) / 16;
    const int ra = gcn_rows_alloc(V);
    const int tid = threadIdx.x;

    float* g = smem;                                   // [ra][4]
    float* e1 = g + ra * 4;                            // [ra][72]   (later: S e)
    float* e = e1 + ra * GCN_LDE;                      // [ra][72]
    float* thph = e + ra * GCN_LDE;                    // [ra][264]  theta | phi (later: output staging)
    float* S = thph + ra * GCN_LDT;                    // [ra][32]

    // -- stage 0: gather geometry of human 0 and batch-normalise ---------------------------------
    for (int idx = tid; idx < rows * 4; idx += GCN_THREADS) {
        const int r = idx >> 2, c = idx & 3;
        const int f = r / V, v = r - f * V;
        const float x = p.xh[((size_t)(b * T + t0 + f) * p.H) * p.Fh + 2048 + v * 4 + c];
        const int ch = c * V + v;
        const float inv = 1.0f / sqrtf(p.var[ch] + 1e-5f);
        g[idx] = (x - p.mean[ch]) * inv * p.gamma[ch] + p.beta[ch];
    }
    // rows beyond the tile's frames feed padded MMA rows only; keep them finite
    for (int idx = rows * GCN_LDE + tid; idx < ra * GCN_LDE; idx += GCN_THREADS) { e1[idx] = 0.0f; e[idx] = 0.0f; }
    for (int idx = rows * GCN_LDT + tid; idx < ra * GCN_LDT; idx += GCN_THREADS) thph[idx] = 0.0f;
    __syncthreads();
    // -- stage 1: e1 = relu(W1 g + b1)  (K = 4: FFMA) ----------------------------------------------
    for (int idx = tid; idx < rows * 64; idx += GCN_THREADS) {
        const int r = idx >> 6, n = idx & 63;
        const float4 w = *reinterpret_cast<const float4*>(p.w1 + n * 4);
        const float4 x = *reinterpret_cast<const float4*>(g + r * 4);
        float a = p.b1[n];
        a = fmaf(w.x, x.x, a); a = fmaf(w.y, x.y, a); a = fmaf(w.z, x.z, a); a = fmaf(w.w, x.w, a);
        e1[r * GCN_LDE + n] = fmaxf(a, 0.0f);
    }
    __syncthreads();
    // -- stage 2: e = relu(W3 e1 + b3) on the tensor cores ------------------------------------------
    gcn_linear64<1, true>(e1, GCN_LDE, mtiles, rows, p.w3, 64, 1, p.b3, e, GCN_LDE);
    __syncthreads();
    // -- stage 3: theta = Ws1 e + bs1 (cols 0..127), phi = Ws2 e + bs2 (cols 128..255) --------------
    gcn_linear64<2, false>(e, GCN_LDE, mtiles, rows, p.ws1, 64, 1, p.bs1, thph, GCN_LDT);
    gcn_linear64<2, false>(e, GCN_LDE, mtiles, rows, p.ws2, 64, 1, p.bs2, thph + 128, GCN_LDT);
    __syncthreads();
    // -- stage 4: logits S[i][j] = theta_i . phi_j per frame (no 1/sqrt(d), models_gcn.py:97-99): both operands are
    //    activations, split in registers; (frame, 16-row tile) pairs are dealt to the warps -------------------------------
    {
        const int lane = tid & 31, warp = tid >> 5, g8 = lane >> 2, t4 = lane & 3;
        const int mt_f = (V + 15) / 16, nt_f = (V + 7) / 8;
        for (int job = warp; job < nf * mt_f; job += GCN_THREADS / 32) {
            const int f = job / mt_f, mt = job - f * mt_f;
            const float* th = thph + (f * V + mt * 16) * GCN_LDT;
            float c[4][4];
#pragma unroll
            for (int j = 0; j < 4; ++j) c[j][0] = c[j][1] = c[j][2] = c[j][3] = 0.0f;
#pragma unroll 1
            for (int kh = 0; kh < 2; ++kh) {                    // K = 128 in two halves of 64
                uint32_t ah[4][4], al[4][4];
                gcn_load_a(th + kh * 64, GCN_LDT, g8, t4, ah, al);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (j >= nt_f) continue;
                    const float* ph = thph + (f * V + j * 8 + g8) * GCN_LDT + 128 + kh * 64 + 2 * t4;
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        uint32_t bh[2], bl[2];
                        const float2 y0 = *reinterpret_cast<const float2*>(ph + ks * 16);
                        const float2 y1 = *reinterpret_cast<const float2*>(ph + ks * 16 + 8);
                        split_f16x2(y0.x, y0.y, bh[0], bl[0]);
                        split_f16x2(y1.x, y1.y, bh[1], bl[1]);
                        mma_f16(c[j], al[ks], bh);
                        mma_f16(c[j], ah[ks], bl);
                        mma_f16(c[j], ah[ks], bh);
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (j >= nt_f) continue;
                const int i0 = mt * 16 + g8, j0 = j * 8 + 2 * t4;
                if (i0 < V) {
                    if (j0 < V) S[(f * V + i0) * GCN_LDS + j0] = c[j][0];
                    if (j0 + 1 < V) S[(f * V + i0) * GCN_LDS + j0 + 1] = c[j][1];
                }
                if (i0 + 8 < V) {
                    if (j0 < V) S[(f * V + i0 + 8) * GCN_LDS + j0] = c[j][2];
                    if (j0 + 1 < V) S[(f * V + i0 + 8) * GCN_LDS + j0 + 1] = c[j][3];
                }
            }
        }
    }
    __syncthreads();
    // -- stage 5: row softmax, one warp per row ----------------------------------------------------
    {
        const int lane = tid & 31, warp = tid >> 5;
        for (int r = warp; r < rows; r += GCN_THREADS / 32) {
            const float v = lane < V ? S[r * GCN_LDS + lane] : -INFINITY;
            const float m = warp_max(v);
            const float ex = lane < V ? expf(v - m) : 0.0f;
            const float s = warp_sum(ex);
            if (lane < V) S[r * GCN_LDS + lane] = ex / s;
        }
    }
    __syncthreads();
    // -- stage 6: Se = S e  (into the e1 buffer; K = V <= 32: FFMA) ---------------------------------
    {
        const int c = tid & 63, rg = tid >> 6;
        for (int r = rg; r < rows; r += GCN_THREADS / 64) {
            const int f = r / V;
            const float* srow = S + r * GCN_LDS;
            const float* eb = e + (f * V) * GCN_LDE + c;
            float a = 0.0f;
            for (int j = 0; j < V; ++j) a = fmaf(srow[j], eb[j * GCN_LDE], a);
            e1[r * GCN_LDE + c] = a;
        }
    }
    __syncthreads();
    // -- stage 7: out = Se Wg (64 -> 128) on the tensor cores, staged in the theta|phi buffer --------
    gcn_linear64<2, false>(e1, GCN_LDE, mtiles, rows, p.wg, 1, 128, nullptr, thph, GCN_LDO);
    __syncthreads();
    // -- stage 8: store (B,128,V,T): t fastest ------------------------------------------------------
    for (int idx = tid; idx < 128 * V * nf; idx += GCN_THREADS) {
        const int f = idx % nf;
        const int nv = idx / nf;
        const int v = nv % V, n = nv / V;
        p.out[((size_t)(b * 128 + n) * V + v) * T + t0 + f] = thph[(f * V + v) * GCN_LDO + n];
    }
}

size_t geo_gcn_smem_bytes(int V) {
    return sizeof(float) * (size_t)gcn_rows_alloc(V) * (4 + 2 * GCN_LDE + GCN_LDT + GCN_LDS);
}

// workspace: 2 * 4V floats for batch statistics
int launch_geo_gcn(const float* x_human, const void* const* w, float* out, float* bn_running_mean,
                   float* bn_running_var, int64_t* bn_num_batches, float* stats_ws, int B, int T, int H, int V, int Fh,
                   int bn_train, cudaStream_t stream) {
    TG_REQUIRE(V >= 1 && V <= 32, "geo_gcn: gcn_node=%d unsupported (1..32)", V);
    TG_REQUIRE(Fh == 2048 + 4 * V, "geo_gcn: x_human feature size %d != 2048 + 4*%d", Fh, V);
    GcnParams p;
    p.xh = x_human;
    p.gamma = (const float*)w[TGGCN_W_GCN_BN_W];
    p.beta = (const float*)w[TGGCN_W_GCN_BN_B];
    if (bn_train) {
        TG_REQUIRE(stats_ws && bn_running_mean && bn_running_var, "geo_gcn: bn_train needs stats workspace and running buffers");
        geo_bn_stats_kernel<<<V, 256, 0, stream>>>(x_human, stats_ws, stats_ws + 4 * V, bn_running_mean, bn_running_var,
                                                   (long long*)bn_num_batches, B * T, H, V, Fh);
        TG_LAUNCH_OK();
        p.mean = stats_ws;
        p.var = stats_ws + 4 * V;
    } else {
        p.mean = (const float*)w[TGGCN_W_GCN_BN_MEAN];
        p.var = (const float*)w[TGGCN_W_GCN_BN_VAR];
    }
    p.w1 = (const float*)w[TGGCN_W_GCN_C1_W]; p.b1 = (const float*)w[TGGCN_W_GCN_C1_B];
    p.w3 = (const float*)w[TGGCN_W_GCN_C3_W]; p.b3 = (const float*)w[TGGCN_W_GCN_C3_B];
    p.ws1 = (const float*)w[TGGCN_W_GCN_S1_W]; p.bs1 = (const float*)w[TGGCN_W_GCN_S1_B];
    p.ws2 = (const float*)w[TGGCN_W_GCN_S2_W]; p.bs2 = (const float*)w[TGGCN_W_GCN_S2_B];
    p.wg = (const float*)w[TGGCN_W_GCN_W];
    p.out = out;
    p.B = B; p.T = T; p.H = H; p.V = V; p.Fh = Fh;
    const size_t smem = geo_gcn_smem_bytes(V);
    if (int rc = ensure_smem((const void*)geo_gcn_kernel, smem)) return rc;
    const int tiles = cdiv(T, GCN_TT);
    geo_gcn_kernel<<<B * tiles, GCN_THREADS, smem, stream>>>(p);
    TG_LAUNCH_OK();
    return 0;
}

}  // namespace tg
