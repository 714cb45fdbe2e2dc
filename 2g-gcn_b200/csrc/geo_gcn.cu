// Geometry-level GCN (K-A): BatchNorm -> 1x1 conv 4->64 -> ReLU -> 1x1 conv 64->64 -> ReLU ->
// adaptive adjacency S = softmax(theta(e)^T phi(e)) -> (S e) W, stored as (B,128,V,T) so that the
// reference's un-permuted .view (vhoi/models.py:644-645) is a free reinterpretation.
// Reference: pyrutils/torch/models_gcn.py:30-100, input split vhoi/models.py:631-642.
//
// One CTA handles TT consecutive frames of one video entirely in shared memory; HBM traffic is the
// 16*V input bytes and 512*V output bytes per frame plus 112 KB of weights that stay L2-resident.
#include "common.cuh"

namespace tg {

constexpr int GCN_TT = 4;        // frames per CTA
constexpr int GCN_THREADS = 256;
constexpr int GCN_LDT = 257;     // theta|phi row stride (odd: conflict-free column walks)
constexpr int GCN_LDO = 129;     // output staging row stride
constexpr int GCN_LDS = 32;      // adjacency row stride (V <= 32)

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

__global__ void __launch_bounds__(GCN_THREADS, 1) geo_gcn_kernel(const GcnParams p) {
    extern __shared__ __align__(16) float smem[];
    const int V = p.V, T = p.T;
    const int tiles = (T + GCN_TT - 1) / GCN_TT;
    const int b = blockIdx.x / tiles, t0 = (blockIdx.x % tiles) * GCN_TT;
    const int nf = min(GCN_TT, T - t0);
    const int rows = nf * V;
    const int tid = threadIdx.x;

    float* g = smem;                                   // [TT*V][4]
    float* e1 = g + GCN_TT * V * 4;                    // [TT*V][64]   (later: S e)
    float* e = e1 + GCN_TT * V * 64;                   // [TT*V][64]
    float* thph = e + GCN_TT * V * 64;                 // [TT*V][257]  theta | phi (later: output staging)
    float* S = thph + GCN_TT * V * GCN_LDT;            // [TT*V][32]

    // -- stage 0: gather geometry of human 0 and batch-normalise ---------------------------------
    for (int idx = tid; idx < rows * 4; idx += GCN_THREADS) {
        const int r = idx >> 2, c = idx & 3;
        const int f = r / V, v = r - f * V;
        const float x = p.xh[((size_t)(b * T + t0 + f) * p.H) * p.Fh + 2048 + v * 4 + c];
        const int ch = c * V + v;
        const float inv = 1.0f / sqrtf(p.var[ch] + 1e-5f);
        g[idx] = (x - p.mean[ch]) * inv * p.gamma[ch] + p.beta[ch];
    }
    __syncthreads();
    // -- stage 1: e1 = relu(W1 g + b1) -------------------------------------------------------------
    for (int idx = tid; idx < rows * 64; idx += GCN_THREADS) {
        const int r = idx >> 6, n = idx & 63;
        const float4 w = *reinterpret_cast<const float4*>(p.w1 + n * 4);
        const float4 x = *reinterpret_cast<const float4*>(g + r * 4);
        float a = p.b1[n];
        a = fmaf(w.x, x.x, a); a = fmaf(w.y, x.y, a); a = fmaf(w.z, x.z, a); a = fmaf(w.w, x.w, a);
        e1[idx] = fmaxf(a, 0.0f);
    }
    __syncthreads();
    // -- stage 2: e = relu(W3 e1 + b3); each thread keeps one weight row in registers --------------
    {
        const int n = tid & 63, rg = tid >> 6;
        float w[64];
#pragma unroll
        for (int k = 0; k < 64; k += 4) {
            const float4 t4 = *reinterpret_cast<const float4*>(p.w3 + n * 64 + k);
            w[k] = t4.x; w[k + 1] = t4.y; w[k + 2] = t4.z; w[k + 3] = t4.w;
        }
        const float bias = p.b3[n];
        for (int r = rg; r < rows; r += GCN_THREADS / 64) {
            float a = bias;
#pragma unroll
            for (int k = 0; k < 64; k += 4) {
                const float4 x = *reinterpret_cast<const float4*>(e1 + r * 64 + k);
                a = fmaf(w[k], x.x, a); a = fmaf(w[k + 1], x.y, a); a = fmaf(w[k + 2], x.z, a); a = fmaf(w[k + 3], x.w, a);
            }
            e[r * 64 + n] = fmaxf(a, 0.0f);
        }
    }
    __syncthreads();
    // -- stage 3: theta = Ws1 e + bs1 (cols 0..127), phi = Ws2 e + bs2 (cols 128..255) --------------
    {
        const int n = tid;   // 256 output channels, one per thread
        const float* wrow = n < 128 ? p.ws1 + n * 64 : p.ws2 + (n - 128) * 64;
        float w[64];
#pragma unroll
        for (int k = 0; k < 64; k += 4) {
            const float4 t4 = *reinterpret_cast<const float4*>(wrow + k);
            w[k] = t4.x; w[k + 1] = t4.y; w[k + 2] = t4.z; w[k + 3] = t4.w;
        }
        const float bias = n < 128 ? p.bs1[n] : p.bs2[n - 128];
        for (int r = 0; r < rows; ++r) {
            float a = bias;
#pragma unroll
            for (int k = 0; k < 64; k += 4) {
                const float4 x = *reinterpret_cast<const float4*>(e + r * 64 + k);
                a = fmaf(w[k], x.x, a); a = fmaf(w[k + 1], x.y, a); a = fmaf(w[k + 2], x.z, a); a = fmaf(w[k + 3], x.w, a);
            }
            thph[r * GCN_LDT + n] = a;
        }
    }
    __syncthreads();
    // -- stage 4: logits S[i][j] = theta_i . phi_j (no 1/sqrt(d), models_gcn.py:97-99) --------------
    for (int idx = tid; idx < nf * V * V; idx += GCN_THREADS) {
        const int f = idx / (V * V);
        const int ij = idx - f * V * V;
        const int i = ij / V, j = ij - i * V;
        const float* th = thph + (f * V + i) * GCN_LDT;
        const float* ph = thph + (f * V + j) * GCN_LDT + 128;
        float a = 0.0f;
#pragma unroll 8
        for (int k = 0; k < 128; ++k) a = fmaf(th[k], ph[k], a);
        S[(f * V + i) * GCN_LDS + j] = a;
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
    // -- stage 6: Se = S e  (into the e1 buffer) ----------------------------------------------------
    {
        const int c = tid & 63, rg = tid >> 6;
        for (int r = rg; r < rows; r += GCN_THREADS / 64) {
            const int f = r / V;
            const float* srow = S + r * GCN_LDS;
            const float* eb = e + (f * V) * 64 + c;
            float a = 0.0f;
            for (int j = 0; j < V; ++j) a = fmaf(srow[j], eb[j * 64], a);
            e1[r * 64 + c] = a;
        }
    }
    __syncthreads();
    // -- stage 7: out = Se Wg (64 -> 128), staged in the theta|phi buffer ---------------------------
    {
        const int n = tid & 127, rg = tid >> 7;
        float w[64];
#pragma unroll
        for (int k = 0; k < 64; ++k) w[k] = p.wg[k * 128 + n];
        float* ostage = thph;
        for (int r = rg; r < rows; r += GCN_THREADS / 128) {
            float a = 0.0f;
#pragma unroll
            for (int k = 0; k < 64; k += 4) {
                const float4 x = *reinterpret_cast<const float4*>(e1 + r * 64 + k);
                a = fmaf(x.x, w[k], a); a = fmaf(x.y, w[k + 1], a); a = fmaf(x.z, w[k + 2], a); a = fmaf(x.w, w[k + 3], a);
            }
            ostage[r * GCN_LDO + n] = a;
        }
    }
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
    return sizeof(float) * (size_t)GCN_TT * V * (4 + 64 + 64 + GCN_LDT + GCN_LDS);
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
    static size_t configured = 0;
    if (smem > configured) {
        TG_CUDA_OK(cudaFuncSetAttribute(geo_gcn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    const int tiles = cdiv(T, GCN_TT);
    geo_gcn_kernel<<<B * tiles, GCN_THREADS, smem, stream>>>(p);
    TG_LAUNCH_OK();
    return 0;
}

}  // namespace tg
