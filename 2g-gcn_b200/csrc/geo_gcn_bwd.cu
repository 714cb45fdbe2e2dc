// Backward of the geometry-level GCN (autograd of Geo_gcn.forward, pyrutils/torch/models_gcn.py:30-100).
// The BatchNorm input is model data, so only the affine parameters of the norm get a gradient:
// d gamma = sum dY * x_hat, d beta = sum dY (second pass, geo_bn_bwd_kernel).
// A CTA walks over a contiguous run of frames (one CTA per SM); per frame it recomputes the forward intermediates in shared memory, runs the
// backward stages, and keeps the weight-gradient partials of "its" weight rows in registers across frames, so the
// global atomics are issued once per CTA.
#include "backward.cuh"

namespace tg {

constexpr int GB_THREADS = 256;
constexpr int GB_LDT = 257, GB_LDO = 129, GB_LDS = 32;

__global__ void __launch_bounds__(GB_THREADS, 1) geo_gcn_bwd_kernel(const GcnBwdParams P, int fpc) {
    extern __shared__ __align__(16) float sm[];
    const int V = P.V, T = P.T, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float* xn = sm;                       // [V][4]
    float* e1 = xn + V * 4;               // [V][64]
    float* e = e1 + V * 64;               // [V][64]
    float* thph = e + V * 64;             // [V][257]  theta | phi
    float* S = thph + V * GB_LDT;         // [V][32]
    float* Se = S + V * GB_LDS;           // [V][64]
    float* dout = Se + V * 64;            // [V][129]
    float* dSe = dout + V * GB_LDO;       // [V][64]
    float* dS = dSe + V * 64;             // [V][32]   d S, then d logits
    float* de = dS + V * GB_LDS;          // [V][64]   d e, then d pre3
    float* dth = de + V * 64;             // [V][257]  d theta | d phi
    float* de1 = dth + V * GB_LDT;        // [V][64]   d e1, then d pre1

    // register partials (see the stage comments for the thread -> weight-row mapping)
    float g_wg[32], g_ws[64], g_w3[16], g_w1[4];
    float g_bs = 0.f, g_b3 = 0.f, g_b1 = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) g_wg[i] = 0.f;
#pragma unroll
    for (int i = 0; i < 64; ++i) g_ws[i] = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) g_w3[i] = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) g_w1[i] = 0.f;

    const int N = P.B * T;
    const int f0 = blockIdx.x * fpc, f1 = min(f0 + fpc, N);
    for (int n = f0; n < f1; ++n) {
        const int b = n / T, t = n - b * T;
        __syncthreads();
        // ---------------- forward recompute ----------------
        for (int idx = tid; idx < V * 4; idx += GB_THREADS) {
            const int v = idx >> 2, c = idx & 3;
            const float x = P.xh[((size_t)n * P.H) * P.Fh + 2048 + v * 4 + c];
            const int ch = c * V + v;
            xn[idx] = (x - P.mean[ch]) * (1.0f / sqrtf(P.var[ch] + 1e-5f)) * P.gamma[ch] + P.beta[ch];
        }
        for (int idx = tid; idx < V * 128; idx += GB_THREADS) {       // upstream gradient of this frame: (B,128,V,T)
            const int c = idx / V, v = idx - c * V;
            dout[v * GB_LDO + c] = P.dout[((size_t)(b * 128 + c) * V + v) * T + t];
        }
        __syncthreads();
        for (int idx = tid; idx < V * 64; idx += GB_THREADS) {
            const int r = idx >> 6, k = idx & 63;
            const float4 w = __ldg(reinterpret_cast<const float4*>(P.w1 + k * 4));
            const float4 x = *reinterpret_cast<const float4*>(xn + r * 4);
            e1[idx] = fmaxf(fmaf(w.x, x.x, fmaf(w.y, x.y, fmaf(w.z, x.z, fmaf(w.w, x.w, __ldg(P.b1 + k))))), 0.0f);
        }
        __syncthreads();
        for (int idx = tid; idx < V * 64; idx += GB_THREADS) {
            const int r = idx >> 6, k = idx & 63;
            float a = __ldg(P.b3 + k);
            for (int q = 0; q < 64; ++q) a = fmaf(__ldg(P.w3 + k * 64 + q), e1[r * 64 + q], a);
            e[idx] = fmaxf(a, 0.0f);
        }
        __syncthreads();
        {
            const int c = tid;     // 256 output channels
            const float* wrow = c < 128 ? P.ws1 + c * 64 : P.ws2 + (c - 128) * 64;
            const float bias = c < 128 ? __ldg(P.bs1 + c) : __ldg(P.bs2 + c - 128);
            for (int r = 0; r < V; ++r) {
                float a = bias;
                for (int q = 0; q < 64; ++q) a = fmaf(__ldg(wrow + q), e[r * 64 + q], a);
                thph[r * GB_LDT + c] = a;
            }
        }
        __syncthreads();
        for (int idx = tid; idx < V * V; idx += GB_THREADS) {
            const int i = idx / V, j = idx - i * V;
            float a = 0.0f;
            for (int q = 0; q < 128; ++q) a = fmaf(thph[i * GB_LDT + q], thph[j * GB_LDT + 128 + q], a);
            S[i * GB_LDS + j] = a;
        }
        __syncthreads();
        for (int r = warp; r < V; r += GB_THREADS / 32) {
            const float v = lane < V ? S[r * GB_LDS + lane] : -INFINITY;
            const float m = warp_max(v);
            const float ex = lane < V ? expf(v - m) : 0.0f;
            const float s = warp_sum(ex);
            if (lane < V) S[r * GB_LDS + lane] = ex / s;
        }
        __syncthreads();
        for (int idx = tid; idx < V * 64; idx += GB_THREADS) {
            const int r = idx >> 6, k = idx & 63;
            float a = 0.0f;
            for (int j = 0; j < V; ++j) a = fmaf(S[r * GB_LDS + j], e[j * 64 + k], a);
            Se[idx] = a;
        }
        __syncthreads();
        // ---------------- backward ----------------
        // out = Se Wg:  dSe[v][k] = sum_c dout[v][c] Wg[k][c];  dWg[k][c] += sum_v Se[v][k] dout[v][c]
        for (int idx = tid; idx < V * 64; idx += GB_THREADS) {
            const int r = idx >> 6, k = idx & 63;
            float a = 0.0f;
            for (int c = 0; c < 128; ++c) a = fmaf(dout[r * GB_LDO + c], __ldg(P.wg + k * 128 + c), a);
            dSe[idx] = a;
        }
        {
            const int c = tid & 127, kb = (tid >> 7) * 32;      // thread owns Wg[kb..kb+31][c]
            for (int v = 0; v < V; ++v) {
                const float d = dout[v * GB_LDO + c];
#pragma unroll
                for (int kk = 0; kk < 32; ++kk) g_wg[kk] = fmaf(Se[v * 64 + kb + kk], d, g_wg[kk]);
            }
        }
        __syncthreads();
        // Se = S e:  dS[i][j] = <dSe[i], e[j]>;  de[j] = sum_i S[i][j] dSe[i]
        for (int idx = tid; idx < V * V; idx += GB_THREADS) {
            const int i = idx / V, j = idx - i * V;
            float a = 0.0f;
            for (int k = 0; k < 64; ++k) a = fmaf(dSe[i * 64 + k], e[j * 64 + k], a);
            dS[i * GB_LDS + j] = a;
        }
        for (int idx = tid; idx < V * 64; idx += GB_THREADS) {
            const int j = idx >> 6, k = idx & 63;
            float a = 0.0f;
            for (int i = 0; i < V; ++i) a = fmaf(S[i * GB_LDS + j], dSe[i * 64 + k], a);
            de[idx] = a;
        }
        __syncthreads();
        // softmax backward -> d logits
        for (int r = warp; r < V; r += GB_THREADS / 32) {
            const float sv = lane < V ? S[r * GB_LDS + lane] : 0.0f;
            const float dv = lane < V ? dS[r * GB_LDS + lane] : 0.0f;
            const float dot = warp_sum(sv * dv);
            if (lane < V) dS[r * GB_LDS + lane] = sv * (dv - dot);
        }
        __syncthreads();
        // logits = theta phi^T:  dtheta[i] = sum_j dL[i][j] phi[j];  dphi[j] = sum_i dL[i][j] theta[i]
        {
            const int c = tid;
            if (c < 128) {
                for (int i = 0; i < V; ++i) {
                    float a = 0.0f;
                    for (int j = 0; j < V; ++j) a = fmaf(dS[i * GB_LDS + j], thph[j * GB_LDT + 128 + c], a);
                    dth[i * GB_LDT + c] = a;
                }
            } else {
                for (int j = 0; j < V; ++j) {
                    float a = 0.0f;
                    for (int i = 0; i < V; ++i) a = fmaf(dS[i * GB_LDS + j], thph[i * GB_LDT + (c - 128)], a);
                    dth[j * GB_LDT + c] = a;
                }
            }
        }
        __syncthreads();
        // theta|phi = Wcat e + b:  dWcat[c][k] += sum_v dth[v][c] e[v][k] (thread c owns row c);  de += dth Wcat
        {
            const int c = tid;
            for (int v = 0; v < V; ++v) {
                const float d = dth[v * GB_LDT + c];
                g_bs += d;
#pragma unroll
                for (int k = 0; k < 64; ++k) g_ws[k] = fmaf(d, e[v * 64 + k], g_ws[k]);
            }
        }
        {
            const int k = tid & 63, rg = tid >> 6;
            float acc[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = 0.0f;
            for (int c = 0; c < 256; ++c) {
                const float w = c < 128 ? __ldg(P.ws1 + c * 64 + k) : __ldg(P.ws2 + (c - 128) * 64 + k);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int v = rg + 4 * i;
                    if (v < V) acc[i] = fmaf(dth[v * GB_LDT + c], w, acc[i]);
                }
            }
            __syncthreads();      // all reads of de by nobody yet; writes below are to own elements only
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int v = rg + 4 * i;
                if (v < V) {
                    const float tot = de[v * 64 + k] + acc[i];
                    de[v * 64 + k] = e[v * 64 + k] > 0.0f ? tot : 0.0f;       // through the ReLU: d pre3
                }
            }
        }
        __syncthreads();
        // e = relu(W3 e1 + b3):  dW3[c][k] += sum_v dpre3[v][c] e1[v][k] (thread (c, rg) owns k in [16rg,16rg+16));  de1 = dpre3 W3
        {
            const int c = tid & 63, rg = tid >> 6;
            for (int v = 0; v < V; ++v) {
                const float d = de[v * 64 + c];
                if (rg == 0) g_b3 += d;
#pragma unroll
                for (int kk = 0; kk < 16; ++kk) g_w3[kk] = fmaf(d, e1[v * 64 + rg * 16 + kk], g_w3[kk]);
            }
            float acc[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = 0.0f;
            const int k = c;
            for (int q = 0; q < 64; ++q) {
                const float w = __ldg(P.w3 + q * 64 + k);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int v = rg + 4 * i;
                    if (v < V) acc[i] = fmaf(de[v * 64 + q], w, acc[i]);
                }
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int v = rg + 4 * i;
                if (v < V) de1[v * 64 + k] = e1[v * 64 + k] > 0.0f ? acc[i] : 0.0f;     // d pre1
            }
        }
        __syncthreads();
        // e1 = relu(W1 xn + b1):  dW1[c][q] += sum_v dpre1[v][c] xn[v][q];  d xn = dpre1 W1
        if (tid < 64) {
            for (int v = 0; v < V; ++v) {
                const float d = de1[v * 64 + tid];
                g_b1 += d;
#pragma unroll
                for (int q = 0; q < 4; ++q) g_w1[q] = fmaf(d, xn[v * 4 + q], g_w1[q]);
            }
        }
        for (int idx = tid; idx < V * 4; idx += GB_THREADS) {
            const int v = idx >> 2, q = idx & 3;
            float a = 0.0f;
            for (int c = 0; c < 64; ++c) a = fmaf(de1[v * 64 + c], __ldg(P.w1 + c * 4 + q), a);
            P.dxn[(size_t)n * V * 4 + idx] = a;
        }
    }
    // ---------------- flush the register partials ----------------
    {
        const int c = tid & 127, kb = (tid >> 7) * 32;
#pragma unroll
        for (int kk = 0; kk < 32; ++kk) atomicAdd(P.dwg + (kb + kk) * 128 + c, g_wg[kk]);
    }
    {
        const int c = tid;
        float* dw = c < 128 ? P.dws1 + c * 64 : P.dws2 + (c - 128) * 64;
#pragma unroll
        for (int k = 0; k < 64; ++k) atomicAdd(dw + k, g_ws[k]);
        atomicAdd(c < 128 ? P.dbs1 + c : P.dbs2 + (c - 128), g_bs);
    }
    {
        const int c = tid & 63, rg = tid >> 6;
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) atomicAdd(P.dw3 + c * 64 + rg * 16 + kk, g_w3[kk]);
        if (rg == 0) atomicAdd(P.db3 + c, g_b3);
    }
    if (tid < 64) {
#pragma unroll
        for (int q = 0; q < 4; ++q) atomicAdd(P.dw1 + tid * 4 + q, g_w1[q]);
        atomicAdd(P.db1 + tid, g_b1);
    }
}

// d gamma[ch] = sum_n dY[n][v][c] * x_hat, d beta[ch] = sum_n dY;  one CTA per node v (4 channels)
__global__ void __launch_bounds__(256) geo_bn_bwd_kernel(const float* __restrict__ xh, const float* __restrict__ dxn,
                                                        const float* __restrict__ mean, const float* __restrict__ var,
                                                        float* __restrict__ dgamma, float* __restrict__ dbeta, int N, int H, int V, int Fh) {
    const int v = blockIdx.x;
    float sg[4] = {0, 0, 0, 0}, sb[4] = {0, 0, 0, 0};
    float mu[4], rs[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) { mu[c] = mean[c * V + v]; rs[c] = 1.0f / sqrtf(var[c * V + v] + 1e-5f); }
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        const float4 x = *reinterpret_cast<const float4*>(xh + (size_t)n * H * Fh + 2048 + v * 4);
        const float4 d = *reinterpret_cast<const float4*>(dxn + ((size_t)n * V + v) * 4);
        sg[0] += d.x * (x.x - mu[0]) * rs[0]; sb[0] += d.x;
        sg[1] += d.y * (x.y - mu[1]) * rs[1]; sb[1] += d.y;
        sg[2] += d.z * (x.z - mu[2]) * rs[2]; sb[2] += d.z;
        sg[3] += d.w * (x.w - mu[3]) * rs[3]; sb[3] += d.w;
    }
    __shared__ float sh[8][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        sg[c] = warp_sum(sg[c]); sb[c] = warp_sum(sb[c]);
        if (lane == 0) { sh[warp][c] = sg[c]; sh[warp][4 + c] = sb[c]; }
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        float a = 0.f, bsum = 0.f;
        for (int w = 0; w < 8; ++w) { a += sh[w][threadIdx.x]; bsum += sh[w][4 + threadIdx.x]; }
        dgamma[threadIdx.x * V + v] = a;
        dbeta[threadIdx.x * V + v] = bsum;
    }
}

int launch_geo_gcn_bwd(const GcnBwdParams& P, cudaStream_t stream) {
    TG_REQUIRE(P.V >= 1 && P.V <= 32, "geo_gcn_bwd: gcn_node=%d unsupported", P.V);
    const size_t smem = sizeof(float) * (size_t)P.V * (4 + 64 + 64 + GB_LDT + GB_LDS + 64 + GB_LDO + 64 + GB_LDS + 64 + GB_LDT + 64);
    static size_t configured = 0;
    if (smem > configured) {
        TG_CUDA_OK(cudaFuncSetAttribute(geo_gcn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    const int fpc = cdiv(P.B * P.T, num_sms());          // frames per CTA: one wave, the register partials are flushed once per CTA
    geo_gcn_bwd_kernel<<<cdiv(P.B * P.T, fpc), GB_THREADS, smem, stream>>>(P, fpc);
    TG_LAUNCH_OK();
    return 0;
}

int launch_geo_bn_bwd(const float* xh, const float* dxn, const float* mean, const float* var, const float* gamma, float* dgamma,
                      float* dbeta, int B, int T, int H, int V, int Fh, cudaStream_t stream) {
    (void)gamma;
    geo_bn_bwd_kernel<<<V, 256, 0, stream>>>(xh, dxn, mean, var, dgamma, dbeta, B * T, H, V, Fh);
    TG_LAUNCH_OK();
    return 0;
}

}  // namespace tg
