// Backward of the geometry-level GCN (autograd of Geo_gcn.forward, pyrutils/torch/models_gcn.py:30-100).
// The BatchNorm input is model data, so only the affine parameters of the norm get a gradient:
// d gamma = sum dY * x_hat, d beta = sum dY (second pass, geo_bn_bwd_kernel).
// A CTA walks over a contiguous run of frames (one CTA per SM); per frame it recomputes the forward intermediates in shared memory, runs the
// backward stages, and keeps the weight-gradient partials of "its" weight rows in registers across frames, so the
// global atomics are issued once per CTA.
#include "backward.cuh"
#include "recurrent.cuh"

namespace tg {

constexpr int GB_THREADS = 256;
// row strides ≡ 4 (mod 32) words: the scalar fragment loads of an MMA (8 rows x 4 columns) hit 32 distinct banks
constexpr int GB_LD = 68, GB_LDT = 260, GB_LDO = 132, GB_LDS = 36;

// One warp: C[m-tile mt][n-tiles nt0 .. nt0+NT-1] += A (16 x 8*ksteps) * B (8*ksteps x 8*NT) on mma.sync.m16n8k8 with the 3xTF32
// split (gradients are too small for an fp16 split).  A(row, k) and Bm(k, col) are element accessors (they return 0 outside the
// matrices), so transposed and strided operands — weight gradients contract over the node index — need no staging.
// C fragment: c[j][0] = (16 mt + g8, 8 (nt0 + j) + 2 t4), [1] = column + 1, [2] / [3] = row + 8.
template <int NT, class FA, class FB>
__device__ __forceinline__ void warp_mma_tf32x3(float (&c)[NT][4], int mt, int nt0, int ksteps, FA A, FB Bm) {
    const int lane = threadIdx.x & 31, g8 = lane >> 2, t4 = lane & 3;
    const int r0 = mt * 16 + g8;
#pragma unroll 2
    for (int ks = 0; ks < ksteps; ++ks) {
        const int k0 = ks * 8 + t4;
        uint32_t ah[4], al[4];
        split_tf32(A(r0, k0), ah[0], al[0]);
        split_tf32(A(r0 + 8, k0), ah[1], al[1]);
        split_tf32(A(r0, k0 + 4), ah[2], al[2]);
        split_tf32(A(r0 + 8, k0 + 4), ah[3], al[3]);
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            const int n = (nt0 + j) * 8 + g8;
            uint32_t bh[2], bl[2];
            split_tf32(Bm(k0, n), bh[0], bl[0]);
            split_tf32(Bm(k0 + 4, n), bh[1], bl[1]);
            mma_tf32(c[j], al, bh);
            mma_tf32(c[j], ah, bl);
            mma_tf32(c[j], ah, bh);
        }
    }
}

template <int NT>
__device__ __forceinline__ void frag_zero(float (&c)[NT][4]) {
#pragma unroll
    for (int j = 0; j < NT; ++j) c[j][0] = c[j][1] = c[j][2] = c[j][3] = 0.0f;
}
// f(row, col, value) for every element of the fragments of m-tile mt, n-tiles nt0 .. nt0+NT-1
template <int NT, class F>
__device__ __forceinline__ void frag_each(const float (&c)[NT][4], int mt, int nt0, F f) {
    const int lane = threadIdx.x & 31, g8 = lane >> 2, t4 = lane & 3;
#pragma unroll
    for (int j = 0; j < NT; ++j) {
        const int r = mt * 16 + g8, col = (nt0 + j) * 8 + 2 * t4;
        f(r, col, c[j][0]); f(r, col + 1, c[j][1]); f(r + 8, col, c[j][2]); f(r + 8, col + 1, c[j][3]);
    }
}

__global__ void __launch_bounds__(GB_THREADS, 1) geo_gcn_bwd_kernel(const GcnBwdParams P, int fpc) {
    extern __shared__ __align__(16) float sm[];
    const int V = P.V, T = P.T, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float* xn = sm;                       // [V][4]
    float* e1 = xn + V * 4;               // [V][68]
    float* e = e1 + V * GB_LD;            // [V][68]
    float* thph = e + V * GB_LD;          // [V][260]  theta | phi
    float* S = thph + V * GB_LDT;         // [V][36]
    float* Se = S + V * GB_LDS;           // [V][68]
    float* dout = Se + V * GB_LD;         // [V][132]
    float* dSe = dout + V * GB_LDO;       // [V][68]
    float* dS = dSe + V * GB_LD;          // [V][36]   d S, then d logits
    float* de = dS + V * GB_LDS;          // [V][68]   d e, then d pre3
    float* dth = de + V * GB_LD;          // [V][260]  d theta | d phi
    float* de1 = dth + V * GB_LDT;        // [V][68]   d e1, then d pre1

    // weight-gradient partials: MMA accumulator fragments kept in registers across the CTA's frames
    //   dWg  (64 x 128)  : warp -> m-tile warp & 3, n-tiles 8 (warp >> 2) .. +7
    //   dWs  (256 x 64)  : warp -> m-tiles warp, warp + 8, all 8 n-tiles          (rows 0..127 = Ws1, 128..255 = Ws2)
    //   dW3  (64 x 64)   : warp -> m-tile warp & 3, n-tiles 4 (warp >> 2) .. +3
    float g_wg[8][4], g_ws[2][8][4], g_w3[4][4], g_w1[4];
    float g_bs = 0.f, g_b3 = 0.f, g_b1 = 0.f;
    frag_zero(g_wg); frag_zero(g_ws[0]); frag_zero(g_ws[1]); frag_zero(g_w3);
#pragma unroll
    for (int i = 0; i < 4; ++i) g_w1[i] = 0.f;
    const int kv = (V + 7) / 8;           // k8 steps of a contraction over the nodes

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
            e1[r * GB_LD + k] = fmaxf(fmaf(w.x, x.x, fmaf(w.y, x.y, fmaf(w.z, x.z, fmaf(w.w, x.w, __ldg(P.b1 + k))))), 0.0f);
        }
        __syncthreads();
        {   // e = relu(e1 W3^T + b3): 2 x 8 tiles, warp -> (m-tile warp & 1, n-tiles 2 (warp >> 1), +1)
            float c[2][4];
            frag_zero(c);
            const int mt = warp & 1, nt0 = (warp >> 1) * 2;
            warp_mma_tf32x3<2>(c, mt, nt0, 8, [&](int r, int k) { return r < V ? e1[r * GB_LD + k] : 0.0f; },
                               [&](int k, int nn) { return __ldg(P.w3 + nn * 64 + k); });
            frag_each<2>(c, mt, nt0, [&](int r, int col, float v) { if (r < V) e[r * GB_LD + col] = fmaxf(v + __ldg(P.b3 + col), 0.0f); });
        }
        __syncthreads();
        {   // theta | phi = e Wcat^T + b: 2 x 32 tiles, warp -> (m-tile warp & 1, n-tiles 8 (warp >> 1) .. +7)
            float c[8][4];
            frag_zero(c);
            const int mt = warp & 1, nt0 = (warp >> 1) * 8;
            warp_mma_tf32x3<8>(c, mt, nt0, 8, [&](int r, int k) { return r < V ? e[r * GB_LD + k] : 0.0f; },
                               [&](int k, int nn) { return nn < 128 ? __ldg(P.ws1 + nn * 64 + k) : __ldg(P.ws2 + (nn - 128) * 64 + k); });
            frag_each<8>(c, mt, nt0, [&](int r, int col, float v) {
                if (r < V) thph[r * GB_LDT + col] = v + (col < 128 ? __ldg(P.bs1 + col) : __ldg(P.bs2 + col - 128));
            });
        }
        __syncthreads();
        {   // logits S = theta phi^T: 2 x 4 tiles, one per warp
            float c[1][4];
            frag_zero(c);
            const int mt = warp & 1, nt0 = warp >> 1;
            warp_mma_tf32x3<1>(c, mt, nt0, 16, [&](int r, int k) { return r < V ? thph[r * GB_LDT + k] : 0.0f; },
                               [&](int k, int j) { return j < V ? thph[j * GB_LDT + 128 + k] : 0.0f; });
            frag_each<1>(c, mt, nt0, [&](int r, int col, float v) { if (r < V && col < V) S[r * GB_LDS + col] = v; });
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
            for (int j = 0; j < V; ++j) a = fmaf(S[r * GB_LDS + j], e[j * GB_LD + k], a);
            Se[r * GB_LD + k] = a;
        }
        // ---------------- backward ----------------
        {   // out = Se Wg:  dSe[v][k] = sum_c dout[v][c] Wg[k][c]   (2 x 8 tiles, K = 128)
            float c[2][4];
            frag_zero(c);
            const int mt = warp & 1, nt0 = (warp >> 1) * 2;
            warp_mma_tf32x3<2>(c, mt, nt0, 16, [&](int r, int k) { return r < V ? dout[r * GB_LDO + k] : 0.0f; },
                               [&](int k, int nn) { return __ldg(P.wg + nn * 128 + k); });
            frag_each<2>(c, mt, nt0, [&](int r, int col, float v) { if (r < V) dSe[r * GB_LD + col] = v; });
        }
        __syncthreads();       // Se and dSe complete
        //   dWg[k][c] += sum_v Se[v][k] dout[v][c]   (contraction over the nodes)
        warp_mma_tf32x3<8>(g_wg, warp & 3, (warp >> 2) * 8, kv, [&](int k, int v) { return v < V ? Se[v * GB_LD + k] : 0.0f; },
                           [&](int v, int c) { return v < V ? dout[v * GB_LDO + c] : 0.0f; });
        {   // Se = S e:  dS[i][j] = <dSe[i], e[j]>   (2 x 4 tiles, K = 64)
            float c[1][4];
            frag_zero(c);
            const int mt = warp & 1, nt0 = warp >> 1;
            warp_mma_tf32x3<1>(c, mt, nt0, 8, [&](int r, int k) { return r < V ? dSe[r * GB_LD + k] : 0.0f; },
                               [&](int k, int j) { return j < V ? e[j * GB_LD + k] : 0.0f; });
            frag_each<1>(c, mt, nt0, [&](int r, int col, float v) { if (r < V && col < V) dS[r * GB_LDS + col] = v; });
        }
        //   de[j] = sum_i S[i][j] dSe[i]
        for (int idx = tid; idx < V * 64; idx += GB_THREADS) {
            const int j = idx >> 6, k = idx & 63;
            float a = 0.0f;
            for (int i = 0; i < V; ++i) a = fmaf(S[i * GB_LDS + j], dSe[i * GB_LD + k], a);
            de[j * GB_LD + k] = a;
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
        {   // logits = theta phi^T:  dtheta[i] = sum_j dL[i][j] phi[j];  dphi[j] = sum_i dL[i][j] theta[i]
            // 2 x 32 tiles over the 256 columns of dth, warp -> (m-tile warp & 1, n-tiles 8 (warp >> 1) .. +7): warps 0-3 dtheta, 4-7 dphi
            float c[8][4];
            frag_zero(c);
            const int mt = warp & 1, nt0 = (warp >> 1) * 8;
            if (nt0 < 16)
                warp_mma_tf32x3<8>(c, mt, nt0, kv, [&](int i, int j) { return (i < V && j < V) ? dS[i * GB_LDS + j] : 0.0f; },
                                   [&](int j, int col) { return j < V ? thph[j * GB_LDT + 128 + col] : 0.0f; });
            else
                warp_mma_tf32x3<8>(c, mt, nt0, kv, [&](int j, int i) { return (i < V && j < V) ? dS[i * GB_LDS + j] : 0.0f; },
                                   [&](int i, int col) { return i < V ? thph[i * GB_LDT + (col - 128)] : 0.0f; });
            frag_each<8>(c, mt, nt0, [&](int r, int col, float v) { if (r < V) dth[r * GB_LDT + col] = v; });
        }
        __syncthreads();
        // theta|phi = Wcat e + b:  dWcat[c][k] += sum_v dth[v][c] e[v][k];  d b[c] += sum_v dth[v][c];  de += dth Wcat
        {
            const auto A = [&](int c, int v) { return v < V ? dth[v * GB_LDT + c] : 0.0f; };
            const auto Bm = [&](int v, int k) { return v < V ? e[v * GB_LD + k] : 0.0f; };
            warp_mma_tf32x3<8>(g_ws[0], warp, 0, kv, A, Bm);
            warp_mma_tf32x3<8>(g_ws[1], warp + 8, 0, kv, A, Bm);
            for (int v = 0; v < V; ++v) g_bs += dth[v * GB_LDT + tid];
        }
        {   // de += dth Wcat   (2 x 8 tiles, K = 256), then through the ReLU: d pre3
            float c[2][4];
            frag_zero(c);
            const int mt = warp & 1, nt0 = (warp >> 1) * 2;
            warp_mma_tf32x3<2>(c, mt, nt0, 32, [&](int r, int k) { return r < V ? dth[r * GB_LDT + k] : 0.0f; },
                               [&](int k, int col) { return k < 128 ? __ldg(P.ws1 + k * 64 + col) : __ldg(P.ws2 + (k - 128) * 64 + col); });
            frag_each<2>(c, mt, nt0, [&](int r, int col, float v) {
                if (r < V) de[r * GB_LD + col] = e[r * GB_LD + col] > 0.0f ? de[r * GB_LD + col] + v : 0.0f;   // own elements only
            });
        }
        __syncthreads();
        // e = relu(W3 e1 + b3):  dW3[c][k] += sum_v dpre3[v][c] e1[v][k];  d b3[c] += sum_v dpre3[v][c];  de1 = dpre3 W3
        warp_mma_tf32x3<4>(g_w3, warp & 3, (warp >> 2) * 4, kv, [&](int c, int v) { return v < V ? de[v * GB_LD + c] : 0.0f; },
                           [&](int v, int k) { return v < V ? e1[v * GB_LD + k] : 0.0f; });
        if (tid < 64)
            for (int v = 0; v < V; ++v) g_b3 += de[v * GB_LD + tid];
        {   // de1 = dpre3 W3 (2 x 8 tiles, K = 64), through the ReLU: d pre1
            float c[2][4];
            frag_zero(c);
            const int mt = warp & 1, nt0 = (warp >> 1) * 2;
            warp_mma_tf32x3<2>(c, mt, nt0, 8, [&](int r, int q) { return r < V ? de[r * GB_LD + q] : 0.0f; },
                               [&](int q, int k) { return __ldg(P.w3 + q * 64 + k); });
            frag_each<2>(c, mt, nt0, [&](int r, int col, float v) { if (r < V) de1[r * GB_LD + col] = e1[r * GB_LD + col] > 0.0f ? v : 0.0f; });
        }
        __syncthreads();
        // e1 = relu(W1 xn + b1):  dW1[c][q] += sum_v dpre1[v][c] xn[v][q];  d xn = dpre1 W1
        if (tid < 64) {
            for (int v = 0; v < V; ++v) {
                const float d = de1[v * GB_LD + tid];
                g_b1 += d;
#pragma unroll
                for (int q = 0; q < 4; ++q) g_w1[q] = fmaf(d, xn[v * 4 + q], g_w1[q]);
            }
        }
        for (int idx = tid; idx < V * 4; idx += GB_THREADS) {
            const int v = idx >> 2, q = idx & 3;
            float a = 0.0f;
            for (int c = 0; c < 64; ++c) a = fmaf(de1[v * GB_LD + c], __ldg(P.w1 + c * 4 + q), a);
            P.dxn[(size_t)n * V * 4 + idx] = a;
        }
    }
    // ---------------- flush the register partials ----------------
    frag_each<8>(g_wg, warp & 3, (warp >> 2) * 8, [&](int k, int c, float v) { atomicAdd(P.dwg + k * 128 + c, v); });
#pragma unroll
    for (int h = 0; h < 2; ++h)
        frag_each<8>(g_ws[h], warp + 8 * h, 0, [&](int c, int k, float v) {
            atomicAdd((c < 128 ? P.dws1 + c * 64 : P.dws2 + (c - 128) * 64) + k, v);
        });
    atomicAdd(tid < 128 ? P.dbs1 + tid : P.dbs2 + (tid - 128), g_bs);
    frag_each<4>(g_w3, warp & 3, (warp >> 2) * 4, [&](int c, int k, float v) { atomicAdd(P.dw3 + c * 64 + k, v); });
    if (tid < 64) {
        atomicAdd(P.db3 + tid, g_b3);
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
    const size_t smem = sizeof(float) * (size_t)P.V * (4 + GB_LD + GB_LD + GB_LDT + GB_LDS + GB_LD + GB_LDO + GB_LD + GB_LDS + GB_LD + GB_LDT + GB_LD);
    if (int rc = ensure_smem((const void*)geo_gcn_bwd_kernel, smem)) return rc;
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
