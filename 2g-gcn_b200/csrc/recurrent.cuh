// Building block of the two recurrent kernels (frame-level BiGRU, segment-level gated GRUCell graph):
// a CTA-wide "gate tile"
//     acc[g][r] += sum_k W[g*gate_stride + unit][k] * X[r][k]        g < NG gates, r < RB rows
// with the weight slice and the activation rows streamed through shared memory in K-chunks by
// cp.async (.cg: activations were written by other CTAs of the same persistent kernel, L1 must be
// bypassed) and double-buffered.  Thread (j, rg) owns unit j and rows 2rg, 2rg+1 for all gates, so the
// GRU gate math that follows is thread-local.
#pragma once
#include "common.cuh"

namespace tg {

constexpr int REC_THREADS = 256;

template <int NRG>
struct TileGeom {
    static constexpr int RB = 2 * NRG;               // rows per tile
    static constexpr int J = REC_THREADS / NRG;      // units per tile
};

// shared memory floats needed by tile_accumulate<NG,NRG,KC,XROWS>
template <int NG, int NRG, int KC, int XROWS>
__host__ __device__ constexpr int tile_smem_floats() {
    return 2 * (NG * TileGeom<NRG>::J + XROWS) * (KC + 4);
}

struct NoHook {
    __device__ __forceinline__ void operator()(const float*, int) const {}
};

// xrows: shared-memory array of XROWS global row pointers (nullptr = all-zero row).
// rows [0, RB) feed the gate GEMM; rows [RB, XROWS) are only visible to the hook.
template <int NG, int NRG, int KC, int XROWS, typename Hook>
__device__ __forceinline__ void tile_accumulate(float (&acc)[NG][2], const float* __restrict__ W, int ldw,
                                                int gate_stride, int unit0, int unit_end, int row_end,
                                                const float* const* xrows, int K, float* smem, Hook hook) {
    constexpr int J = TileGeom<NRG>::J;
    constexpr int LD = KC + 4;
    constexpr int F4 = KC / 4;
    float* Ws = smem;                              // [2][NG*J][LD]
    float* Xs = smem + 2 * NG * J * LD;            // [2][XROWS][LD]
    const int tid = threadIdx.x;
    const int j = (tid & 15) + 16 * (tid / (16 * NRG));
    const int rg = (tid >> 4) % NRG;
    const int nchunks = K / KC;
    if (nchunks <= 0) return;

    auto issue = [&](int c, int buf) {
        for (int f = tid; f < NG * J * F4; f += REC_THREADS) {
            const int row = f / F4, q = f - row * F4;
            const int g = row / J, jj = row - g * J;
            const int unit = unit0 + jj;
            const int wrow = g * gate_stride + unit;
            float* dst = Ws + (buf * NG * J + row) * LD + q * 4;
            if (unit < unit_end && wrow < row_end) cp_async16(dst, W + (size_t)wrow * ldw + c * KC + q * 4);
            else *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        for (int f = tid; f < XROWS * F4; f += REC_THREADS) {
            const int row = f / F4, q = f - row * F4;
            const float* src = xrows[row];
            float* dst = Xs + (buf * XROWS + row) * LD + q * 4;
            if (src != nullptr) cp_async16(dst, src + c * KC + q * 4);
            else *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        cp_async_commit();
    };

    issue(0, 0);
    for (int c = 0; c < nchunks; ++c) {
        const int buf = c & 1;
        if (c + 1 < nchunks) { issue(c + 1, buf ^ 1); cp_async_wait<1>(); }
        else                 { cp_async_wait<0>(); }
        __syncthreads();
        const float* wb = Ws + (buf * NG * J + j) * LD;
        const float* xb = Xs + (buf * XROWS + 2 * rg) * LD;
#pragma unroll
        for (int q = 0; q < F4; ++q) {
            const float4 x0 = *reinterpret_cast<const float4*>(xb + q * 4);
            const float4 x1 = *reinterpret_cast<const float4*>(xb + LD + q * 4);
#pragma unroll
            for (int g = 0; g < NG; ++g) {
                const float4 w = *reinterpret_cast<const float4*>(wb + g * J * LD + q * 4);
                acc[g][0] = fmaf(w.x, x0.x, acc[g][0]); acc[g][0] = fmaf(w.y, x0.y, acc[g][0]);
                acc[g][0] = fmaf(w.z, x0.z, acc[g][0]); acc[g][0] = fmaf(w.w, x0.w, acc[g][0]);
                acc[g][1] = fmaf(w.x, x1.x, acc[g][1]); acc[g][1] = fmaf(w.y, x1.y, acc[g][1]);
                acc[g][1] = fmaf(w.z, x1.z, acc[g][1]); acc[g][1] = fmaf(w.w, x1.w, acc[g][1]);
            }
        }
        hook(Xs + buf * XROWS * LD, LD);
        __syncthreads();
    }
}

// GRU cell update, gate order (r, z, n) as torch.nn.GRU / GRUCell (vhoi/models.py:267,:294):
//   r = s(xr + hr), z = s(xz + hz), n = tanh(xn + r*hn), h' = n + z*(h - n)
__device__ __forceinline__ float gru_update(float xr, float xz, float xn, float hr, float hz, float hn, float hprev) {
    const float r = sigmoidf_acc(xr + hr);
    const float z = sigmoidf_acc(xz + hz);
    const float n = tanhf(xn + r * hn);
    return n + z * (hprev - n);
}

}  // namespace tg
