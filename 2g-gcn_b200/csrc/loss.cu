// Fused criterion of the 2G-GCN model (K-G, loss half): multi_task_loss (pyrutils/torch/losses.py:39-51) with the function tuple
// of vhoi/losses.py:41-60 — budget_loss (:24-36), binary_cross_entropy_loss (:7-21, positive_class_weight == 1) and
// F.nll_loss(ignore_index=-1, reduction='mean') — for all 6 (12) outputs of the model at once.
//   forward : one pass over every output: per-term sum of the element losses and count of valid targets (target != -1),
//             then  loss_i = w_i * sum_i / n_i  (0 when nothing is valid, like the reference's early return)
//   backward: d out_i = g_i * w_i / n_i * d(element loss)   written densely in the layout of the output
// No host synchronisation (the reference's two `.item()` calls per loss are gone); the loss values stay on the device.
#include "common.cuh"

namespace tg {

namespace {

constexpr int LOSS_MAX_TERMS = 16;
struct LossTerms {
    tggcn_loss_term t[LOSS_MAX_TERMS];
    int n;
};

__device__ __forceinline__ float bce_elem(float o, float t) {
    // F.binary_cross_entropy clamps each log term at -100
    return -(t * fmaxf(logf(o), -100.0f) + (1.0f - t) * fmaxf(logf(1.0f - o), -100.0f));
}

__global__ void __launch_bounds__(256) loss_sum_kernel(const LossTerms L, float* __restrict__ acc) {
    const tggcn_loss_term& t = L.t[blockIdx.y];
    float s = 0.0f, n = 0.0f;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < t.numel; i += stride) {
        if (t.kind == TGGCN_LOSS_NLL) {
            const long long tg_ = ((const long long*)t.target)[i];          // position (b, t, e)
            if (tg_ >= t.C) {             // F.nll_loss raises for a class index >= C: no out-of-bounds read here, the loss turns NaN
                s = __int_as_float(0x7fc00000);
                n += 1.0f;
            } else if (tg_ >= 0) {
                const long long e = i % t.E, bt = i / t.E, tt = bt % t.T, b = bt / t.T;
                s -= t.out[((b * t.C + tg_) * t.T + tt) * t.E + e];
                n += 1.0f;
            }
        } else {
            const float tv = ((const float*)t.target)[i];
            if (tv != -1.0f) {
                const float o = t.out[i];
                s += t.kind == TGGCN_LOSS_BCE ? bce_elem(o, tv) : o;
                n += 1.0f;
            }
        }
    }
    s = warp_sum(s); n = warp_sum(n);
    __shared__ float ss[8], sn[8];
    if ((threadIdx.x & 31) == 0) { ss[threadIdx.x >> 5] = s; sn[threadIdx.x >> 5] = n; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.f, b = 0.f;
        for (int w = 0; w < 8; ++w) { a += ss[w]; b += sn[w]; }
        atomicAdd(acc + 2 * blockIdx.y, a);
        atomicAdd(acc + 2 * blockIdx.y + 1, b);
    }
}

__global__ void loss_finalize_kernel(const LossTerms L, const float* __restrict__ acc, float* __restrict__ losses) {
    const int i = threadIdx.x;
    if (i >= L.n) return;
    const float n = acc[2 * i + 1];
    losses[i] = n > 0.0f ? L.t[i].weight * acc[2 * i] / n : 0.0f;
}

__global__ void __launch_bounds__(256) loss_grad_kernel(const LossTerms L, const float* __restrict__ acc, const float* __restrict__ gup) {
    const tggcn_loss_term& t = L.t[blockIdx.y];
    if (t.d_out == nullptr) return;
    const float n = acc[2 * blockIdx.y + 1];
    const float scale = n > 0.0f ? (gup != nullptr ? gup[blockIdx.y] : 1.0f) * t.weight / n : 0.0f;
    const long long total = t.kind == TGGCN_LOSS_NLL ? t.numel * t.C : t.numel;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        float d = 0.0f;
        if (t.kind == TGGCN_LOSS_NLL) {
            // i indexes (b, c, t, e)
            const long long e = i % t.E, r1 = i / t.E, tt = r1 % t.T, r2 = r1 / t.T, c = r2 % t.C, b = r2 / t.C;
            const long long tg_ = ((const long long*)t.target)[(b * t.T + tt) * t.E + e];
            d = tg_ == c ? -scale : 0.0f;
        } else {
            const float tv = ((const float*)t.target)[i];
            if (tv != -1.0f) {
                if (t.kind == TGGCN_LOSS_BCE) {
                    const float o = t.out[i];
                    d = scale * (o - tv) / fmaxf((1.0f - o) * o, 1e-12f);     // binary_cross_entropy_backward
                } else {
                    d = scale;
                }
            }
        }
        t.d_out[i] = d;
    }
}

int pack(const tggcn_loss_term* terms, int n_terms, LossTerms& L) {
    TG_REQUIRE(terms != nullptr && n_terms >= 1 && n_terms <= LOSS_MAX_TERMS, "loss: 1..%d terms expected", LOSS_MAX_TERMS);
    L.n = n_terms;
    for (int i = 0; i < n_terms; ++i) {
        L.t[i] = terms[i];
        TG_REQUIRE(terms[i].out && terms[i].target && terms[i].numel > 0, "loss: term %d has null / empty tensors", i);
        TG_REQUIRE(terms[i].kind >= TGGCN_LOSS_BUDGET && terms[i].kind <= TGGCN_LOSS_NLL, "loss: term %d has unknown kind", i);
        if (terms[i].kind == TGGCN_LOSS_NLL)
            TG_REQUIRE((long long)terms[i].B * terms[i].T * terms[i].E == terms[i].numel && terms[i].C >= 1, "loss: term %d: bad NLL shape", i);
    }
    return 0;
}

}  // namespace
}  // namespace tg

using namespace tg;

extern "C" {

int tggcn_loss_fwd(const tggcn_loss_term* terms, int n_terms, float* losses, float* scratch, void* stream_) {
    TG_REQUIRE(losses && scratch, "loss_fwd: null output");
    LossTerms L;
    if (int rc = pack(terms, n_terms, L)) return rc;
    cudaStream_t stream = (cudaStream_t)stream_;
    TG_CUDA_OK(cudaMemsetAsync(scratch, 0, sizeof(float) * 2 * n_terms, stream));
    loss_sum_kernel<<<dim3(32, n_terms), 256, 0, stream>>>(L, scratch);
    TG_LAUNCH_OK();
    loss_finalize_kernel<<<1, 32, 0, stream>>>(L, scratch, losses);
    TG_LAUNCH_OK();
    return 0;
}

int tggcn_loss_bwd(const tggcn_loss_term* terms, int n_terms, const float* scratch, const float* grad_losses, void* stream_) {
    TG_REQUIRE(scratch, "loss_bwd: null scratch");
    LossTerms L;
    if (int rc = pack(terms, n_terms, L)) return rc;
    loss_grad_kernel<<<dim3(64, n_terms), 256, 0, (cudaStream_t)stream_>>>(L, scratch, grad_losses);
    TG_LAUNCH_OK();
    return 0;
}

}  // extern "C"
