// Adam over the flat parameter / gradient buffers of the training driver (2g-gcn_b200/optim.py: FlatAdam) — the update of
// torch.optim.Adam (the optimiser train.py:40 builds; no amsgrad, L2 weight decay folded into the gradient) as ONE launch over all
// parameters instead of torch's ~20 multi-tensor launches: p, g, m, v are read once and p, m, v written once (28 bytes per
// parameter: HBM-bound).
#include "common.cuh"

namespace tg {

namespace {

__global__ void __launch_bounds__(256) adam_step_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                        float* __restrict__ v, size_t n4, size_t n, float lr_over_bc1, float beta1,
                                                        float beta2, float eps, float weight_decay, float inv_sqrt_bc2) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    auto upd = [&](float& pp, float gg, float& mm, float& vv) {
        if (weight_decay != 0.0f) gg = fmaf(weight_decay, pp, gg);
        mm = fmaf(1.0f - beta1, gg - mm, mm);                       // exp_avg.lerp_(grad, 1 - beta1)
        vv = fmaf(1.0f - beta2, gg * gg, beta2 * vv);               // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
        const float denom = sqrtf(vv) * inv_sqrt_bc2 + eps;
        pp -= lr_over_bc1 * (mm / denom);
    };
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 pp = reinterpret_cast<float4*>(p)[i], mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
        const float4 gg = __ldg(reinterpret_cast<const float4*>(g) + i);
        upd(pp.x, gg.x, mm.x, vv.x); upd(pp.y, gg.y, mm.y, vv.y); upd(pp.z, gg.z, mm.z, vv.z); upd(pp.w, gg.w, mm.w, vv.w);
        reinterpret_cast<float4*>(p)[i] = pp; reinterpret_cast<float4*>(m)[i] = mm; reinterpret_cast<float4*>(v)[i] = vv;
    }
    for (size_t i = n4 * 4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) upd(p[i], g[i], m[i], v[i]);
}

}  // namespace

}  // namespace tg

using namespace tg;

extern "C" {

int tggcn_adam_step(float* p, const float* g, float* m, float* v, size_t n, float lr, float beta1, float beta2, float eps,
                    float weight_decay, int step, void* stream) {
    TG_REQUIRE(p && g && m && v && n > 0 && step >= 1, "adam_step: null buffer or step < 1");
    TG_REQUIRE(((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                 reinterpret_cast<uintptr_t>(v)) & 15) == 0, "adam_step: buffers must be 16-byte aligned");
    const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
    const size_t n4 = n / 4;
    const int blocks = (int)(((n4 + 255) / 256) < (size_t)(num_sms() * 8) ? ((n4 + 255) / 256) : (size_t)(num_sms() * 8));
    adam_step_kernel<<<blocks > 0 ? blocks : 1, 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n4, n, (float)((double)lr / bc1), beta1, beta2, eps,
                                                                               weight_decay, (float)(1.0 / sqrt(bc2)));
    TG_LAUNCH_OK();
    return 0;
}

}  // extern "C"
