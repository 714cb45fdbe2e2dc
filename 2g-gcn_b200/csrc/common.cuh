// Shared device/host helpers for the sm_100a kernels of the 2G-GCN hot path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <math.h>

#include "../../include/tggcn_b200.h"

namespace tg {

// ---------------------------------------------------------------------------------------------
// error plumbing (host)
// ---------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
}  // namespace tg
namespace tg {

#define TG_CUDA_OK(expr)                                                                        \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess) {                                                                \
            tg::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return 1;                                                                           \
        }                                                                                       \
    } while (0)

#define TG_REQUIRE(cond, ...)                \
    do {                                     \
        if (!(cond)) {                       \
            tg::set_error(__VA_ARGS__);      \
            return 2;                        \
        }                                    \
    } while (0)

extern unsigned long long g_launches;

bool debug_sync();   // TGGCN_DEBUG_SYNC=1: synchronise after every launch so an asynchronous fault names its kernel

#define TG_LAUNCH_OK()                                                                          \
    do {                                                                                        \
        ++tg::g_launches;                                                                       \
        cudaError_t _e = cudaGetLastError();                                                    \
        if (_e == cudaSuccess && tg::debug_sync()) _e = cudaDeviceSynchronize();                \
        if (_e != cudaSuccess) {                                                                \
            tg::set_error("%s:%d: launch failed -> %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
            return 1;                                                                           \
        }                                                                                       \
    } while (0)

inline int cdiv(int a, int b) { return (a + b - 1) / b; }
inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

int num_sms();   // SM count of the CURRENT device (cached per device)
// Raise a kernel's dynamic shared-memory limit to `bytes` on the CURRENT device if it is lower (function attributes are
// per device: the cache is keyed by (device, function)).
int ensure_smem(const void* func, size_t bytes);

// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Distance-based attention (compute_distance_based_attention_weights, vhoi/models.py:1757-1775).  dist = {hh (B,T,H,H), ho (B,T,H,O),
// oo (B,T,O,O)}, n = b*T + t, message kind 0 hh, 1 oh (receiver human, sender object), 2 ho (receiver object, sender human), 3 oo.
// Returns false when the kind keeps its dot-product attention (no distances given); else `logit` = 1 / (d + 1e-7) and `valid` =
// the distance is non-zero (zero distances are masked out like virtual senders).
__device__ __forceinline__ const float* dist_of_kind(const float* const* dist, int kind) {
    return kind == 0 ? dist[0] : (kind == 3 ? dist[2] : dist[1]);
}
__device__ __forceinline__ bool dist_logit(const float* const* dist, int kind, size_t n, int H, int O, int r, int s, float& logit,
                                           bool& valid) {
    const float* p = dist_of_kind(dist, kind);
    if (p == nullptr) return false;
    const float d = kind == 0 ? __ldg(p + (n * H + r) * H + s)
                  : kind == 1 ? __ldg(p + (n * H + r) * O + s)
                  : kind == 2 ? __ldg(p + (n * H + s) * O + r)
                              : __ldg(p + (n * O + r) * O + s);
    logit = 1.0f / (d + 1e-7f);
    valid = d != 0.0f;
    return true;
}

__device__ __forceinline__ float sigmoidf_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

// 16-byte async copy global -> shared, bypassing L1 (.cg): used for data other CTAs produced.
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// L2-coherent scalar / vector loads (never served from a stale L1 line).
__device__ __forceinline__ float ld_cg(const float* p) {
    float v;
    asm volatile("ld.global.cg.f32 %0, [%1];\n" : "=f"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ float4 ld_cg4(const float* p) {
    float4 v;
    asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];\n" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

// ---------------------------------------------------------------------------------------------
// grid-wide barrier for persistent (cooperatively launched) kernels.
// A monotonically increasing arrival counter: barrier number k (1-based) completes when the counter
// reaches k * nblocks.  Spins are bounded so a scheduling mistake surfaces as an error flag instead of
// a hung device.
// ---------------------------------------------------------------------------------------------
struct GridSync {
    unsigned int* counter;   // zeroed before launch
    unsigned int* error;     // set to 1 on spin timeout
};

// Split form: grid_arrive publishes this CTA's writes (release at GPU scope, cumulative over the CTA through the
// __syncthreads) and returns at once; work that does not depend on other CTAs can run before grid_wait, which blocks
// until every CTA has arrived and makes their writes visible to the whole CTA.
__device__ __forceinline__ void grid_arrive(const GridSync& gs, unsigned int& epoch) {
    __syncthreads();
    epoch += 1;
#ifdef GRID_FENCE_BARRIER
    if (threadIdx.x == 0) { __threadfence(); atomicAdd(gs.counter, 1u); }
#else
    if (threadIdx.x == 0) asm volatile("red.release.gpu.global.add.u32 [%0], %1;\n" ::"l"(gs.counter), "r"(1u) : "memory");
#endif
}

__device__ __forceinline__ bool grid_wait(const GridSync& gs, unsigned int epoch, unsigned int nblocks, volatile int* s_fail) {
    if (threadIdx.x == 0) {
        const unsigned int target = epoch * nblocks;
        unsigned int spins = 0;
        while (true) {
            unsigned int v;
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(gs.counter) : "memory");
            if (v >= target) break;
            if (++spins > (1u << 22)) {   // far beyond any legitimate wait (each probe is an L2 round trip)
                atomicOr(gs.error, 1u);   // bit 0: timeout (bit 1: operand range of the fp16-split tiles)
                *s_fail = 1;
                break;
            }
            if ((spins & 1023u) == 0 && (*((volatile unsigned int*)gs.error) & 1u)) { *s_fail = 1; break; }
        }
#ifdef GRID_FENCE_BARRIER
        __threadfence();
#endif
    }
    __syncthreads();
    return *s_fail == 0;
}

__device__ __forceinline__ bool grid_barrier(const GridSync& gs, unsigned int& epoch, unsigned int nblocks,
                                             volatile int* s_fail) {
    grid_arrive(gs, epoch);
    return grid_wait(gs, epoch, nblocks, s_fail);
}

}  // namespace tg
