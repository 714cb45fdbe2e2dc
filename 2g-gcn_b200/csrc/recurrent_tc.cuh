// tcgen05 version of the recurrent "gate tile" (recurrent.cuh): same contract as tile_accumulate —
//     out[g][r][u] = sum_k Wrow[g*16 + u][k] * X[r][k]        g < NG groups (NG*16 <= 128 weight rows), r < 8*NT rows
// over up to two K segments addressed through pointer tables — but the products run on the 5th-generation tensor
// cores with the accumulator in TMEM:
//   * the WEIGHT rows are the UMMA M operand (M = 128; rows >= NG*16 are never loaded, their accumulator lanes are
//     ignored), the ACTIVATION rows the N operand (N = 16 or 32): tcgen05.mma cta_group::1 kind::tf32, K = 8 per
//     instruction, so one step of a recurrence costs 128*N/256 cycles per 8 columns of K instead of dozens of
//     mma.sync issue slots;
//   * fp32 accuracy through the same 3xTF32 split as everywhere else (a_lo*b_hi + a_hi*b_lo + a_hi*b_hi, fp32
//     accumulate): the 8 producer warps load the fp32 rows from L2 (ld.global.cg: activations were written by other CTAs
//     of the same persistent kernel), split them in registers and store hi / lo tiles in the canonical K-major
//     SWIZZLE_128B layout; a 4-stage mbarrier ring hands k-blocks of 32 floats to the single MMA-issuing thread
//     (warp 8), tcgen05.commit returns the stages;
//   * after the last k-block warps 0-3 read the accumulator with tcgen05.ld (lane = weight row), transpose it through
//     shared memory, and every thread receives the sums of "its" (unit, row) pairs exactly like tile_accumulate
//     delivers them, so the callers' epilogues (gate math, blends, saves) are unchanged.
// The ring's mbarrier phases and the TMEM allocation persist across tiles and steps of the persistent kernel (RtcState).
// K1 and K2 must be multiples of 32 floats.  Block = 288 threads (8 producer/epilogue warps + 1 MMA warp).
#pragma once
#include <stdlib.h>
#include "recurrent.cuh"
#include "tcgen05.cuh"

namespace tg {

constexpr int RTC_THREADS = REC_THREADS + 32;
constexpr int RTC_STAGES = 4;
constexpr int RTC_PF = 3;                                    // k-blocks each producer thread keeps in flight in registers
constexpr int RTC_BK = 32;                                   // floats per k-block: one 128-byte swizzle row
constexpr int RTC_A_BYTES = 128 * RTC_BK * 4;                // 16 KB: weight tile (hi or lo)
constexpr int RTC_B_BYTES = 32 * RTC_BK * 4;                 // 4 KB: activation tile (hi or lo)
constexpr int RTC_STAGE_BYTES = 2 * RTC_A_BYTES + 2 * RTC_B_BYTES;
constexpr int RTC_OUT_LD = 33;
constexpr int RTC_OUT_BYTES = 2 * 128 * RTC_OUT_LD * 4;     // two partial-sum buffers (accumulators 0-5 / 6-11)
constexpr int RTC_SMEM_BYTES = RTC_STAGES * RTC_STAGE_BYTES + RTC_OUT_BYTES + 1024;     // + alignment slack
constexpr uint32_t RTC_TMEM_COLS = 512;        // 12 accumulators of 32 columns (power-of-two allocation)
constexpr int RTC_ACCS = 12;

// Host: whether the recurrent kernels use the tcgen05 gate tile for this hidden size (every K segment is a multiple of D).
// Measured on B200: with only 16-32 activation rows per step the SS-mode tcgen05.mma is bound by the
// shared-memory read of its 128 x 8 weight operand (~75 cycles per instruction, 36 per k-block with the 3xTF32 split), so
// this tile is SLOWER than the mma.sync tile (6.9 vs 4.1 ms per segment pass) and is opt-in: TGGCN_REC_TC=1.
inline bool rec_use_tc(int D) {
    static int env = -1;
    if (env < 0) {
        const char* e = getenv("TGGCN_REC_TC");
        env = (e != nullptr && e[0] == '1') ? 1 : 0;
    }
    return env == 1 && D % RTC_BK == 0;
}

struct RtcShared {
    uint64_t bars[2 * RTC_STAGES + 1];     // full[STAGES], empty[STAGES], accumulator-complete
    uint32_t tmem_base;
};

struct RtcState {            // identical in every thread of the CTA
    uint32_t tmem_base;
    uint32_t kb_count;       // k-blocks pushed through the ring so far
    uint32_t tiles;          // tiles completed so far
    uint8_t* ring;           // 1024-byte aligned stage buffers
    float* outbuf;           // [2][128][RTC_OUT_LD] accumulator staging
    int dbg;                 // timing experiments: 1 = producers skip loads/stores, 2 = issuer skips the MMAs
};

// Once per kernel, by all RTC_THREADS threads.
__device__ __forceinline__ void rtc_init(RtcShared& sh, RtcState& st, uint8_t* smem_raw) {
    const int tid = threadIdx.x, warp = tid >> 5;
    st.ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    st.outbuf = reinterpret_cast<float*>(st.ring + RTC_STAGES * RTC_STAGE_BYTES);
    st.kb_count = 0;
    st.tiles = 0;
    st.dbg = 0;
    if (tid == 0) {
        for (int s = 0; s < RTC_STAGES; ++s) {
            mbar_init(smem_u32(&sh.bars[s]), REC_WARPS);
            mbar_init(smem_u32(&sh.bars[RTC_STAGES + s]), 1);
        }
        mbar_init(smem_u32(&sh.bars[2 * RTC_STAGES]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == REC_WARPS) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&sh.tmem_base)), "r"(RTC_TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    st.tmem_base = sh.tmem_base;
}

// Once per kernel at the end, by all threads (every exit path of the kernel must reach it).
__device__ __forceinline__ void rtc_finish(RtcState& st) {
    tc_fence_before();
    __syncthreads();
    if ((threadIdx.x >> 5) == REC_WARPS) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(st.tmem_base), "r"(RTC_TMEM_COLS) : "memory");
    }
}

// Pointer tables as in tile_accumulate: rows [0, NG*16) weight rows, rows [NG*16, NG*16 + 8*NT) activation rows; tab1 covers
// K columns [0, K1), tab2 columns [K1, K1 + K2) (pointers to the START of the segment); null = all-zero row.
// Must be called by all RTC_THREADS threads; the tables must be visible (a __syncthreads after filling them).
template <int NG, int NT>
__device__ __forceinline__ void tile_accumulate_tc(float (&out)[NG][(NT + 1) / 2], const float* const* tab1, const float* const* tab2,
                                                   int K1, int K2, RtcShared& sh, RtcState& st) {
    constexpr int WR = NG * REC_J, NB = 8 * NT, NA = (WR + 31) / 32;
    static_assert(WR <= 128 && (NB == 16 || NB == 32), "tile shape not supported by the tcgen05 gate tile");
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nkb1 = K1 / RTC_BK, nkb = (K1 + K2) / RTC_BK;
#pragma unroll
    for (int g = 0; g < NG; ++g)
#pragma unroll
        for (int p = 0; p < (NT + 1) / 2; ++p) out[g][p] = 0.0f;
    if (nkb == 0) return;                                       // uniform over the CTA
    const uint32_t full0 = smem_u32(&sh.bars[0]), empty0 = smem_u32(&sh.bars[RTC_STAGES]), tfull = smem_u32(&sh.bars[2 * RTC_STAGES]);
    const uint32_t ring_u32 = smem_u32(st.ring);
    const uint32_t kb_base = st.kb_count;

    if (warp < REC_WARPS) {
        // ------------------------------ producers ------------------------------
        const int c = tid & 7, r0 = tid >> 3;                   // 16-byte chunk c of rows r0 + 32*i
        const float* a1[NA];
        const float* a2[NA];
#pragma unroll
        for (int i = 0; i < NA; ++i) {
            const int r = r0 + 32 * i;
            const float* p1 = r < WR ? tab1[r] : nullptr;
            const float* p2 = (r < WR && K2 > 0) ? tab2[r] : nullptr;
            a1[i] = p1 != nullptr ? p1 + c * 4 : nullptr;
            a2[i] = p2 != nullptr ? p2 + c * 4 : nullptr;
        }
        const float* b1 = nullptr;
        const float* b2 = nullptr;
        if (r0 < NB) {
            const float* p1 = tab1[WR + r0];
            const float* p2 = K2 > 0 ? tab2[WR + r0] : nullptr;
            b1 = p1 != nullptr ? p1 + c * 4 : nullptr;
            b2 = p2 != nullptr ? p2 + c * 4 : nullptr;
        }
        const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
        float4 pa[RTC_PF][NA], pb[RTC_PF];                      // register prefetch, RTC_PF k-blocks deep (static slots)
        auto load = [&](int kb, float4 (&xa)[NA], float4& xb) {
            const bool s1 = kb < nkb1;
            const int off = (s1 ? kb : kb - nkb1) * RTC_BK;
#pragma unroll
            for (int i = 0; i < NA; ++i) {
                const float* p = s1 ? a1[i] : a2[i];
                xa[i] = p != nullptr ? ld_cg4(p + off) : zero4;
            }
            const float* q = s1 ? b1 : b2;
            xb = q != nullptr ? ld_cg4(q + off) : zero4;
        };
        auto split4 = [](const float4& x, float4& hi, float4& lo) {
            hi.x = __uint_as_float(__float_as_uint(x.x) & 0xffffe000u); lo.x = x.x - hi.x;
            hi.y = __uint_as_float(__float_as_uint(x.y) & 0xffffe000u); lo.y = x.y - hi.y;
            hi.z = __uint_as_float(__float_as_uint(x.z) & 0xffffe000u); lo.z = x.z - hi.z;
            hi.w = __uint_as_float(__float_as_uint(x.w) & 0xffffe000u); lo.w = x.w - hi.w;
        };
        auto produce = [&](int kb, float4 (&xa)[NA], float4& xb) {
            const uint32_t n = kb_base + (uint32_t)kb;
            const int s = (int)(n % RTC_STAGES);
            if (lane == 0) mbar_wait_backoff(empty0 + 8 * s, ((n / RTC_STAGES) & 1) ^ 1);     // one poller per warp, sleeping between probes
            __syncwarp();
            uint8_t* stg = st.ring + s * RTC_STAGE_BYTES;
#pragma unroll
            for (int i = 0; i < NA; ++i) {
                const int r = r0 + 32 * i;
                if (r < WR && !(st.dbg & 1)) {
                    const uint32_t off = sw128_off(r, c);
                    float4 hi, lo;
                    split4(xa[i], hi, lo);
                    *reinterpret_cast<float4*>(stg + off) = hi;
                    *reinterpret_cast<float4*>(stg + RTC_A_BYTES + off) = lo;
                }
            }
            if (r0 < NB) {
                const uint32_t off = sw128_off(r0, c);
                float4 hi, lo;
                split4(xb, hi, lo);
                *reinterpret_cast<float4*>(stg + 2 * RTC_A_BYTES + off) = hi;
                *reinterpret_cast<float4*>(stg + 2 * RTC_A_BYTES + RTC_B_BYTES + off) = lo;
            }
            if (kb + RTC_PF < nkb && !(st.dbg & 1)) load(kb + RTC_PF, xa, xb);
            if (!(st.dbg & 4)) asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // generic-proxy writes -> visible to the tensor core
            __syncwarp();
            if (lane == 0) mbar_arrive(full0 + 8 * s);
        };
#pragma unroll
        for (int j = 0; j < RTC_PF; ++j)
            if (j < nkb) load(j, pa[j], pb[j]);
#pragma unroll 1
        for (int kb0 = 0; kb0 < nkb; kb0 += RTC_PF) {
#pragma unroll
            for (int j = 0; j < RTC_PF; ++j)
                if (kb0 + j < nkb) produce(kb0 + j, pa[j], pb[j]);
        }
    } else {
        // ------------------------------ MMA issuer ------------------------------
        const uint32_t idesc = umma_idesc_tf32(128, NB);
        const uint64_t dbase = umma_desc(ring_u32);                // descriptors of other tiles / k-steps differ in the address field only
#pragma unroll 1
        for (int kb = 0; kb < nkb; ++kb) {
            const uint32_t n = kb_base + (uint32_t)kb;
            const int s = (int)(n % RTC_STAGES);
            if (lane == 0) {
                mbar_wait(full0 + 8 * s, (n / RTC_STAGES) & 1);
                tc_fence_after();
                const uint64_t a_hi = dbase + (uint64_t)((s * RTC_STAGE_BYTES) >> 4), a_lo = a_hi + (RTC_A_BYTES >> 4);
                const uint64_t b_hi = a_hi + ((2 * RTC_A_BYTES) >> 4), b_lo = b_hi + (RTC_B_BYTES >> 4);
                const uint32_t acc = kb != 0;
#pragma unroll
                for (int kk = 0; kk < RTC_BK / 8; ++kk) {
                    if ((st.dbg & 2) && kb != 0) continue;
                    // 12 independent accumulators (4 k-steps x 3 split products): back-to-back MMAs never wait on each other's D;
                    // 8 tf32 = 32 bytes along the swizzled row = 2 in the descriptor's 16-byte units
                    const uint32_t d0 = st.tmem_base + (uint32_t)(kk * 3) * 32;
                    umma_tf32(d0, a_lo + 2 * kk, b_hi + 2 * kk, idesc, acc);
                    umma_tf32(d0 + 32, a_hi + 2 * kk, b_lo + 2 * kk, idesc, acc);
                    umma_tf32(d0 + 64, a_hi + 2 * kk, b_hi + 2 * kk, idesc, acc);
                }
                umma_commit(empty0 + 8 * s);                  // frees the stage when these MMAs have read it
                if (kb == nkb - 1) umma_commit(tfull);        // accumulator complete
            }
            __syncwarp();
        }
    }
    // ------------------------------ accumulator -> per-thread sums ------------------------------
    mbar_wait(tfull, st.tiles & 1);
    tc_fence_after();
    if (warp < REC_WARPS && (warp & 3) * 32 < WR) {            // TMEM lane quarter (w & 3) = weight rows [32(w&3), +32); two warps per quarter
        const int q = warp & 3, half = warp >> 2;
        float* dst = st.outbuf + half * (128 * RTC_OUT_LD) + (q * 32 + lane) * RTC_OUT_LD;
        const uint32_t t0 = st.tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * (RTC_ACCS / 2)) * 32;
        if (NB == 32) {
            float sum[32], v[32];
            tmem_ld32(t0, sum);
#pragma unroll 1
            for (int a = 1; a < RTC_ACCS / 2; ++a) {
                tmem_ld32(t0 + (uint32_t)a * 32, v);
#pragma unroll
                for (int j = 0; j < 32; ++j) sum[j] += v[j];
            }
#pragma unroll
            for (int j = 0; j < 32; ++j) dst[j] = sum[j];
        } else {
            float sum[16], v[16];
            tmem_ld16(t0, sum);
#pragma unroll 1
            for (int a = 1; a < RTC_ACCS / 2; ++a) {
                tmem_ld16(t0 + (uint32_t)a * 32, v);
#pragma unroll
                for (int j = 0; j < 16; ++j) sum[j] += v[j];
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) dst[j] = sum[j];
        }
    }
    tc_fence_before();
    __syncthreads();                                           // also orders the tcgen05.ld before the next tile's first MMA
    if (tid < REC_THREADS) {
        const int u = tid & 15;
#pragma unroll
        for (int p = 0; p < (NT + 1) / 2; ++p) {
            const int row = (tid >> 4) + 16 * p;
            if (row < NB) {
#pragma unroll
                for (int g = 0; g < NG; ++g)
                    out[g][p] = st.outbuf[(g * REC_J + u) * RTC_OUT_LD + row] + st.outbuf[128 * RTC_OUT_LD + (g * REC_J + u) * RTC_OUT_LD + row];
            }
        }
    }
    st.kb_count += (uint32_t)nkb;
    st.tiles += 1;
    __syncthreads();                                           // outbuf may be rewritten by the next tile
}

}  // namespace tg
