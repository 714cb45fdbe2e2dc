// Gate tile with ON-CHIP RESIDENT weights for the persistent segment-level kernel (third variant of recurrent.cuh's tile).
//
// Measured (profiles/r01_segment_phase_timing.txt): the cell phase of a segment step streams 57 MB through L2 at ~6.7 TB/s —
// the L2->SM fabric limit — and two thirds of it are the cell weights, which are the SAME on every one of the T steps.  In the
// persistent kernel a CTA owns the same cell tile on every step, so its weights can stay on the SM:
//   * each thread keeps ITS OWN mma.sync A-fragments (the 4 fp32 values of a 16x8 weight block that this lane feeds to
//     mma.m16n8k8) for all K chunks its warp processes: 288 words per thread at D = 512;
//   * 256 of them live in TENSOR MEMORY, used as a per-thread register-file extension: tcgen05.st 32x32b once at kernel start,
//     tcgen05.ld 32x32b.x4 inside the K loop (lane i of warp w owns TMEM lane 32*(w%4)+i; warps w and w+4 share a lane
//     quarter and take columns [0,256) / [256,512)); the remaining words sit in shared memory, one float4 column per thread.
//     Because every lane reads back exactly what it stored, no layout conversion is ever needed;
//   * only the ACTIVATION rows (aggregated messages, previous state: 8*NT rows) still stream through the warp-private
//     cp.async ring, so the ring is a third of its old size and the L2 traffic of the phase drops to the activations.
// Arithmetic: 3xFP16 split products on mma.sync.m16n8k16 (x = hi + lo with hi = fp16(x), lo = fp16(x - hi): 22 mantissa bits;
// hi*hi + hi*lo + lo*hi, fp32 accumulate) — the same accuracy class as the 3xTF32 split of the streaming tile (measured: max
// error 1.7e-7 on a K = 1536 gate pre-activation against 7.8e-7 for 3xTF32 and 2.4e-6 for a sequential fp32 sum) at HALF the
// tensor-pipe instructions, which is what bounds this tile once the weights no longer stream (ncu r01: tensor pipe 24 % of the
// whole kernel, math_pipe_throttle).  The resident weights are stored PRE-SPLIT: one (hi, lo) pair of f16x2 words takes exactly
// the space of the two fp32 values it replaces, so the per-step weight split disappears from the K loop.  Weights are scaled by
// 2^8 before the split (keeps the lo parts of default-init-sized weights out of the fp16 subnormals; undone on the reduced sums);
// a weight beyond 65504 / 2^8 or an activation beyond 65504 raises the kernel's range flag (GridSync::error bit 1).
// K split over the 8 warps, partials reduced through shared memory (in two halves: the reduction buffer is small).
#pragma once
#include <cuda_fp16.h>
#include "recurrent.cuh"
#include "tcgen05.cuh"

namespace tg {

constexpr int RES_TMEM_WORDS = 256;          // per-thread words in tensor memory (512 columns shared by two warps per lane quarter)
constexpr int RES_SMEM_WORDS = 32;           // per-thread overflow words in shared memory
constexpr int RES_GROUPS = 3;                // weight groups active in a K chunk of the cell tile: (r, z, n_i) or (r, z, n_h)
constexpr int RES_CHUNK_WORDS = 2 * RES_GROUPS * 4;   // one k16 step x 3 groups x (hi, lo) x 4 fragment registers
constexpr int RES_RS = 16;                   // ring row stride in floats: no padding, the 16-byte quarters of a row are XOR-swizzled
constexpr int RES_STAGES = 3;                // cp.async ring depth per warp (activation rows only)
constexpr int RES_MSG_GROUPS = 4;            // message tile: 4 resident weight groups of 16 units (+ the receivers' rows as a streamed 5th)

// Ring layout: row r, 16-byte quarter q lives at float offset r*16 + 4*(q ^ (r & 2)); with it the 64-bit fragment loads of a
// half-warp (rows g8 = 0..3 or 4..7, columns 2 t4) hit 16 distinct 8-byte slots.
__device__ __forceinline__ int res_ring_off(int row, int quarter) { return row * RES_RS + ((quarter ^ (row & 2)) << 2); }
// float offset of column 2 t4 (first fragment register) within a row whose index is congruent to g8 modulo 8; the second
// register (column + 8) is at this offset ^ 8
__device__ __forceinline__ int res_frag_col(int g8, int t4) { return ((((t4 >> 1) ^ (g8 & 2))) << 2) + ((t4 & 1) << 1); }
constexpr float RES_WSCALE = 256.0f;         // weights are split as fp16(w * 2^8)
constexpr float RES_F16_MAX = 65504.0f;

struct ResState {
    uint32_t tmem_base;
    uint4* wovf;             // [RES_SMEM_WORDS / 4][REC_THREADS] overflow fragments
    int ready;               // fragments of this CTA's cell tile are loaded
};

__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint32_t (&v)[4]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};\n" ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3])
                 : "memory");
}

// D += A(16x16, row) * B(16x8, col), fp16 operands, fp32 accumulate.
__device__ __forceinline__ void mma_f16(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// (x0, x1) = hi + lo, both f16x2 words with x0 in the low half (the smaller k index of an MMA fragment register).
__device__ __forceinline__ void split_f16x2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
    const __half2 h = __floats2half2_rn(x0, x1);
    const float2 f = __half22float2(h);
    const __half2 l = __floats2half2_rn(x0 - f.x, x1 - f.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}
__device__ __forceinline__ void tmem_ld4_nowait(uint32_t taddr, uint32_t (&r)[4]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];\n" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }

// TMEM address of word j of this thread
__device__ __forceinline__ uint32_t res_taddr(const ResState& rs, int j) {
    const int warp = threadIdx.x >> 5;
    return rs.tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * RES_TMEM_WORDS + j);
}

// Once per kernel (all threads): allocate tensor memory.  Pair with res_finish on every exit path.
__device__ __forceinline__ void res_init(ResState& rs, uint32_t* tmem_slot, uint4* wovf) {
    if ((threadIdx.x >> 5) == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    rs.tmem_base = *tmem_slot;
    rs.wovf = wovf;
    rs.ready = 0;
}
__device__ __forceinline__ void res_finish(ResState& rs) {
    tc_fence_before();
    __syncthreads();
    if ((threadIdx.x >> 5) == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(rs.tmem_base), "r"(512u) : "memory");
    }
}

// Load this thread's weight fragments of one cell tile, pre-split into fp16 (hi, lo) MMA A-fragments.
// wrow1[g*16 + u] / wrow2[...]: weight row pointers of K segment 1 / 2 (null = group absent in that segment), exactly the first
// 64 entries of the streaming tile's pointer tables.  A-fragment of m16n8k16 for lane (g8, t4): register 0 = row g8, k = 2 t4 + {0,1};
// 1 = row g8 + 8, same k; 2 = row g8, k + 8; 3 = row g8 + 8, k + 8.
// Words of chunk n (chunk id = warp + 8 n), active group gi, half hl (0 hi, 1 lo), register r:  j = ((n*3 + gi)*2 + hl)*4 + r.
// Returns false (uniformly) when the tile needs more words than fit; *range_bad is set when a weight leaves the fp16 split range.
__device__ __forceinline__ bool res_fill_cell(ResState& rs, const float* const* wrow1, const float* const* wrow2, int K1, int K2,
                                              unsigned int* range_flag) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g8 = lane >> 2, t4 = lane & 3;
    const int chunks1 = K1 / REC_CK, total = (K1 + K2) / REC_CK;
    const int nmine = warp < total ? (total - warp + REC_WARPS - 1) / REC_WARPS : 0;
    const int nmax = (total + REC_WARPS - 1) / REC_WARPS;
    if (nmax * RES_CHUNK_WORDS > RES_TMEM_WORDS + RES_SMEM_WORDS) return false;
    float wmax = 0.0f;
    for (int n = 0; n < nmine; ++n) {
        const int chunk = warp + n * REC_WARPS;
        const bool seg1 = chunk < chunks1;
        const int k = (seg1 ? chunk : chunk - chunks1) * REC_CK + 2 * t4;
#pragma unroll
        for (int gi = 0; gi < RES_GROUPS; ++gi) {
            const int g = (gi == 2 && !seg1) ? 3 : gi;                         // seg 1: r, z, n_i ; seg 2: r, z, n_h
            const float* r0 = (seg1 ? wrow1 : wrow2)[g * REC_J + g8];
            const float* r1 = (seg1 ? wrow1 : wrow2)[g * REC_J + g8 + 8];
            float2 v[4];
            v[0] = r0 != nullptr ? __ldg(reinterpret_cast<const float2*>(r0 + k)) : make_float2(0.f, 0.f);
            v[1] = r1 != nullptr ? __ldg(reinterpret_cast<const float2*>(r1 + k)) : make_float2(0.f, 0.f);
            v[2] = r0 != nullptr ? __ldg(reinterpret_cast<const float2*>(r0 + k + 8)) : make_float2(0.f, 0.f);
            v[3] = r1 != nullptr ? __ldg(reinterpret_cast<const float2*>(r1 + k + 8)) : make_float2(0.f, 0.f);
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                wmax = fmaxf(wmax, fmaxf(fabsf(v[r].x), fabsf(v[r].y)));
                split_f16x2(v[r].x * RES_WSCALE, v[r].y * RES_WSCALE, hi[r], lo[r]);
            }
            const int j = ((n * RES_GROUPS + gi) * 2) * 4;
            // warp-uniform branches (j depends on n and gi only)
            if (j < RES_TMEM_WORDS) tmem_st4(res_taddr(rs, j), hi);
            else rs.wovf[((j - RES_TMEM_WORDS) >> 2) * REC_THREADS + tid] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            if (j + 4 < RES_TMEM_WORDS) tmem_st4(res_taddr(rs, j + 4), lo);
            else rs.wovf[((j + 4 - RES_TMEM_WORDS) >> 2) * REC_THREADS + tid] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
    }
    if (!(wmax * RES_WSCALE < RES_F16_MAX)) atomicOr(range_flag, 2u);        // also catches NaN weights
    tmem_wait_st();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    rs.ready = 1;
    return true;
}

// out[g][p] as tile_accumulate<4, NT, 3> delivers it (groups 0 r, 1 z, 2 n_i, 3 n_h), weights from the resident fragments,
// activation rows act1 (K segment 1) / act2 (segment 2) through the ring.  K1, K2 multiples of 16 and the same on every call.
template <int NT>
__device__ __forceinline__ void tile_accumulate_res(float (&out)[4][(NT + 1) / 2], const float* const* act1, const float* const* act2,
                                                    int K1, int K2, const ResState& rs, const float* gdummy, float* smem) {
    constexpr int NG = 4, STAGES = RES_STAGES, ROWS = 8 * NT;
    constexpr int STAGE_F = ROWS * RES_RS;
    constexpr int NP = (ROWS * 4 + 31) / 32;            // 16-byte pieces per lane per chunk
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g8 = lane >> 2, t4 = lane & 3, fc = res_frag_col(g8, t4);
    float* ring = smem + warp * (STAGES * STAGE_F);

    float c[NG][NT][4];
#pragma unroll
    for (int m = 0; m < NG; ++m)
#pragma unroll
        for (int n = 0; n < NT; ++n)
#pragma unroll
            for (int r = 0; r < 4; ++r) c[m][n][r] = 0.0f;

    const int chunks1 = K1 / REC_CK, total_chunks = (K1 + K2) / REC_CK;
    const int nmine = warp < total_chunks ? (total_chunks - warp + REC_WARPS - 1) / REC_WARPS : 0;
    if (nmine > 0) {
        const float* s1[NP];
        const float* s2[NP];
        int dst[NP];
        bool in[NP];
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            const int piece = lane + p * 32;
            const int row = piece >> 2, quarter = piece & 3;
            in[p] = piece < ROWS * 4;
            const float* b1 = in[p] ? act1[row] : nullptr;
            const float* b2 = in[p] ? act2[row] : nullptr;
            s1[p] = b1 != nullptr ? b1 + quarter * 4 : nullptr;
            s2[p] = b2 != nullptr ? b2 + quarter * 4 - K1 : nullptr;      // indexed with the global k offset
            dst[p] = res_ring_off(row, quarter);
        }
        auto issue = [&](int n, int st) {
            const int chunk = warp + n * REC_WARPS;
            const bool seg1 = chunk < chunks1;
            const size_t koff = (size_t)chunk * REC_CK;
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                if (!in[p]) continue;
                const float* base = seg1 ? s1[p] : s2[p];
                const bool ok = base != nullptr;
                cp_async16_zfill(ring + st * STAGE_F + dst[p], ok ? base + koff : gdummy, ok);
            }
        };
#pragma unroll
        for (int st = 0; st < STAGES - 1; ++st) {
            if (st < nmine) issue(st, st);
            cp_async_commit();
        }
#pragma unroll 1
        for (int n = 0; n < nmine; ++n) {
            // this chunk's pre-split weight fragments: tensor memory (or the shared-memory overflow), fetched while the copies land
            uint32_t w[RES_GROUPS][2][4];
            const int j0 = n * RES_CHUNK_WORDS;
#pragma unroll
            for (int gi = 0; gi < RES_GROUPS; ++gi)
#pragma unroll
                for (int hl = 0; hl < 2; ++hl) {
                    const int j = j0 + (gi * 2 + hl) * 4;
                    if (j < RES_TMEM_WORDS) {
                        tmem_ld4_nowait(res_taddr(rs, j), w[gi][hl]);
                    } else {
                        const uint4 v = rs.wovf[((j - RES_TMEM_WORDS) >> 2) * REC_THREADS + tid];
                        w[gi][hl][0] = v.x; w[gi][hl][1] = v.y; w[gi][hl][2] = v.z; w[gi][hl][3] = v.w;
                    }
                }
            cp_async_wait<STAGES - 2>();
            __syncwarp();
            {
                const int nn = n + STAGES - 1;
                if (nn < nmine) issue(nn, nn % STAGES);
                cp_async_commit();
            }
            const bool seg1 = (warp + n * REC_WARPS) < chunks1;
            // B fragments of m16n8k16 for lane (g8, t4): register 0 = (k = 2 t4 + {0,1}, row g8), register 1 = (k + 8, row g8)
            const float* xb = ring + (n % STAGES) * STAGE_F + g8 * RES_RS;
            uint32_t bh[NT][2], bl[NT][2];
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                const float2 v0 = *reinterpret_cast<const float2*>(xb + nt * 8 * RES_RS + fc);
                const float2 v1 = *reinterpret_cast<const float2*>(xb + nt * 8 * RES_RS + (fc ^ 8));
                split_f16x2(v0.x, v0.y, bh[nt][0], bl[nt][0]);
                split_f16x2(v1.x, v1.y, bh[nt][1], bl[nt][1]);
            }
            tmem_wait_ld();
#pragma unroll
            for (int gi = 0; gi < RES_GROUPS; ++gi) {
                // seg 1 feeds groups (r, z, n_i) = accumulators 0, 1, 2; seg 2 feeds (r, z, n_h) = 0, 1, 3
                if (gi < 2 || seg1) {
                    float (&cc)[NT][4] = c[gi];
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt) mma_f16(cc[nt], w[gi][1], bh[nt]);
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt) mma_f16(cc[nt], w[gi][0], bl[nt]);
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt) mma_f16(cc[nt], w[gi][0], bh[nt]);
                } else {
                    float (&cc)[NT][4] = c[3];
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt) mma_f16(cc[nt], w[gi][1], bh[nt]);
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt) mma_f16(cc[nt], w[gi][0], bl[nt]);
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt) mma_f16(cc[nt], w[gi][0], bh[nt]);
                }
            }
        }
        cp_async_wait<0>();
    }
    __syncthreads();                             // every warp's ring is dead: reuse the memory for the reduction
    // cross-warp reduction of the K split in two halves (the buffer holds 4 warps' partials): red[w4][m][n][reg][lane]
    float* red = smem;
    if (warp >= 4) {
#pragma unroll
        for (int m = 0; m < NG; ++m)
#pragma unroll
            for (int n = 0; n < NT; ++n)
#pragma unroll
                for (int r = 0; r < 4; ++r) red[((((warp - 4) * NG + m) * NT + n) * 4 + r) * 32 + lane] = c[m][n][r];
    }
    __syncthreads();
    if (warp < 4) {
#pragma unroll
        for (int m = 0; m < NG; ++m)
#pragma unroll
            for (int n = 0; n < NT; ++n)
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const int idx = (((warp * NG + m) * NT + n) * 4 + r) * 32 + lane;
                    red[idx] += c[m][n][r];
                }
    }
    __syncthreads();
#pragma unroll
    for (int p = 0; p < (NT + 1) / 2; ++p) {
        const int u = tid & 15, row = (tid >> 4) + 16 * p;
        const int n = row >> 3, col = row & 7;
        const int l = (u & 7) * 4 + (col >> 1), r = (u >> 3) * 2 + (col & 1);
#pragma unroll
        for (int m = 0; m < NG; ++m) {
            float s = 0.0f;
            if (n < NT) {
#pragma unroll
                for (int w4 = 0; w4 < 4; ++w4) s += red[(((w4 * NG + m) * NT + n) * 4 + r) * 32 + l];
            }
            out[m][p] = s * (1.0f / RES_WSCALE);
        }
    }
    __syncthreads();                             // smem may be reused by the caller right away
}

// ---- message tile with SMEM-RESIDENT weights ---------------------------------------------------------------------------------
// The message MLP rows of a tile (4 groups x 16 units x K) never change either: they sit in shared memory as pre-split fp16
// A-fragments in per-thread order, wmsg[((n*4 + g)*2 + hl) * REC_THREADS + tid] = the four fragment registers of chunk
// (warp + 8 n), group g, half hl — one conflict-free LDS.128 each.  Only the 16 receiver rows (the A operand of the logit group)
// and the 16 sender rows (B operand) stream through the ring: 64 KB per step instead of 224 KB.
__device__ __forceinline__ void res_fill_msg(uint4* wmsg, const float* const* wrow, int K, unsigned int* range_flag) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g8 = lane >> 2, t4 = lane & 3;
    const int total = K / REC_CK;
    const int nmine = warp < total ? (total - warp + REC_WARPS - 1) / REC_WARPS : 0;
    float wmax = 0.0f;
    for (int n = 0; n < nmine; ++n) {
        const int k = (warp + n * REC_WARPS) * REC_CK + 2 * t4;
#pragma unroll
        for (int g = 0; g < RES_MSG_GROUPS; ++g) {
            const float* r0 = wrow[g * REC_J + g8];
            const float* r1 = wrow[g * REC_J + g8 + 8];
            float2 v[4];
            v[0] = r0 != nullptr ? __ldg(reinterpret_cast<const float2*>(r0 + k)) : make_float2(0.f, 0.f);
            v[1] = r1 != nullptr ? __ldg(reinterpret_cast<const float2*>(r1 + k)) : make_float2(0.f, 0.f);
            v[2] = r0 != nullptr ? __ldg(reinterpret_cast<const float2*>(r0 + k + 8)) : make_float2(0.f, 0.f);
            v[3] = r1 != nullptr ? __ldg(reinterpret_cast<const float2*>(r1 + k + 8)) : make_float2(0.f, 0.f);
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                wmax = fmaxf(wmax, fmaxf(fabsf(v[r].x), fabsf(v[r].y)));
                split_f16x2(v[r].x * RES_WSCALE, v[r].y * RES_WSCALE, hi[r], lo[r]);
            }
            wmsg[((n * RES_MSG_GROUPS + g) * 2 + 0) * REC_THREADS + tid] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            wmsg[((n * RES_MSG_GROUPS + g) * 2 + 1) * REC_THREADS + tid] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
    }
    if (!(wmax * RES_WSCALE < RES_F16_MAX)) atomicOr(range_flag, 2u);
    __syncthreads();
}

// out[g][0] for g < 4: message pre-activations (unit = tid % 16 of group g, sender row = tid / 16); out[4][0]: <receiver tid % 16,
// sender tid / 16> — what tile_accumulate<5, 2, 3> delivers.  act[0..15]: receiver rows, act[16..31]: sender rows (null = zeros).
__device__ __forceinline__ void tile_accumulate_msg_res(float (&out)[RES_MSG_GROUPS + 1][1], const float* const* act, int K,
                                                        const uint4* wmsg, const float* gdummy, float* smem) {
    constexpr int NG = RES_MSG_GROUPS + 1, NT = 2, STAGES = RES_STAGES, ROWS = 32;
    constexpr int STAGE_F = ROWS * RES_RS;
    constexpr int NP = ROWS * 4 / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g8 = lane >> 2, t4 = lane & 3, fc = res_frag_col(g8, t4);
    float* ring = smem + warp * (STAGES * STAGE_F);

    float c[NG][NT][4];
#pragma unroll
    for (int m = 0; m < NG; ++m)
#pragma unroll
        for (int n = 0; n < NT; ++n)
#pragma unroll
            for (int r = 0; r < 4; ++r) c[m][n][r] = 0.0f;

    const int total_chunks = K / REC_CK;
    const int nmine = warp < total_chunks ? (total_chunks - warp + REC_WARPS - 1) / REC_WARPS : 0;
    if (nmine > 0) {
        const float* src[NP];
        int dst[NP];
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            const int piece = lane + p * 32;
            const int row = piece >> 2, quarter = piece & 3;
            const float* b = act[row];
            src[p] = b != nullptr ? b + quarter * 4 : nullptr;
            dst[p] = res_ring_off(row, quarter);
        }
        auto issue = [&](int n, int st) {
            const size_t koff = (size_t)(warp + n * REC_WARPS) * REC_CK;
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                const bool ok = src[p] != nullptr;
                cp_async16_zfill(ring + st * STAGE_F + dst[p], ok ? src[p] + koff : gdummy, ok);
            }
        };
#pragma unroll
        for (int st = 0; st < STAGES - 1; ++st) {
            if (st < nmine) issue(st, st);
            cp_async_commit();
        }
#pragma unroll 1
        for (int n = 0; n < nmine; ++n) {
            uint4 w[RES_MSG_GROUPS][2];
#pragma unroll
            for (int g = 0; g < RES_MSG_GROUPS; ++g) {
                w[g][0] = wmsg[((n * RES_MSG_GROUPS + g) * 2 + 0) * REC_THREADS + tid];
                w[g][1] = wmsg[((n * RES_MSG_GROUPS + g) * 2 + 1) * REC_THREADS + tid];
            }
            cp_async_wait<STAGES - 2>();
            __syncwarp();
            {
                const int nn = n + STAGES - 1;
                if (nn < nmine) issue(nn, nn % STAGES);
                cp_async_commit();
            }
            const float* xb = ring + (n % STAGES) * STAGE_F + g8 * RES_RS;
            // logit group: A fragments from the receiver rows 0..15
            uint32_t rh[4], rl[4];
            {
                const float2 x0 = *reinterpret_cast<const float2*>(xb + fc);
                const float2 x1 = *reinterpret_cast<const float2*>(xb + 8 * RES_RS + fc);
                const float2 x2 = *reinterpret_cast<const float2*>(xb + (fc ^ 8));
                const float2 x3 = *reinterpret_cast<const float2*>(xb + 8 * RES_RS + (fc ^ 8));
                split_f16x2(x0.x, x0.y, rh[0], rl[0]);
                split_f16x2(x1.x, x1.y, rh[1], rl[1]);
                split_f16x2(x2.x, x2.y, rh[2], rl[2]);
                split_f16x2(x3.x, x3.y, rh[3], rl[3]);
            }
            // B fragments from the sender rows 16..31
            uint32_t bh[NT][2], bl[NT][2];
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                const float2 v0 = *reinterpret_cast<const float2*>(xb + (16 + nt * 8) * RES_RS + fc);
                const float2 v1 = *reinterpret_cast<const float2*>(xb + (16 + nt * 8) * RES_RS + (fc ^ 8));
                split_f16x2(v0.x, v0.y, bh[nt][0], bl[nt][0]);
                split_f16x2(v1.x, v1.y, bh[nt][1], bl[nt][1]);
            }
#pragma unroll
            for (int g = 0; g < RES_MSG_GROUPS; ++g) {
                const uint32_t ah[4] = {w[g][0].x, w[g][0].y, w[g][0].z, w[g][0].w};
                const uint32_t al[4] = {w[g][1].x, w[g][1].y, w[g][1].z, w[g][1].w};
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) mma_f16(c[g][nt], al, bh[nt]);
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) mma_f16(c[g][nt], ah, bl[nt]);
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) mma_f16(c[g][nt], ah, bh[nt]);
            }
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) mma_f16(c[RES_MSG_GROUPS][nt], rl, bh[nt]);
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) mma_f16(c[RES_MSG_GROUPS][nt], rh, bl[nt]);
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) mma_f16(c[RES_MSG_GROUPS][nt], rh, bh[nt]);
        }
        cp_async_wait<0>();
    }
    __syncthreads();                             // every warp's ring is dead: reuse the memory for the reduction
    float* red = smem;                           // red[warp][m][n][reg][lane]: 8 x 5 x 2 x 4 x 32 floats = 40 KB (ring: 48 KB)
#pragma unroll
    for (int m = 0; m < NG; ++m)
#pragma unroll
        for (int n = 0; n < NT; ++n)
#pragma unroll
            for (int r = 0; r < 4; ++r) red[(((warp * NG + m) * NT + n) * 4 + r) * 32 + lane] = c[m][n][r];
    __syncthreads();
    {
        const int u = tid & 15, row = tid >> 4;
        const int n = row >> 3, col = row & 7;
        const int l = (u & 7) * 4 + (col >> 1), r = (u >> 3) * 2 + (col & 1);
#pragma unroll
        for (int m = 0; m < NG; ++m) {
            float s = 0.0f;
#pragma unroll
            for (int w = 0; w < REC_WARPS; ++w) s += red[(((w * NG + m) * NT + n) * 4 + r) * 32 + l];
            out[m][0] = m < RES_MSG_GROUPS ? s * (1.0f / RES_WSCALE) : s;
        }
    }
    __syncthreads();
}


// ---- fp32 resident fragments for the BACKWARD recurrences (3xTF32) -------------------------------------------------------------
// The BPTT tiles multiply gradient rows (far below the fp16 range, different at every step) by TRANSPOSED weights that never
// change: the weights stay on chip as raw fp32 mma.m16n8k8 A-fragment words (tensor memory first, shared-memory overflow after
// RES_TMEM_WORDS words per thread) and are split into TF32 (hi, lo) in registers, exactly as the streaming tile does after its
// ring loads — only the activation rows still stream (ncu r02: segment_bwd moved 424 MB of DRAM and ~110 MB of L2 -> SM traffic
// per reverse step, 60 % of it the same weights again).  Several tiles of a CTA share the per-thread word space through `word0`.
// Word of chunk n (chunk id = warp + 8 n), group m, k8 step kk, register r:  j = word0 + ((n*NG + m)*2 + kk)*4 + r.
__device__ __forceinline__ uint32_t res32_taddr(const ResState& rs, int j) {
    const int warp = threadIdx.x >> 5;
    return rs.tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * RES_TMEM_WORDS + j);
}

// wrow[g*16 + u]: K-contiguous weight row pointers (K = K1 of tab1 followed by K2 of tab2, null = zero row).  `ovf`: shared-memory
// overflow [(words beyond RES_TMEM_WORDS) / 4][REC_THREADS] uint4.  Returns the number of words this tile occupies per thread.
template <int NG>
__device__ __forceinline__ int res32_fill(const ResState& rs, uint4* ovf, int word0, const float* const* wrow1, const float* const* wrow2,
                                          int K1, int K2) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g8 = lane >> 2, t4 = lane & 3;
    const int chunks1 = K1 / REC_CK, total = (K1 + K2) / REC_CK;
    const int nmax = (total + REC_WARPS - 1) / REC_WARPS;
    const int nmine = warp < total ? (total - warp + REC_WARPS - 1) / REC_WARPS : 0;
    for (int n = 0; n < nmine; ++n) {
        const int chunk = warp + n * REC_WARPS;
        const bool seg1 = chunk < chunks1;
        const int k = (seg1 ? chunk : chunk - chunks1) * REC_CK + t4;
#pragma unroll
        for (int m = 0; m < NG; ++m) {
            const float* r0 = (seg1 ? wrow1 : wrow2)[m * REC_J + g8];
            const float* r1 = (seg1 ? wrow1 : wrow2)[m * REC_J + g8 + 8];
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
                uint32_t w4[4];
                w4[0] = r0 != nullptr ? __float_as_uint(__ldg(r0 + k + kk * 8)) : 0u;
                w4[1] = r1 != nullptr ? __float_as_uint(__ldg(r1 + k + kk * 8)) : 0u;
                w4[2] = r0 != nullptr ? __float_as_uint(__ldg(r0 + k + kk * 8 + 4)) : 0u;
                w4[3] = r1 != nullptr ? __float_as_uint(__ldg(r1 + k + kk * 8 + 4)) : 0u;
                const int j = word0 + ((n * NG + m) * 2 + kk) * 4;
                if (j < RES_TMEM_WORDS) tmem_st4(res32_taddr(rs, j), w4);           // (warp-uniform: j depends on n, m, kk only)
                else ovf[((j - RES_TMEM_WORDS) >> 2) * REC_THREADS + tid] = make_uint4(w4[0], w4[1], w4[2], w4[3]);
            }
        }
    }
    tmem_wait_st();
    return nmax * NG * 8;
}

// out[g][p] as tile_accumulate<NG, NT, 3> delivers it; weights from the resident fp32 fragments (word0 as given to res32_fill),
// activation rows act1 (K segment 1) / act2 (segment 2) through the warp-private cp.async ring.
template <int NG, int NT>
__device__ __forceinline__ void tile_accumulate_res32(float (&out)[NG][(NT + 1) / 2], const float* const* act1, const float* const* act2,
                                                      int K1, int K2, const ResState& rs, const uint4* ovf, int word0, const float* gdummy,
                                                      float* smem) {
    constexpr int STAGES = RES_STAGES, ROWS = 8 * NT;
    constexpr int STAGE_F = ROWS * REC_RS;
    constexpr int NP = (ROWS * 4 + 31) / 32;            // 16-byte pieces per lane per chunk
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g8 = lane >> 2, t4 = lane & 3;
    float* ring = smem + warp * (STAGES * STAGE_F);

    float c[NG][NT][4];
#pragma unroll
    for (int m = 0; m < NG; ++m)
#pragma unroll
        for (int n = 0; n < NT; ++n)
#pragma unroll
            for (int r = 0; r < 4; ++r) c[m][n][r] = 0.0f;

    const int chunks1 = K1 / REC_CK, total_chunks = (K1 + K2) / REC_CK;
    const int nmine = warp < total_chunks ? (total_chunks - warp + REC_WARPS - 1) / REC_WARPS : 0;
    if (nmine > 0) {
        const float* s1[NP];
        const float* s2[NP];
        int dst[NP];
        bool in[NP];
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            const int piece = lane + p * 32;
            const int row = piece >> 2, quarter = piece & 3;
            in[p] = piece < ROWS * 4;
            const float* b1 = in[p] ? act1[row] : nullptr;
            const float* b2 = (in[p] && K2 > 0) ? act2[row] : nullptr;
            s1[p] = b1 != nullptr ? b1 + quarter * 4 : nullptr;
            s2[p] = b2 != nullptr ? b2 + quarter * 4 - K1 : nullptr;      // indexed with the global k offset
            dst[p] = row * REC_RS + quarter * 4;
        }
        auto issue = [&](int n, int st) {
            const int chunk = warp + n * REC_WARPS;
            const bool seg1 = chunk < chunks1;
            const size_t koff = (size_t)chunk * REC_CK;
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                if (!in[p]) continue;
                const float* base = seg1 ? s1[p] : s2[p];
                const bool ok = base != nullptr;
                cp_async16_zfill(ring + st * STAGE_F + dst[p], ok ? base + koff : gdummy, ok);
            }
        };
#pragma unroll
        for (int st = 0; st < STAGES - 1; ++st) {
            if (st < nmine) issue(st, st);
            cp_async_commit();
        }
#pragma unroll 1
        for (int n = 0; n < nmine; ++n) {
            // this chunk's weight words: tensor memory (or the shared-memory overflow), fetched while the copies land
            uint32_t w[NG][2][4];
#pragma unroll
            for (int m = 0; m < NG; ++m)
#pragma unroll
                for (int kk = 0; kk < 2; ++kk) {
                    const int j = word0 + ((n * NG + m) * 2 + kk) * 4;
                    if (j < RES_TMEM_WORDS) {
                        tmem_ld4_nowait(res32_taddr(rs, j), w[m][kk]);
                    } else {
                        const uint4 v = ovf[((j - RES_TMEM_WORDS) >> 2) * REC_THREADS + tid];
                        w[m][kk][0] = v.x; w[m][kk][1] = v.y; w[m][kk][2] = v.z; w[m][kk][3] = v.w;
                    }
                }
            cp_async_wait<STAGES - 2>();
            __syncwarp();
            {
                const int nn = n + STAGES - 1;
                if (nn < nmine) issue(nn, nn % STAGES);
                cp_async_commit();
            }
            tmem_wait_ld();
            const float* xb = ring + (n % STAGES) * STAGE_F + g8 * REC_RS + t4;
#pragma unroll
            for (int kk = 0; kk < REC_CK / 8; ++kk) {
                uint32_t bh[NT][2], bl[NT][2];
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    split_tf32(xb[nt * 8 * REC_RS + kk * 8], bh[nt][0], bl[nt][0]);
                    split_tf32(xb[nt * 8 * REC_RS + kk * 8 + 4], bh[nt][1], bl[nt][1]);
                }
#pragma unroll
                for (int m = 0; m < NG; ++m) {
                    uint32_t ah[4], al[4];
#pragma unroll
                    for (int r = 0; r < 4; ++r) split_tf32(__uint_as_float(w[m][kk][r]), ah[r], al[r]);
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt) mma_tf32(c[m][nt], al, bh[nt]);
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt) mma_tf32(c[m][nt], ah, bl[nt]);
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt) mma_tf32(c[m][nt], ah, bh[nt]);
                }
            }
        }
        cp_async_wait<0>();
    }
    __syncthreads();                             // every warp's ring is dead: reuse the memory for the reduction
    float* red = smem;                           // red[warp][m][n][reg][lane]
#pragma unroll
    for (int m = 0; m < NG; ++m)
#pragma unroll
        for (int n = 0; n < NT; ++n)
#pragma unroll
            for (int r = 0; r < 4; ++r) red[(((warp * NG + m) * NT + n) * 4 + r) * 32 + lane] = c[m][n][r];
    __syncthreads();
#pragma unroll
    for (int p = 0; p < (NT + 1) / 2; ++p) {
        const int u = tid & 15, row = (tid >> 4) + 16 * p;
        const int n = row >> 3, col = row & 7;
        const int l = (u & 7) * 4 + (col >> 1), r = (u >> 3) * 2 + (col & 1);
#pragma unroll
        for (int m = 0; m < NG; ++m) {
            float s = 0.0f;
            if (n < NT) {
#pragma unroll
                for (int w8 = 0; w8 < REC_WARPS; ++w8) s += red[(((w8 * NG + m) * NT + n) * 4 + r) * 32 + l];
            }
            out[m][p] = s;
        }
    }
    __syncthreads();                             // smem may be reused by the caller right away
}

// floats of shared memory tile_accumulate_res32<NG, NT> needs (activation ring or reduction buffer, whichever is larger)
__host__ __device__ constexpr int res32_smem_floats(int NG, int NT) {
    return REC_WARPS * RES_STAGES * 8 * NT * REC_RS > REC_WARPS * NG * NT * 4 * 32 ? REC_WARPS * RES_STAGES * 8 * NT * REC_RS
                                                                                  : REC_WARPS * NG * NT * 4 * 32;
}

}  // namespace tg
