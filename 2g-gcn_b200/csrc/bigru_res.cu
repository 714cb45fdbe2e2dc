// Frame-level bidirectional GRUs with SMEM-RESIDENT recurrent weights (K-C, second design).
//
// The streaming kernel (bigru.cu) re-reads the 18.9 MB of W_hh from L2 on every one of the T steps.  W_hh of all three
// groups and both directions fits in the shared memory of the chip: each CTA owns up to three blocks of 8 hidden units
// (x 3 gates = 72 weight rows of D floats, 145 KB at D = 512) of ONE (group, direction), loads them once, and keeps them
// for all T steps.  Per step a CTA
//   1. stages the previous hidden state of its group's rows (<= 32 rows x D floats) from L2 into shared memory (cp.async),
//   2. multiplies on the tensor cores from shared memory only: mma.sync m16n8k16 with the 3xFP16 split (recurrent_res.cuh:
//      same accuracy class as 3xTF32 at half the tensor-pipe instructions), A = state rows split in registers, B = resident
//      weight rows stored PRE-SPLIT as (hi, lo) f16x2 word pairs — the same bytes as the fp32 values they replace —, K split
//      over the 8 warps in k16 steps (each warp: all 2 x 9 accumulator tiles),
//   3. reduces the 8 partial accumulators through shared memory (the staging buffer is reused), applies the GRU gate math
//      for its 24 units and publishes the new state (= the output row) through L2,
// followed by one grid barrier.  Same rounding class and outputs as bigru_kernel (parity tests unchanged).
#include <stdlib.h>
#include "recurrent.cuh"
#include "recurrent_res.cuh"
#include "bigru.h"

namespace tg {

namespace {

constexpr int BR_UB = 8;          // hidden units per block (one n8 MMA tile per gate)
// unit blocks per CTA = template parameter NBLK: 3 (22 CTAs per (group, direction): the whole stage on 132 CTAs) or 2 (32 CTAs: the
// lone recurrence of the hybrid launch, on the SMs the clusters leave free)
constexpr int BR_ROWS = 32;       // state rows per pass (two m16 tiles)

__device__ __forceinline__ int br_ld(int D) { return D + 8; }    // padded row stride (words): 64-bit fragment loads of a half-warp hit 16 distinct 8-byte slots

template <int BR_NBLK>
__global__ void __launch_bounds__(REC_THREADS, 1) bigru_res_kernel(const BiGruParams P, int ctas_per_gd) {
    constexpr int BR_NT = 3 * BR_NBLK;       // n8 tiles per CTA: [gate][block]
    constexpr int BR_ACC = 2 * BR_NT * 4;    // accumulator floats per lane
    extern __shared__ __align__(16) float smem[];
    __shared__ int s_fail;
    __shared__ long long rowoff[BR_ROWS];                // element offset of (video, entity) of each state row at t = 0 (first row block)
    const int D = P.D, T = P.T, LD = br_ld(D);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g8 = lane >> 2, t4 = lane & 3;
    float* wsm = smem;                                   // [BR_NT * 8][LD]   row (nt*8 + n): gate = nt / NBLK, block = nt % NBLK
    float* hsm = wsm + BR_NT * 8 * LD;                   // [BR_ROWS][LD] staged previous state; reused as the reduction buffer
    // CTA -> (group, direction, row-block group, unit-block slice).  A group with several row blocks is spread over n_rb CTAs per
    // weight slice when the grid has room (each keeps its own copy of the slice and walks every n_rb-th row block)
    int group = 0;
#pragma unroll 1
    for (int i = 1; i < P.ngroups; ++i)
        if ((int)blockIdx.x >= P.g[i].tile_begin) group = i;
    const BiGruGroup& G = P.g[group];
    const int RG = G.n_rb;
    const int local = blockIdx.x - G.tile_begin;
    const int dslot = local / (RG * ctas_per_gd);                     // slot among the directions this launch owns (skip_dirs)
    const int dir = G.skip_dirs == 1 ? 1 : (G.skip_dirs == 2 ? 0 : dslot);
    const int rg = (local - dslot * RG * ctas_per_gd) / ctas_per_gd, c = local % ctas_per_gd;
    const int rbase = rg * BR_ROWS;
    const int nblk = D / BR_UB;
    const int j0 = c * BR_NBLK;
    const int nb = min(BR_NBLK, nblk - j0);              // blocks this CTA owns (>= 1 by construction)
    if (tid == 0) s_fail = 0;
    const bool single_rb = rbase + RG * BR_ROWS >= G.rows;   // this CTA owns one row block: every step sees the same rows -> offsets precomputed
    if (tid < BR_ROWS) {
        const int r = rbase + tid < G.rows ? rbase + tid : 0, b = r / G.E, e = r - b * G.E;
        rowoff[tid] = ((long long)b * T * G.E + e) * 2 * D + dir * D;
    }

    // ---- load the resident weight rows once, pre-split: word pair kp of a row = (hi, lo) f16x2 of columns 2kp, 2kp+1 -------
    {
        const float* W = G.whh[dir];
        const int d2 = D / 2;
        float wmax = 0.0f;
        for (int i = tid; i < BR_NT * 8 * d2; i += REC_THREADS) {
            const int row = i / d2, kp = i - row * d2;
            const int nt = row >> 3, n = row & 7, gate = nt / BR_NBLK, blk = nt - gate * BR_NBLK;
            float2 v = make_float2(0.f, 0.f);
            if (blk < nb) v = __ldg(reinterpret_cast<const float2*>(W + (size_t)(gate * D + (j0 + blk) * BR_UB + n) * D) + kp);
            wmax = fmaxf(wmax, fmaxf(fabsf(v.x), fabsf(v.y)));
            uint2 hl;
            split_f16x2(v.x * RES_WSCALE, v.y * RES_WSCALE, hl.x, hl.y);
            *reinterpret_cast<uint2*>(wsm + row * LD + kp * 2) = hl;
        }
        if (!(wmax * RES_WSCALE < RES_F16_MAX)) atomicOr(P.sync.error, 2u);
    }
    // epilogue constants: this thread's outputs o = tid + 256*q  ->  (m tile, block, accumulator register, lane)
    float bh[BR_NBLK][3];
    int o_row[BR_NBLK], o_unit[BR_NBLK], o_base[BR_NBLK];
    bool o_ok[BR_NBLK];
#pragma unroll
    for (int q = 0; q < BR_NBLK; ++q) {
        const int o = tid + REC_THREADS * q;
        const int ol = o & 31, reg = (o >> 5) & 3, blk = (o >> 7) % BR_NBLK, m = o / (128 * BR_NBLK);
        o_row[q] = m * 16 + (ol >> 2) + ((reg & 2) ? 8 : 0);
        o_unit[q] = (j0 + blk) * BR_UB + 2 * (ol & 3) + (reg & 1);
        o_base[q] = ((m * BR_NT + blk) * 4 + reg) * 32 + ol;        // + gate * NBLK * 4 * 32 per gate, + warp * BR_ACC * 32 per partial
        o_ok[q] = m < 2 && blk < nb;
#pragma unroll
        for (int gt = 0; gt < 3; ++gt) bh[q][gt] = o_ok[q] ? __ldg(G.bhh[dir] + gt * D + o_unit[q]) : 0.0f;
    }
    long long o_fe0[BR_NBLK];                                  // (video, entity) part of the frame-entity index at t = 0 (first row block)
#pragma unroll
    for (int q = 0; q < BR_NBLK; ++q) {
        const int r = rbase + o_row[q] < G.rows ? rbase + o_row[q] : 0, b = r / G.E, e = r - b * G.E;
        o_fe0[q] = (long long)b * T * G.E + e;
    }
    __syncthreads();

    unsigned int epoch = 0;
    bool ok = true;
    // epilogue operands of one (step, row block): input pre-activations, previous state, output offsets
    float xg[BR_NBLK][3], hprev[BR_NBLK];
    size_t orow[BR_NBLK];
    float* gsave[BR_NBLK];
    bool valid[BR_NBLK];
    auto fetch = [&](int s, int rb0, int nrows) {
        const int t = dir == 0 ? s : T - 1 - s;
        const int tprev = dir == 0 ? t - 1 : t + 1;
#pragma unroll
        for (int q = 0; q < BR_NBLK; ++q) {
            valid[q] = o_ok[q] && o_row[q] < nrows;
            xg[q][0] = xg[q][1] = xg[q][2] = 0.0f;
            if (!single_rb || s == 0) hprev[q] = 0.0f;           // single row block: hprev is carried in the register (own last output)
            orow[q] = 0;
            gsave[q] = nullptr;
            if (valid[q]) {
                const int unit = o_unit[q];
                size_t fe0;
                if (single_rb) {
                    fe0 = (size_t)o_fe0[q];
                } else {
                    const int r = rb0 + o_row[q], b = r / G.E, e = r - b * G.E;
                    fe0 = (size_t)b * T * G.E + e;
                }
                const size_t fe = fe0 + (size_t)t * G.E;
                const float* gi = G.gi + (fe * 2 + dir) * 3 * D;
                xg[q][0] = __ldg(gi + unit); xg[q][1] = __ldg(gi + D + unit); xg[q][2] = __ldg(gi + 2 * D + unit);
                if (!single_rb && s > 0) hprev[q] = ld_cg(G.hfr + (fe0 + (size_t)tprev * G.E) * 2 * D + dir * D + unit);
                orow[q] = fe * 2 * D + dir * D + unit;
                if (G.gates != nullptr) gsave[q] = G.gates + (fe * 2 + dir) * 4 * D + unit;
            }
        }
    };
    const int own_rows = min(BR_ROWS, G.rows - rbase);      // rows of the CTA's first row block (> 0 by construction)
    if (single_rb) fetch(0, rbase, own_rows);
    for (int s = 0; s < T && ok; ++s) {
        const int t = dir == 0 ? s : T - 1 - s;
        const int tprev = dir == 0 ? t - 1 : t + 1;
        for (int rb0 = rbase; rb0 < G.rows; rb0 += RG * BR_ROWS) {
            const int nrows = min(BR_ROWS, G.rows - rb0);
            // ---- 1. stage the previous state of these rows: every warp copies exactly the K columns it multiplies (its k16
            //         steps ks = warp, warp + 8, ...), so only a __syncwarp separates the copies from the MMAs -------------
            if (s > 0) {
                const size_t tstep = (size_t)tprev * G.E * 2 * D;
                const int quarter = lane & 3;
#pragma unroll
                for (int p = 0; p < BR_ROWS / 8; ++p) {
                    const int row = (lane >> 2) + 8 * p;
                    const bool rv = row < nrows;
                    const float* base;
                    if (single_rb) {
                        base = G.hfr + rowoff[row] + tstep + quarter * 4;
                    } else {
                        const int r = rb0 + row;
                        const int b = rv ? r / G.E : 0, e = rv ? r - b * G.E : 0;
                        base = G.hfr + ((size_t)(b * T + tprev) * G.E + e) * 2 * D + dir * D + quarter * 4;
                    }
                    float* dstp = hsm + row * LD + quarter * 4;
                    for (int ks = warp; ks < D / 16; ks += REC_WARPS) cp_async16_zfill(dstp + ks * 16, rv ? base + ks * 16 : G.hfr, rv);
                }
            }
            cp_async_commit();
            if (!single_rb) fetch(s, rb0, nrows);
            cp_async_wait<0>();
            __syncwarp();
            float sum[BR_NBLK][3];
#pragma unroll
            for (int q = 0; q < BR_NBLK; ++q) sum[q][0] = sum[q][1] = sum[q][2] = 0.0f;
            if (s > 0) {
                // ---- 2. W_hh h on the tensor cores, operands from shared memory only --------------------------------
                float acc[2][BR_NT][4];
#pragma unroll
                for (int m = 0; m < 2; ++m)
#pragma unroll
                    for (int nt = 0; nt < BR_NT; ++nt)
#pragma unroll
                        for (int r = 0; r < 4; ++r) acc[m][nt][r] = 0.0f;
                const bool two = nrows > 16;             // second m16 tile holds rows
#pragma unroll 1
                for (int ks = warp; ks < D / 16; ks += REC_WARPS) {
                    const int k0 = ks * 16;
                    // A fragments (state rows): register 0 = (row g8, k0 + 2 t4 + {0,1}), 1 = row + 8, 2 = k + 8, 3 = both
                    uint32_t ah[2][4], al[2][4];
#pragma unroll
                    for (int m = 0; m < 2; ++m) {
                        if (m == 1 && !two) continue;
                        const float* hp = hsm + (m * 16 + g8) * LD + k0 + 2 * t4;
                        const float2 x0 = *reinterpret_cast<const float2*>(hp);
                        const float2 x1 = *reinterpret_cast<const float2*>(hp + 8 * LD);
                        const float2 x2 = *reinterpret_cast<const float2*>(hp + 8);
                        const float2 x3 = *reinterpret_cast<const float2*>(hp + 8 * LD + 8);
                        split_f16x2(x0.x, x0.y, ah[m][0], al[m][0]);
                        split_f16x2(x1.x, x1.y, ah[m][1], al[m][1]);
                        split_f16x2(x2.x, x2.y, ah[m][2], al[m][2]);
                        split_f16x2(x3.x, x3.y, ah[m][3], al[m][3]);
                    }
#pragma unroll
                    for (int gt = 0; gt < 3; ++gt) {     // one gate = BR_NBLK n-tiles: 2*NBLK independent accumulators per pass
                        uint32_t bhi[BR_NBLK][2], blo[BR_NBLK][2];
#pragma unroll
                        for (int j = 0; j < BR_NBLK; ++j) {
                            const float* wp = wsm + ((gt * BR_NBLK + j) * 8 + g8) * LD + k0 + 2 * t4;
                            const uint2 w0 = *reinterpret_cast<const uint2*>(wp);
                            const uint2 w1 = *reinterpret_cast<const uint2*>(wp + 8);
                            bhi[j][0] = w0.x; blo[j][0] = w0.y;
                            bhi[j][1] = w1.x; blo[j][1] = w1.y;
                        }
#pragma unroll
                        for (int j = 0; j < BR_NBLK; ++j) {
                            mma_f16(acc[0][gt * BR_NBLK + j], al[0], bhi[j]);
                            if (two) mma_f16(acc[1][gt * BR_NBLK + j], al[1], bhi[j]);
                        }
#pragma unroll
                        for (int j = 0; j < BR_NBLK; ++j) {
                            mma_f16(acc[0][gt * BR_NBLK + j], ah[0], blo[j]);
                            if (two) mma_f16(acc[1][gt * BR_NBLK + j], ah[1], blo[j]);
                        }
#pragma unroll
                        for (int j = 0; j < BR_NBLK; ++j) {
                            mma_f16(acc[0][gt * BR_NBLK + j], ah[0], bhi[j]);
                            if (two) mma_f16(acc[1][gt * BR_NBLK + j], ah[1], bhi[j]);
                        }
                    }
                }
                __syncthreads();                         // every warp is done reading the staged state
                // ---- 3. reduce the 8 K-slices through shared memory ------------------------------------------------
                float* red = hsm;
#pragma unroll
                for (int m = 0; m < 2; ++m)
#pragma unroll
                    for (int nt = 0; nt < BR_NT; ++nt)
#pragma unroll
                        for (int r = 0; r < 4; ++r) red[(warp * BR_ACC + (m * BR_NT + nt) * 4 + r) * 32 + lane] = acc[m][nt][r];
                __syncthreads();
#pragma unroll
                for (int q = 0; q < BR_NBLK; ++q) {
                    if (!o_ok[q]) continue;
#pragma unroll
                    for (int gt = 0; gt < 3; ++gt) {
                        float v = 0.0f;
#pragma unroll
                        for (int w = 0; w < REC_WARPS; ++w) v += red[w * BR_ACC * 32 + gt * BR_NBLK * 4 * 32 + o_base[q]];
                        sum[q][gt] = v * (1.0f / RES_WSCALE);
                    }
                }
            }
#pragma unroll
            for (int q = 0; q < BR_NBLK; ++q)
                if (valid[q]) {
                    const float hnew = gru_update(xg[q][0], xg[q][1], xg[q][2], sum[q][0] + bh[q][0], sum[q][1] + bh[q][1],
                                                  sum[q][2] + bh[q][2], hprev[q], gsave[q], D);
                    G.hfr[orow[q]] = hnew;
                    hprev[q] = hnew;
                }
            if (!single_rb) __syncthreads();             // the reduction buffer becomes the staging buffer again
        }
        if (s + 1 < T) {
            grid_arrive(P.sync, epoch);                  // (its __syncthreads also retires the reduction buffer)
            if (single_rb) fetch(s + 1, rbase, own_rows);      // next step's input pre-activations arrive in the shadow of the barrier
            if (!grid_wait(P.sync, epoch, gridDim.x, &s_fail)) ok = false;
        }
    }
}

}  // namespace

// Returns 0 when the resident kernel was launched, -1 when this shape does not qualify (caller falls back), > 0 on error.
int launch_bigru_resident(BiGruParams& P, cudaStream_t stream, int nblk) {
    TG_REQUIRE(nblk == 2 || nblk == 3, "bigru (resident): %d unit blocks per CTA not built", nblk);
    const int BR_NBLK = nblk, BR_NT = 3 * nblk, BR_ACC = 2 * BR_NT * 4;
    static int enabled = -1;
    if (enabled < 0) {
        const char* e = getenv("TGGCN_BIGRU_RES");
        enabled = (e != nullptr && e[0] == '0') ? 0 : 1;
    }
    const int D = P.D;
    if (!enabled || P.no_fp16_split || D % 64 != 0) return -1;
    const int LD = D + 8;
    const size_t red_floats = (size_t)REC_WARPS * BR_ACC * 32, stage_floats = (size_t)BR_ROWS * LD;
    const size_t smem = sizeof(float) * ((size_t)BR_NT * 8 * LD + (red_floats > stage_floats ? red_floats : stage_floats));
    if (smem > 227 * 1024) return -1;
    const int ctas_per_gd = cdiv(D / BR_UB, BR_NBLK);
    // row-block groups per (group, direction): as many as the row blocks while the grid fits on the GPU
    int rgs[3] = {1, 1, 1};
    auto ndirs = [&](int i) { return 2 - (P.g[i].skip_dirs & 1) - ((P.g[i].skip_dirs >> 1) & 1); };
    auto total = [&]() { int t = 0; for (int i = 0; i < P.ngroups; ++i) t += ndirs(i) * rgs[i] * ctas_per_gd; return t; };
    for (bool grown = true; grown;) {
        grown = false;
        for (int i = 0; i < P.ngroups; ++i) {
            if (ndirs(i) > 0 && rgs[i] < cdiv(P.g[i].rows, BR_ROWS)) {
                ++rgs[i];
                if (total() <= num_sms()) grown = true; else --rgs[i];
            }
        }
    }
    int grid = 0;
    for (int i = 0; i < P.ngroups; ++i) {
        P.g[i].n_rb = rgs[i];
        P.g[i].tile_begin = grid;
        grid += ndirs(i) * rgs[i] * ctas_per_gd;
    }
    auto kern = nblk == 2 ? bigru_res_kernel<2> : bigru_res_kernel<3>;
    if (int rc = ensure_smem((const void*)kern, smem)) return rc;
    int per_sm = 0;
    TG_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, REC_THREADS, smem));
    if (per_sm < 1 || grid > per_sm * num_sms()) return -1;
    TG_CUDA_OK(cudaMemsetAsync(P.sync.counter, 0, sizeof(unsigned int), stream));      // the error word belongs to the caller (tggcn_forward zeroes it once)
    int cpg = ctas_per_gd;
    void* args[] = {(void*)&P, (void*)&cpg};
    TG_CUDA_OK(cudaLaunchCooperativeKernel((const void*)kern, dim3(grid), dim3(REC_THREADS), args, smem, stream));
    ++g_launches;
    return 0;
}

}  // namespace tg
