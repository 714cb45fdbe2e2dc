// Building block of the two recurrent kernels (frame-level BiGRU, segment-level gated GRUCell graph):
// a CTA-wide "gate tile"
//     out[g][r][u] = sum_k Wrow[g*16 + u][k] * X[r][k]          g < NG groups, r < 8*NT rows, u < 16 units
// over a K range made of up to two segments (e.g. the segment-message columns of W_ih against the aggregated
// messages, then W_hh against the previous state), each with its own pointer tables.  Weight rows and
// activation rows are addressed through pointer tables (a null pointer is an all-zero row), so the same routine
// serves W_hh h, the segment columns of W_ih, the message MLPs and — with the receivers' state rows as an extra
// "weight" group — the attention logits of the message tiles.
//
// How it got here (ncu captures under profiles/, round 1): three FFMA versions (CTA-wide K-chunk barrier;
// K split across warps; warp-private pipelines with 512 threads) all stalled at IPC ~0.4 — first on
// short_scoreboard/barrier, finally on mio_throttle: a 16-row tile reuses every weight only 16 times, so an
// FFMA formulation issues one LDS.128 per ~9 FFMA and the shared-memory instruction queue, not the FMA pipe,
// is the limit.  This version therefore runs the products on the tensor cores:
//  * mma.sync.m16n8k8 TF32 with the 3xTF32 split (a = hi + lo, hi*hi + hi*lo + lo*hi, fp32 accumulate), which
//    keeps fp32-class accuracy (relative error ~2^-21 per product) at 1/3 of the instruction count of FFMA and
//    a third of the shared-memory instructions;
//  * the K range is split across the 8 warps in 16-float chunks (round-robin) and EVERY WARP RUNS ITS OWN
//    cp.async PIPELINE in a private shared-memory ring: no block-wide barrier inside the K loop, only __syncwarp;
//  * ring rows are padded to 20 floats so that the 8 rows x 4 columns of a fragment load hit 32 distinct banks;
//  * copies are branch-free (cp.async zero-fill form; .cg because activations were written by other CTAs of the
//    same persistent kernel and L1 must be bypassed);
//  * partial accumulators of the 8 warps are reduced through shared memory once per tile.
#pragma once
#include "common.cuh"

namespace tg {

constexpr int REC_THREADS = 256;
constexpr int REC_WARPS = REC_THREADS / 32;
constexpr int REC_J = 16;         // units per tile  (one m16 MMA tile per group)
constexpr int REC_CK = 16;        // floats of K per chunk (two k8 MMA steps, 64 bytes per row)
constexpr int REC_RS = 20;        // ring row stride in floats (16 data + 4 pad)

__host__ __device__ constexpr int tile_ring_floats(int NG, int NT, int STAGES) {
    return REC_WARPS * STAGES * (NG * REC_J + 8 * NT) * REC_RS;
}
__host__ __device__ constexpr int tile_red_floats(int NG, int NT) { return REC_WARPS * NG * NT * 32 * 4; }
__host__ __device__ constexpr int tile_smem_floats(int NG, int NT, int STAGES) {
    return tile_ring_floats(NG, NT, STAGES) > tile_red_floats(NG, NT) ? tile_ring_floats(NG, NT, STAGES)
                                                                       : tile_red_floats(NG, NT);
}

// 16-byte async copy with zero-fill: copies `valid ? 16 : 0` bytes and zero-fills the rest.
__device__ __forceinline__ void cp_async16_zfill(void* smem, const void* gmem, bool valid) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    const int n = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(n) : "memory");
}

// D += A(16x8, row) * B(8x8, col), TF32 operands, fp32 accumulate.
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// x = hi + lo with hi exactly representable in TF32 (low 13 mantissa bits cleared) and lo = x - hi exact in fp32.
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
    hi = __float_as_uint(x) & 0xffffe000u;
    lo = __float_as_uint(x - __uint_as_float(hi));
}

// Pointer tables in shared memory: rows [0, NG*16) are weight rows, rows [NG*16, NG*16 + 8*NT) activation rows.
// tab1 covers K columns [0, K1), tab2 columns [K1, K1 + K2) (row pointers are to the START of the segment).
// K1, K2 multiples of 16; rows 16-byte aligned.  A null pointer is an all-zero row for that segment.
// skip1 / skip2: bit m set = group m is all-zero in segment 1 / 2 (its MMAs are skipped).
// gdummy: any valid 16-byte aligned global address (source operand of the zero-size copies).
// On return out[g][p] holds the full sum for this thread's epilogue pair p: unit = tid % 16, row = tid / 16 + 16*p.
template <int NG, int NT, int STAGES>
__device__ __forceinline__ void tile_accumulate(float (&out)[NG][(NT + 1) / 2], const float* const* tab1,
                                                const float* const* tab2, int K1, int K2, unsigned skip1, unsigned skip2,
                                                const float* gdummy, float* smem) {
    constexpr int WR = NG * REC_J, ROWS = WR + 8 * NT;
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
        // this lane's copy pieces: (row, quarter) pairs; per segment a source pointer that only moves along K
        const float* s1[NP];
        const float* s2[NP];
        int dst[NP];
        bool in[NP];
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            const int piece = lane + p * 32;
            const int row = piece >> 2, quarter = piece & 3;
            in[p] = piece < ROWS * 4;
            const float* b1 = in[p] ? tab1[row] : nullptr;
            const float* b2 = (in[p] && K2 > 0) ? tab2[row] : nullptr;
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
                cp_async16_zfill(ring + st * STAGE_F + dst[p], ok ? base + koff : gdummy, ok);   // !ok: size 0, zero-fill
            }
        };
#pragma unroll
        for (int st = 0; st < STAGES - 1; ++st) {
            if (st < nmine) issue(st, st);
            cp_async_commit();
        }
#pragma unroll 1
        for (int n = 0; n < nmine; ++n) {
            cp_async_wait<STAGES - 2>();         // this lane's pieces of chunk n have landed
            __syncwarp();                        // ... all lanes'; and all lanes are done reading chunk n-1
            {
                const int nn = n + STAGES - 1;
                if (nn < nmine) issue(nn, nn % STAGES);
                cp_async_commit();
            }
            const unsigned skip = (warp + n * REC_WARPS) < chunks1 ? skip1 : skip2;
            const float* wb = ring + (n % STAGES) * STAGE_F + g8 * REC_RS + t4;
            const float* xb = ring + (n % STAGES) * STAGE_F + (WR + g8) * REC_RS + t4;
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
                    if ((skip >> m) & 1u) continue;              // warp-uniform
                    uint32_t ah[4], al[4];
                    const float* wm = wb + m * REC_J * REC_RS + kk * 8;
                    split_tf32(wm[0], ah[0], al[0]);
                    split_tf32(wm[8 * REC_RS], ah[1], al[1]);
                    split_tf32(wm[4], ah[2], al[2]);
                    split_tf32(wm[8 * REC_RS + 4], ah[3], al[3]);
                    // three passes over the NT independent accumulator tiles (small terms first)
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
    // cross-warp reduction of the K split: red[warp][m][n][reg][lane]
    float* red = smem;
#pragma unroll
    for (int m = 0; m < NG; ++m)
#pragma unroll
        for (int n = 0; n < NT; ++n)
#pragma unroll
            for (int r = 0; r < 4; ++r) red[(((warp * NG + m) * NT + n) * 4 + r) * 32 + lane] = c[m][n][r];
    __syncthreads();
#pragma unroll
    for (int p = 0; p < (NT + 1) / 2; ++p) {
        // accumulator element (unit u, row) of group m lives in: n = row / 8, lane = (u % 8) * 4 + (row % 8) / 2,
        // reg = (u / 8) * 2 + (row % 2)        (m16n8 C fragment layout)
        const int u = tid & 15, row = (tid >> 4) + 16 * p;
        const int n = row >> 3, col = row & 7;
        const int l = (u & 7) * 4 + (col >> 1), r = (u >> 3) * 2 + (col & 1);
#pragma unroll
        for (int m = 0; m < NG; ++m) {
            float s = 0.0f;
            if (n < NT) {
#pragma unroll
                for (int w = 0; w < REC_WARPS; ++w) s += red[(((w * NG + m) * NT + n) * 4 + r) * 32 + l];
            }
            out[m][p] = s;
        }
    }
    __syncthreads();                             // smem may be reused by the caller right away
}

// GRU cell update, gate order (r, z, n) as torch.nn.GRU / GRUCell (vhoi/models.py:267,:294):
//   r = s(xr + hr), z = s(xz + hz), n = tanh(xn + r*hn), h' = n + z*(h - n)
// When `gates` is not null the values the backward needs are saved: gates[0]=r, [stride]=z, [2*stride]=n, [3*stride]=hn.
__device__ __forceinline__ float gru_update(float xr, float xz, float xn, float hr, float hz, float hn, float hprev,
                                            float* gates = nullptr, int stride = 0) {
    const float r = sigmoidf_acc(xr + hr);
    const float z = sigmoidf_acc(xz + hz);
    const float n = tanhf(xn + r * hn);
    if (gates != nullptr) { gates[0] = r; gates[stride] = z; gates[2 * stride] = n; gates[3 * stride] = hn; }
    return n + z * (hprev - n);
}

}  // namespace tg
