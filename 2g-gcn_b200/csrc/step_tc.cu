// Large-batch recurrent path (dims.recurrent_mode 2, or chosen automatically from the rows per step): every recurrent step of the
// frame-level BiGRUs (vhoi/models.py:983-1002) and of the segment-level graph (:785-880) is ONE dense contraction
//     gates[rows, 3D] = act_rows[rows, K] * W[3D, K]^T            rows = B * entities (hundreds to thousands)
// on the 5th-generation tensor cores, with the GRU cell fused into the epilogue.
//
// step_tc_kernel, one (128 * MT)-row x 64-unit tile per CTA (MT = 2: two accumulators share every weight tile — the kernel runs at
// the L2 -> SM bandwidth, and a 256-row tile moves 30 % fewer operand bytes per FLOP), 320 threads, warp-specialised:
//   warp 0      TMA producer: one elected lane streams the operand tiles with cp.async.bulk.tensor (3-D tensor maps: K x rows x
//               plane) straight into the K-major SWIZZLE_128B layout the UMMA descriptors read — no register pass, no split at
//               run time: activations are WRITTEN as fp16 (hi, lo) pairs by the previous step's epilogue, weights are split
//               once per forward by pack16_kernel.  mbarrier expect_tx / complete_tx hands a stage to the MMA warp.
//   warp 1      MMA issuer: one lane issues tcgen05.mma kind::f16, M = 128 activation rows, N = 192 (three gates of 64 units),
//               K = 16, accumulators in tensor memory.  fp32-class accuracy from the 3-term split x*w = hi*hi + hi*lo + lo*hi
//               (22 mantissa bits; weights pre-scaled by 2^8 so their lo parts stay normal), or plain bf16 (dims.precision 1).
//               TMEM columns [n_i | r | z | n_h]: the input-part k-blocks (aggregated segment messages) accumulate into
//               [n_i r z], the hidden-part k-blocks into [r z n_h] — a GRU needs the two n contributions apart.
//   warps 2-9   epilogue: tcgen05.ld (lane = activation row), bias + hoisted input pre-activations, sigmoid / tanh, blend with
//               the hard segmentation gate, fp32 state / output row, the next step's fp16 (hi, lo) operand row, and — in
//               training — the gate values the BPTT kernels need.
// The same kernel runs the segment-level message MLPs (ReLU epilogue, N = 128).  seg_attend_kernel turns their outputs into the
// aggregated messages (scaled-dot-product attention over the previous states, vhoi/models.py:1051-1381) between the two GEMMs.
#include <stdlib.h>
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>

#include "step_tc.cuh"
#include "tcgen05.cuh"

namespace tg {

constexpr int ST_BM = 128;                  // activation rows per tile (UMMA M)
constexpr int ST_BK = 64;                   // K elements per k-block: 128 bytes of fp16 = one swizzle row
constexpr int ST_U = 64;                    // units per GRU tile (UMMA N = 3 * 64)
constexpr int ST_BN_RELU = 128;             // output columns per tile of the message MLPs
constexpr int ST_EPI_WARPS = 8;
constexpr int ST_THREADS = (2 + ST_EPI_WARPS) * 32;
constexpr int ST_MAX_PROBLEMS = 8;
constexpr int ST_MAX_MAPS = 14;
constexpr int ST_ACC_COLS = 256;            // TMEM columns of one 128-row accumulator: [n_i | r | z | n_h] x 64 units
constexpr int ST_A_TILE = ST_BM * 128;      // bytes of one 128-row operand tile
constexpr int ST_B_TILE = 3 * ST_U * 128;   // bytes of one weight tile (three gate boxes, or one 128-row box + slack)
constexpr float ST_W_SCALE = 256.0f;        // fp16 split: weights are stored times 2^8

// CG: CTAs per tile (tcgen05 cta_group).  CG = 2: a CTA pair on the two SMs of a TPC computes a 256-row tile — each CTA
// stages its own 128 activation rows and HALF of the weight rows, one tcgen05.mma.cta_group::2 reads both halves — so every SM
// moves 30 % fewer bytes through its shared memory per FLOP, which is what bounds this kernel (profiles/r02_step_tile_height.txt).
template <int PREC, int MT, int CG> struct StCfg {
    static_assert(CG == 1 || MT == 1, "256-row tiles are either two accumulators in one CTA or one accumulator in each CTA of a pair");
    static constexpr int PLANES = PREC == 0 ? 2 : 1;                       // (hi, lo) or a single bf16 plane
    static constexpr int A_PLANE = MT * ST_A_TILE;                         // one plane of this CTA's activation rows: 128 * MT rows
    static constexpr int A_BYTES = PLANES * A_PLANE;
    static constexpr int B_PLANE = ST_B_TILE / CG;                         // this CTA's share of the weight tile
    static constexpr int STAGE_BYTES = A_BYTES + PLANES * B_PLANE;         // fp16 split: 80 / 112 / 56 KB; bf16: half of that
    static constexpr int STAGES = (227 * 1024 - 2048) / STAGE_BYTES > 8 ? 8 : (227 * 1024 - 2048) / STAGE_BYTES;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024;
    static constexpr int TMEM_COLS = MT * ST_ACC_COLS;
    static constexpr int WBOX = ST_U / CG;                                 // rows of one weight TMA box (the maps are encoded with it)
};

enum { ST_GRU = 0, ST_RELU = 1 };

struct StepProblem {
    int mode;
    int rows, E;                        // activation rows (videos x entities), entities per video
    int m_tiles, n_tiles, tile_begin;
    int D;                              // layer width: gate row period of the weight matrices / output columns (RELU)
    int nkb1, nkb2;                     // k-blocks of the input part (aggregated messages) and of the hidden part
    int a1_map, a1_plane, a2_map, a2_plane;
    int b1_map, b1_plane, b2_map, b2_plane;      // weight maps: whole gate tile (CG = 1) or my gate of (r, z) (CG = 2) ...
    int b1n_map, b2n_map;                        // ... and my half of the n rows (CG = 2 only)
    int T, t, tprev, dir, first;
    const float* xg;                    // (B,T,E,2,3D) hoisted input pre-activations incl. b_ih
    const float* bhh;                   // (3D)
    const float* ugate;                 // (B,T,E) hard gates, or null (plain GRU)
    float* hx;                          // (B,T,E,2D) state = output rows [fwd | bwd]
    float* gsave;                       // (B,T,E,2,4D) r, z, n, hn for the backward, or null
    void* ring_out;                     // [rows][D] 16-bit hi plane of the state ring this step writes (lo plane follows)
    const float* bias;                  // RELU: (D)
    float* out;                         // RELU: row (b, e) -> out + b * out_bstride + e * D
    long long out_bstride;
};

struct StepLaunch {
    CUtensorMap maps[ST_MAX_MAPS];
    StepProblem p[ST_MAX_PROBLEMS];
    int count;
    float acc_scale;
};

// ---- PTX helpers (TMA, CTA pairs) ------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}\n" ::"r"(bar), "r"(bytes) : "memory");
}
// CG = 2: the copy lands in THIS CTA's shared memory and completes bytes on the LEADER's barrier (bar: shared::cluster address)
template <int CG> __device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
    if (CG == 1)
        asm volatile(
            "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n" ::"r"(dst),
            "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
            : "memory");
    else
        asm volatile(
            "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n" ::"r"(dst),
            "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
            : "memory");
}
template <int CG> __device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint32_t bar) {
    if (CG == 1)
        asm volatile(
            "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];\n" ::"r"(dst),
            "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar)
            : "memory");
    else
        asm volatile(
            "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];\n" ::"r"(dst),
            "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar)
            : "memory");
}
// Programmatic dependent launch: the next kernel of the stream may start its prologue (barriers, tensor-memory allocation, tensor
// map prefetch) while this grid drains; it blocks in pdl_wait until the previous grid has completed and its writes are visible.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;\n" ::: "memory"); }
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];\n" ::"l"(p)); }
// kind::f16 instruction descriptor: D = F32, A = B = F16 (0) or BF16 (1), both K-major, N >> 3 at bits 17-22, M >> 4 at 24-28
__device__ __forceinline__ uint32_t umma_idesc_16(int bf16, int M, int N) {
    return (1u << 4) | ((uint32_t)bf16 << 7) | ((uint32_t)bf16 << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
template <int CG> __device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    if (CG == 1)
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
            "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
            : "memory");
    else
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
            "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
            : "memory");
}
// arrive on the barrier at the same offset in every CTA of the pair once the MMAs issued so far have completed
template <int CG> __device__ __forceinline__ void umma_commit_cg(uint32_t bar) {
    if (CG == 1)
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(bar) : "memory");
    else
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n" ::"r"(bar),
                     "h"((uint16_t)3)
                     : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t cta_addr, uint32_t rank) {       // shared::cluster address of the same offset in CTA `rank`
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(r) : "r"(cta_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];\n" ::"r"(cluster_addr) : "memory");
}

template <int PREC> __device__ __forceinline__ void store16(void* hi_plane, size_t plane_elems, size_t off, const float (&v)[16]) {
    if (PREC == 0) {
        __align__(16) __half hi[16];
        __align__(16) __half lo[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            hi[j] = __float2half_rn(v[j]);
            lo[j] = __float2half_rn(v[j] - __half2float(hi[j]));
        }
        __half* ph = reinterpret_cast<__half*>(hi_plane) + off;
        __half* pl = ph + plane_elems;
        reinterpret_cast<uint4*>(ph)[0] = reinterpret_cast<const uint4*>(hi)[0];
        reinterpret_cast<uint4*>(ph)[1] = reinterpret_cast<const uint4*>(hi)[1];
        reinterpret_cast<uint4*>(pl)[0] = reinterpret_cast<const uint4*>(lo)[0];
        reinterpret_cast<uint4*>(pl)[1] = reinterpret_cast<const uint4*>(lo)[1];
    } else {
        __align__(16) __nv_bfloat16 hi[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) hi[j] = __float2bfloat16_rn(v[j]);
        __nv_bfloat16* ph = reinterpret_cast<__nv_bfloat16*>(hi_plane) + off;
        reinterpret_cast<uint4*>(ph)[0] = reinterpret_cast<const uint4*>(hi)[0];
        reinterpret_cast<uint4*>(ph)[1] = reinterpret_cast<const uint4*>(hi)[1];
    }
}

// Gate non-linearities of the epilogue on the fast exponential (ex2.approx: relative error ~2^-22, i.e. ~1e-7 absolute on values
// in [-1, 1]).  The epilogue is issue-bound — 128 rows x 64 units of (2 sigmoids + 1 tanh) on 8 warps; with expf / tanhf it
// took 26k of a cell tile's 79k cycles (clock64 trace, profiles/r02_step_trace.txt) — and only 2 warps per scheduler hide latency.
__device__ __forceinline__ float sigmoid_fast(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float tanh_fast(float x) {
    const float t = __expf(-2.0f * fabsf(x));                      // in (0, 1]: no overflow, no cancellation for large |x|
    return copysignf(__fdividef(1.0f - t, 1.0f + t), x);
}

__device__ __forceinline__ void load16(const float* p, float (&v)[16]) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float4 x = __ldg(reinterpret_cast<const float4*>(p) + j);
        v[4 * j] = x.x; v[4 * j + 1] = x.y; v[4 * j + 2] = x.z; v[4 * j + 3] = x.w;
    }
}
__device__ __forceinline__ void load16_cg(const float* p, float (&v)[16]) {     // data another kernel wrote moments ago: L2 path
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float4 x = ld_cg4(p + 4 * j);
        v[4 * j] = x.x; v[4 * j + 1] = x.y; v[4 * j + 2] = x.z; v[4 * j + 3] = x.w;
    }
}
__device__ __forceinline__ void store16f(float* p, const float (&v)[16]) {
#pragma unroll
    for (int j = 0; j < 4; ++j) reinterpret_cast<float4*>(p)[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
}

#ifdef ST_TRACE      // timing experiments only (TGGCN_NVCC_DEFS=-DST_TRACE): clock64 stamps of CTA 0, read back with tggcn_debug_trace
__device__ long long g_st_trace[64];
__device__ int g_st_trace_kind = 0;      // which launches record: 0 = cell GEMM (GRU with an input part), 1 = plain GRU (BiGRU), 2 = message MLPs
#define ST_STAMP(i) do { if (blockIdx.x == 0 && st_trace_on) g_st_trace[i] = clock64(); } while (0)
#else
#define ST_STAMP(i) do { } while (0)
#endif

// Accumulator columns of a GRU tile: [r (64) | z (64) | n_i (64) | n_h (64)].  Weight tile in shared memory: [r z] rows (one
// N = 128 operand) followed by the n rows (N = 64) — with CG = 2 the first CTA of the pair holds the r rows and the first half of
// the n rows, the second one the z rows and the other half, which is exactly how cta_group::2 splits an N-operand.
template <int PREC, int MT, int CG>
__global__ void __launch_bounds__(ST_THREADS, 1) step_tc_kernel(const __grid_constant__ StepLaunch L) {
    using Cfg = StCfg<PREC, MT, CG>;
    constexpr int STAGES = Cfg::STAGES, PLANES = Cfg::PLANES, WBOX = Cfg::WBOX;
    constexpr int RZ_BYTES = 2 * ST_U * 128 / CG, NB_BYTES = ST_U * 128 / CG;        // this CTA's rows of the two weight operands
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[2 * STAGES + 1];
    __shared__ uint32_t tmem_base_smem;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    pdl_launch_dependents();
#ifdef ST_TRACE
    const long long st_t0 = clock64();
    const int st_kind = L.p[0].mode == ST_RELU ? 2 : (L.p[0].nkb1 > 0 ? 0 : 1);
    const bool st_trace_on = st_kind == g_st_trace_kind;
    if (tid == 0 && blockIdx.x == 0 && st_trace_on) g_st_trace[0] = st_t0;
#endif
    uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t tiles_u32 = smem_u32(tiles);
    const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[STAGES]), tfull = smem_u32(&bars[2 * STAGES]);
    const int rank = CG == 2 ? (int)cluster_ctarank() : 0;
    const bool leader = rank == 0;

    const int vtile = CG == 2 ? (int)blockIdx.x >> 1 : (int)blockIdx.x;          // a CTA pair works on one tile
    int pi = 0;
#pragma unroll 1
    for (int i = 1; i < L.count; ++i)
        if (vtile >= L.p[i].tile_begin) pi = i;
    const StepProblem& P = L.p[pi];
    const int tile = vtile - P.tile_begin;
    const int mt = tile / P.n_tiles, nt = tile - mt * P.n_tiles;
    const int m0 = mt * ST_BM * MT * CG + rank * ST_BM;                          // first activation row of THIS CTA
    const int nkb = P.nkb1 + P.nkb2;
    const bool gru = P.mode == ST_GRU;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full0 + 8 * s, CG);              // leader's expect_tx arrive (+ the peer's plain arrive)
            mbar_init(empty0 + 8 * s, 1);              // one tcgen05.commit (multicast to both CTAs of a pair)
        }
        mbar_init(tfull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 2) {
        if (CG == 1) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_smem)), "r"((uint32_t)Cfg::TMEM_COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_smem)), "r"((uint32_t)Cfg::TMEM_COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n" ::: "memory");
        }
    }
    tc_fence_before();
    if (CG == 2) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;
    if (tid == 0) ST_STAMP(1);
    pdl_wait();                                   // everything above overlapped the previous kernel of this stream

    if (warp == 0) {
        // ------------------------------ TMA producer (every CTA) ------------------------------
        if (lane == 0) {
            const uint32_t stage_tx = (uint32_t)PLANES * (Cfg::A_PLANE + (gru ? RZ_BYTES + NB_BYTES : ST_BN_RELU * 128 / CG));
#pragma unroll 1
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % STAGES;
                mbar_wait_backoff(empty0 + 8 * s, ((kb / STAGES) & 1) ^ 1);
                if (kb < 28) ST_STAMP(4 + kb);
                const uint32_t my_full = full0 + 8 * s;
                const uint32_t bar = CG == 2 ? mapa_u32(my_full, 0) : my_full;          // bytes complete on the leader's barrier
                if (leader) mbar_expect_tx(my_full, stage_tx * CG);
                else        mbar_arrive_remote(bar);
                const bool seg2 = kb >= P.nkb1;
                const int k0 = (seg2 ? kb - P.nkb1 : kb) * ST_BK;
                const uint32_t st = tiles_u32 + s * Cfg::STAGE_BYTES;
                const CUtensorMap* am = &L.maps[seg2 ? P.a2_map : P.a1_map];
                const CUtensorMap* bm = &L.maps[seg2 ? P.b2_map : P.b1_map];
                const int ap = seg2 ? P.a2_plane : P.a1_plane, bp = seg2 ? P.b2_plane : P.b1_plane;
                // Few, large copies: a tensor-map copy costs ~200 cycles of TMA issue whatever its size (measured: 8 copies per
                // k-block gave a 2200-2460 cycle period for 56-112 KB stages, profiles/r02_step_tma_ops.txt), so the (hi, lo)
                // planes travel in one box (last box dimension = PLANES) and the three gate slices in one 4-D box.
                tma_load_3d<CG>(st, am, k0, m0, ap, bar);                                   // box {64 k, 128 * MT rows, PLANES}
                const uint32_t b = st + Cfg::A_BYTES;
                if (gru) {
                    const int u0 = nt * ST_U;
                    if (CG == 1) {
                        tma_load_4d<CG>(b, bm, k0, u0, 0, bp, bar);                         // box {64 k, 64 units, 3 gates, PLANES}: [hi: r z n][lo: r z n]
                    } else {
                        const CUtensorMap* bn = &L.maps[seg2 ? P.b2n_map : P.b1n_map];
                        tma_load_4d<CG>(b, bm, k0, u0, rank, bp, bar);                      // box {64, 64 units, 1 gate, PLANES}: my gate of (r, z)
                        tma_load_4d<CG>(b + PLANES * RZ_BYTES, bn, k0, u0 + rank * WBOX, 2, bp, bar);   // box {64, 32 units, 1, PLANES}: my half of n
                    }
                } else {
                    tma_load_3d<CG>(b, bm, k0, nt * ST_BN_RELU + rank * (ST_BN_RELU / CG), bp, bar);    // box {64 k, 128 / CG rows, PLANES}
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------ MMA issuer (leader CTA only) ------------------------------
        if (leader) {
            constexpr int MM = ST_BM * CG;
            const uint32_t id_rz = umma_idesc_16(PREC, MM, 2 * ST_U), id_n = umma_idesc_16(PREC, MM, ST_U), id_relu = umma_idesc_16(PREC, MM, ST_BN_RELU);
#pragma unroll 1
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % STAGES;
                mbar_wait(full0 + 8 * s, (kb / STAGES) & 1);
                tc_fence_after();
                if (lane == 0) {
                    if (kb < 28) ST_STAMP(32 + kb);
                    const uint32_t st = tiles_u32 + s * Cfg::STAGE_BYTES;
                    const uint32_t a_hi = st, a_lo = st + Cfg::A_PLANE;
                    const uint32_t b0 = st + Cfg::A_BYTES;
                    // weight tile: CG = 1 [hi: r z n][lo: r z n]; CG = 2 [rz_hi][rz_lo][n_hi][n_lo] (this CTA's rows); RELU [hi][lo]
                    const uint32_t rz_hi = b0, rz_lo = b0 + (gru ? (CG == 1 ? ST_B_TILE : RZ_BYTES) : ST_BN_RELU * 128 / CG);
                    const uint32_t n_hi = b0 + (CG == 1 ? 2 * ST_U * 128 : PLANES * RZ_BYTES), n_lo = n_hi + (CG == 1 ? ST_B_TILE : NB_BYTES);
                    const bool seg2 = kb >= P.nkb1;
                    const bool seg_first_kb = seg2 ? kb == P.nkb1 : kb == 0;
                    const uint32_t ncol = seg2 ? 3 * ST_U : 2 * ST_U;               // n_h or n_i accumulator columns
#pragma unroll
                    for (int kk = 0; kk < ST_BK / 16; ++kk) {
                        const uint32_t ko = kk * 32;            // 16 halves = 32 bytes along the swizzled row
#pragma unroll
                        for (int term = 0; term < (PREC == 0 ? 3 : 1); ++term) {
                            // small terms first: lo*hi, hi*lo, hi*hi
                            const uint32_t a = (PREC == 0 && term == 0) ? a_lo : a_hi;
                            const uint32_t brz = (PREC == 0 && term == 1) ? rz_lo : rz_hi, bn_ = (PREC == 0 && term == 1) ? n_lo : n_hi;
                            const bool tile_first = kb == 0 && kk == 0 && term == 0;
                            const bool seg_first = seg_first_kb && kk == 0 && term == 0;
#pragma unroll
                            for (int mh = 0; mh < MT; ++mh) {                   // the 128-row halves of an MT = 2 tile share the weight operand
                                const uint64_t ad = umma_desc(a + mh * ST_A_TILE + ko);
                                const uint32_t tacc = tmem_base + mh * ST_ACC_COLS;
                                if (!gru) {
                                    umma_f16<CG>(tacc, ad, umma_desc(brz + ko), id_relu, !tile_first);
                                } else {
                                    umma_f16<CG>(tacc, ad, umma_desc(brz + ko), id_rz, !tile_first);              // r, z: both K parts
                                    umma_f16<CG>(tacc + ncol, ad, umma_desc(bn_ + ko), id_n, !seg_first);          // n_i / n_h apart
                                }
                            }
                        }
                    }
                    umma_commit_cg<CG>(empty0 + 8 * s);
                    if (kb == nkb - 1) umma_commit_cg<CG>(tfull);
                }
                __syncwarp();
            }
        }
    } else {
        // ------------------------------ epilogue (every CTA: its own 128 * MT rows) ------------------------------
        // TMEM lane quarter of a warp = warp id % 4.  MT = 1: the two warps of a quarter split the column chunks;
        // MT = 2: they take one 128-row accumulator each.
        const int ew = warp - 2, q = warp & 3, half = ew >> 2;
        const int acc_i = MT == 2 ? half : 0, c_first = MT == 2 ? 0 : half, c_step = MT == 2 ? 1 : 2;
        const int row = m0 + acc_i * ST_BM + q * 32 + lane;
        const bool valid = row < P.rows;
        const int b = valid ? row / P.E : 0, e = valid ? row - b * P.E : 0;
        const uint32_t tq = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc_i * ST_ACC_COLS);
        const float sc = L.acc_scale;
        const int D = P.D;
        if (gru) {
            const size_t fe = (size_t)(b * P.T + P.t) * P.E + e, fp = (size_t)(b * P.T + P.tprev) * P.E + e;
            const float* xg = P.xg + (fe * 2 + P.dir) * 3 * D;
            const float* hp = P.hx + fp * 2 * D + (size_t)P.dir * D;
            float* ho = P.hx + fe * 2 * D + (size_t)P.dir * D;
            float* gs = P.gsave != nullptr ? P.gsave + (fe * 2 + P.dir) * 4 * D : nullptr;
            const float ug = (valid && P.ugate != nullptr) ? __ldg(P.ugate + fe) : 1.0f;
            // The hoisted pre-activations are read exactly once, from HBM, by every CTA at the same moment (all tiles leave their main
            // loops together): pull this thread's lines into L2 while the main loop runs (measured: epilogue 27k -> cycles, see profiles)
            if (valid) {
#pragma unroll 1
                for (int c = c_first; c < ST_U / 16; c += c_step) {
                    const int ub = nt * ST_U + c * 16;
                    prefetch_l2(xg + ub); prefetch_l2(xg + D + ub); prefetch_l2(xg + 2 * D + ub);
                    if (!P.first) prefetch_l2(hp + ub);
                }
            }
            mbar_wait_backoff(tfull, 0);
            tc_fence_after();
            if (tid == 64) ST_STAMP(2);
#pragma unroll 1
            for (int c = c_first; c < ST_U / 16; c += c_step) {
                const int ub = nt * ST_U + c * 16;
                float ni[16], ar[16], az[16], nh[16];
                tmem_ld16(tq + (uint32_t)(c * 16), ar);
                tmem_ld16(tq + (uint32_t)(ST_U + c * 16), az);
                if (P.nkb1 > 0) tmem_ld16(tq + (uint32_t)(2 * ST_U + c * 16), ni);
                tmem_ld16(tq + (uint32_t)(3 * ST_U + c * 16), nh);
                if (!valid) continue;
                float xr[16], xz[16], xn[16], hprev[16], br[16], bz[16], bn[16], outv[16];
                load16(xg + ub, xr); load16(xg + D + ub, xz); load16(xg + 2 * D + ub, xn);
                load16(P.bhh + ub, br); load16(P.bhh + D + ub, bz); load16(P.bhh + 2 * D + ub, bn);
                if (P.first) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) hprev[j] = 0.0f;
                } else {
                    load16_cg(hp + ub, hprev);
                }
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    // gate order (r, z, n) as torch.nn.GRU / GRUCell: r = s(xr + hr), z = s(xz + hz), n = tanh(xn + r*hn), h' = n + z*(h - n)
                    const float r = sigmoid_fast(xr[j] + sc * ar[j] + br[j]);
                    const float z = sigmoid_fast(xz[j] + sc * az[j] + bz[j]);
                    const float hn = sc * nh[j] + bn[j];
                    const float n = tanh_fast(xn[j] + (P.nkb1 > 0 ? sc * ni[j] : 0.0f) + r * hn);
                    const float hnew = n + z * (hprev[j] - n);
                    outv[j] = ug * hnew + (1.0f - ug) * hprev[j];
                    if (gs != nullptr) {
                        xr[j] = r; xz[j] = z; xn[j] = n; br[j] = hn;
                    }
                }
                store16f(ho + ub, outv);
                store16<PREC>(P.ring_out, (size_t)P.rows * D, (size_t)row * D + ub, outv);
                if (gs != nullptr) {
                    store16f(gs + ub, xr); store16f(gs + D + ub, xz); store16f(gs + 2 * D + ub, xn); store16f(gs + 3 * D + ub, br);
                }
            }
            if (tid == 64) ST_STAMP(60);
        } else {
            float* orow = P.out + (size_t)b * P.out_bstride + (size_t)e * D;
            mbar_wait_backoff(tfull, 0);
            tc_fence_after();
#pragma unroll 1
            for (int c = c_first; c < ST_BN_RELU / 16; c += c_step) {
                const int nb = nt * ST_BN_RELU + c * 16;
                float v[16], bias[16];
                tmem_ld16(tq + (uint32_t)(c * 16), v);
                if (!valid || nb >= D) continue;
                load16(P.bias + nb, bias);
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = fmaxf(sc * v[j] + bias[j], 0.0f);
                store16f(orow + nb, v);
            }
        }
    }
    tc_fence_before();
    if (CG == 2) cluster_sync_all(); else __syncthreads();      // the peer's epilogue must have read its accumulator before the pair's columns go
    if (tid == 0) ST_STAMP(3);
    if (warp == 2) {
        tc_fence_after();
        if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"((uint32_t)Cfg::TMEM_COLS) : "memory");
        else         asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"((uint32_t)Cfg::TMEM_COLS) : "memory");
    }
}

#ifdef ST_TRACE
}  // namespace tg
extern "C" __attribute__((visibility("default"))) int tggcn_debug_trace(long long* out_host) {
    return cudaMemcpyFromSymbol(out_host, tg::g_st_trace, sizeof(long long) * 64) == cudaSuccess ? 0 : 1;
}
extern "C" __attribute__((visibility("default"))) int tggcn_debug_trace_kind(int kind) {
    return cudaMemcpyToSymbol(tg::g_st_trace_kind, &kind, sizeof(int)) == cudaSuccess ? 0 : 1;
}
namespace tg {
#endif

// ---- operand preparation: fp32 weights -> fp16 (hi, lo) planes (scaled) or one bf16 plane ------------------------------------
constexpr int PACK_MAX_JOBS = 20;
struct PackJob {
    const float* src;
    int ld, rows, cols;
    void* hi;                   // [rows][cols]; the lo plane follows at + rows*cols elements
};
struct PackJobs {
    PackJob j[PACK_MAX_JOBS];
    int count;
    unsigned int* err;
};

template <int PREC> __global__ void pack16_kernel(const PackJobs jobs) {
    const PackJob& J = jobs.j[blockIdx.y];
    const size_t n4 = (size_t)J.rows * J.cols / 4, plane = (size_t)J.rows * J.cols;
    const int c4 = J.cols / 4;
    bool bad = false;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / c4), c = (int)(i - (size_t)r * c4) * 4;
        const float4 x = __ldg(reinterpret_cast<const float4*>(J.src + (size_t)r * J.ld + c));
        const float v[4] = {x.x, x.y, x.z, x.w};
        const size_t o = (size_t)r * J.cols + c;
        if (PREC == 0) {
            __align__(8) __half hi[4];
            __align__(8) __half lo[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float s = v[k] * ST_W_SCALE;
                bad |= !(fabsf(s) < 65504.0f);
                hi[k] = __float2half_rn(s);
                lo[k] = __float2half_rn(s - __half2float(hi[k]));
            }
            *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(J.hi) + o) = *reinterpret_cast<const uint2*>(hi);
            *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(J.hi) + plane + o) = *reinterpret_cast<const uint2*>(lo);
        } else {
            __align__(8) __nv_bfloat16 hi[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) hi[k] = __float2bfloat16_rn(v[k]);
            *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(J.hi) + o) = *reinterpret_cast<const uint2*>(hi);
        }
    }
    if (bad && jobs.err != nullptr) atomicOr(jobs.err, 2u);        // bit 1: operand range of the fp16-split tiles
}

// ---- segment-level attention + aggregation between the two GEMMs of a step -------------------------------------------------------
// One CTA per (video, direction).  Senders' post-ReLU messages come from the message GEMM (per kind), the attention logits are
// <receiver state, sender state> / sqrt(D) over the PREVIOUS step's states (vhoi/models.py:1743-1753), masked softmax (NaN -> 0),
// mean pooling under message_aggregation 'mp'.  Writes the aggregated messages as the 16-bit operand rows of the cell GEMM.
constexpr int AT_MAXE = 16;
constexpr int AT_THREADS = 256;
struct AttendParams {
    int B, T, H, O, D, hh, nk_h, mean_pool, att_noscale, first, s, dir_base;
    const float* hx_h; const float* hx_o; const float* om;
    const float* dist[3];                                  // distance-based attention (hh, ho, oo), each may be null
    const float* msg[2][4]; long long msg_bstride[4];      // per direction and kind: sender row (b, e) at msg + b*bstride + e*D
    void* mg16_h; void* mg16_o;                            // planes [dir][hi, lo] of [rows][nk*D]
    float* mg32_h; float* mg32_o; int mg_T;                // (2,B,mg_T,E,nk*D) fp32 copies for the backward, or null
    float* salpha[4];                                      // (2,B,T,receivers,senders) or null
    float* att_f; float* att_b;                            // (B,H,T,O) inspect outputs or null
    unsigned int* err;
};

template <int PREC> __global__ void __launch_bounds__(AT_THREADS) seg_attend_kernel(const AttendParams P) {
    extern __shared__ __align__(16) float at_smem[];
    __shared__ float alpha[4][AT_MAXE][AT_MAXE];
    const int b = blockIdx.x, dir = blockIdx.y + P.dir_base, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int H = P.H, O = P.O, D = P.D, E = H + O, T = P.T;
    pdl_launch_dependents();
    pdl_wait();
    const int t = dir == 0 ? P.s : T - 1 - P.s, tprev = dir == 0 ? t - 1 : t + 1;
    float* hs = at_smem;                                  // [E][D] previous states: humans then objects
    const int D4 = D / 4;
    for (int i = tid; i < E * D4; i += AT_THREADS) {
        const int e = i / D4, c = (i - e * D4) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (!P.first) {
            const float* src = e < H ? P.hx_h + ((size_t)(b * T + tprev) * H + e) * 2 * D + (size_t)dir * D
                                     : P.hx_o + ((size_t)(b * T + tprev) * O + (e - H)) * 2 * D + (size_t)dir * D;
            v = ld_cg4(src + c);
        }
        *reinterpret_cast<float4*>(hs + (size_t)e * D + c) = v;
    }
    __syncthreads();
    // logits into alpha[k][r][s]
    const float scale = P.att_noscale ? 1.0f : 1.0f / sqrtf((float)D);
    for (int k = P.hh ? 0 : 1; k < 4; ++k) {
        const bool send_h = (k == 0 || k == 2), recv_h = (k == 0 || k == 1);
        const int Es = send_h ? H : O, Er = recv_h ? H : O;
        for (int pr = warp; pr < Er * Es; pr += AT_THREADS / 32) {
            const int r = pr / Es, sd = pr - r * Es;
            const float* a = hs + (size_t)(recv_h ? r : H + r) * D;
            const float* c = hs + (size_t)(send_h ? sd : H + sd) * D;
            float acc = 0.0f;
            for (int i = lane; i < D; i += 32) acc = fmaf(a[i], c[i], acc);
            acc = warp_sum(acc);
            if (lane == 0) alpha[k][r][sd] = acc * scale;
        }
    }
    __syncthreads();
    // masked softmax per (kind, receiver)
    if (tid < 4 * AT_MAXE) {
        const int k = tid / AT_MAXE, r = tid - k * AT_MAXE;
        const bool send_h = (k == 0 || k == 2), recv_h = (k == 0 || k == 1), same = send_h == recv_h;
        const int Es = send_h ? H : O, Er = recv_h ? H : O;
        if (r < Er && (k != 0 || P.hh)) {
            bool ok[AT_MAXE];
            float m = -INFINITY;
            for (int sd = 0; sd < Es; ++sd) {
                ok[sd] = !(same && sd == r) && (send_h || __ldg(P.om + b * O + sd) != 0.0f);
                float lg; bool dv;
                if (!P.mean_pool && dist_logit(P.dist, k, (size_t)b * T + t, H, O, r, sd, lg, dv)) {   // models.py:1757-1775
                    alpha[k][r][sd] = lg;
                    ok[sd] = ok[sd] && dv;
                }
                if (ok[sd]) m = fmaxf(m, alpha[k][r][sd]);
            }
            float ex[AT_MAXE], sum = 0.0f;
            for (int sd = 0; sd < Es; ++sd) {
                ex[sd] = ok[sd] ? (P.mean_pool ? 1.0f : expf(alpha[k][r][sd] - m)) : 0.0f;
                sum += ex[sd];
            }
            const float inv = sum > 0.0f ? 1.0f / sum : 0.0f;
            float* sal = P.salpha[k] != nullptr ? P.salpha[k] + ((((size_t)dir * P.B + b) * T + t) * Er + r) * Es : nullptr;
            float* att = (k == 1) ? (dir == 0 ? P.att_f : P.att_b) : nullptr;
            if (att != nullptr) att += ((size_t)(b * H + r) * T + t) * O;
            for (int sd = 0; sd < Es; ++sd) {
                const float a = ex[sd] * inv;
                alpha[k][r][sd] = a;
                if (sal != nullptr) sal[sd] = a;
                if (att != nullptr) att[sd] = a;
            }
        }
    }
    __syncthreads();
    // aggregation: receiver rows of humans (slots: [hh], oh) and objects (slots: ho, oo)
    bool bad = false;
    for (int recv = 0; recv < 2; ++recv) {
        const bool recv_h = recv == 0;
        const int Er = recv_h ? H : O, nk = recv_h ? P.nk_h : 2;
        const size_t rows = (size_t)P.B * Er, KW = (size_t)nk * D;
        void* mg16 = recv_h ? P.mg16_h : P.mg16_o;
        float* mg32 = recv_h ? P.mg32_h : P.mg32_o;
        for (int item = tid; item < Er * nk * (D / 16); item += AT_THREADS) {
            const int cu = item % (D / 16), rs = item / (D / 16), slot = rs % nk, r = rs / nk;
            const int k = recv_h ? (slot == nk - 1 ? 1 : 0) : 2 + slot;
            const int Es = (k == 0 || k == 2) ? H : O;
            float acc[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[j] = 0.0f;
            // the senders' messages were written by the previous launch: plain cached loads (a row is read by every receiver)
            const float* mbase = P.msg[dir][k] + (size_t)b * P.msg_bstride[k] + cu * 16;
            for (int sd = 0; sd < Es; ++sd) {
                const float a = alpha[k][r][sd];
                if (a == 0.0f) continue;                     // masked sender
                float m[16];
                load16(mbase + (size_t)sd * D, m);
#pragma unroll
                for (int j = 0; j < 16; ++j) acc[j] = fmaf(a, m[j], acc[j]);
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) bad |= !(acc[j] < 65504.0f);
            const size_t row = (size_t)b * Er + r, col = (size_t)slot * D + cu * 16;
            uint8_t* plane0 = reinterpret_cast<uint8_t*>(mg16) + (size_t)dir * 2 * rows * KW * 2;       // [dir][hi, lo] planes of 16-bit
            store16<PREC>(plane0, rows * KW, row * KW + col, acc);
            if (mg32 != nullptr)
                store16f(mg32 + ((((size_t)dir * P.B + b) * P.mg_T + (P.mg_T > 1 ? t : 0)) * Er + r) * KW + col, acc);
        }
    }
    if (PREC == 0 && bad && P.err != nullptr) atomicOr(P.err, 2u);
}

// ---- host side ------------------------------------------------------------------------------------------------------
static size_t up1k(size_t x) { return (x + 1023) / 1024 * 1024; }

void big_layout(int B, int H, int O, int D, int hh, BigLayout& L) {
    const size_t h2 = 2;                                   // bytes per 16-bit element
    const size_t nkh = hh ? 2 : 1;
    const size_t rows_g[3] = {(size_t)B * H, (size_t)B * O, (size_t)B};
    size_t off = 0;
    auto take = [&](size_t bytes) { const size_t o = off; off += up1k(bytes); return o; };
    // BiGRU regions first: their offsets must not depend on hh (the BiGRU launcher does not know it)
    for (int g = 0; g < 3; ++g) L.whh_g[g] = take(4 * 3 * (size_t)D * D * h2);
    for (int g = 0; g < 3; ++g) L.ring_g[g] = take(8 * rows_g[g] * D * h2);
    L.wih_h = take(4 * 3 * (size_t)D * nkh * D * h2);
    L.wih_o = take(4 * 3 * (size_t)D * 2 * D * h2);
    L.whh_h = take(4 * 3 * (size_t)D * D * h2);
    L.whh_o = take(4 * 3 * (size_t)D * D * h2);
    L.wm = take(8 * (size_t)D * D * h2);
    L.ring_h = take(8 * rows_g[0] * D * h2);
    L.ring_o = take(8 * rows_g[1] * D * h2);
    L.mg_h = take(4 * rows_g[0] * nkh * D * h2);
    L.mg_o = take(4 * rows_g[1] * 2 * D * h2);
    L.msg = take(2 * 2 * (rows_g[0] + rows_g[1]) * D * sizeof(float));
    L.total = off;
}

// stage 0: frame-level BiGRUs, 1: segment-level graph.  Measured crossovers against the latency path (round-2 sweeps, CUDA events):
// segment level — MPHOI B=32 (128 rows per step) 9.61 -> 9.21 ms, CAD-120 B=32 (160 rows) 44.4 -> 37.1 ms on the step kernels,
// CAD-120 B=16 (80 rows) 25.4 -> 35.3 ms (slower); BiGRUs — the cluster / resident kernels win up to 160 rows (CAD-120 B=32:
// 8.6 vs 12.6 ms), the step kernels from CAD-120 B=64 (320 rows).
bool use_big_path(const tggcn_dims& d, int stage) {
    if (d.D % 64 != 0 || d.D < 128 || d.H > AT_MAXE || d.O > AT_MAXE) return false;
    if (d.recurrent_mode == 2) return true;
    if (d.recurrent_mode == 1) return false;
    const int rows = d.B * (d.H > d.O ? d.H : d.O);
    return rows >= (stage == 1 ? 128 : 192);
}

namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// 3-D tensor map over `planes` row-major [rows][K] matrices of 16-bit elements: box = 64 K-elements x box_rows rows x 1 plane,
// 128-byte swizzle (the K-major layout of the UMMA descriptors), out-of-range rows / columns read as zeros.
int encode_map(CUtensorMap* m, const void* base, int precision, int rank, const size_t* dims_, const int* box_) {
    EncodeTiledFn enc = encode_fn();
    TG_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from this driver");
    TG_REQUIRE(dims_[0] % 8 == 0 && dims_[0] >= ST_BK, "tensor map: K=%zu must be a multiple of 8 and at least %d", dims_[0], ST_BK);
    TG_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "tensor map: base must be 16-byte aligned");
    cuuint64_t dims[4], strides[3];
    cuuint32_t box[4], estr[4];
    cuuint64_t pitch = 2;                                  // bytes per element; the tensors are dense in the order of dims
    for (int i = 0; i < rank; ++i) {
        dims[i] = (cuuint64_t)dims_[i];
        box[i] = (cuuint32_t)box_[i];
        estr[i] = 1;
        pitch *= dims[i];
        if (i < rank - 1) strides[i] = pitch;
    }
    const CUresult r = enc(m, precision ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank,
                           const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    TG_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d): rank %d, K=%zu, box rows %d", (int)r, rank, dims_[0], box_[1]);
    return 0;
}

// 3-D map over `pairs` (hi, lo) plane pairs of row-major [rows][K] matrices of 16-bit elements: one copy moves a box of
// 64 K-elements x box_rows rows x PLANES planes (bf16: the lo plane exists but is not read), 128-byte swizzle (the K-major layout
// of the UMMA descriptors), out-of-range rows / columns read as zeros.
int make_map(CUtensorMap* m, const void* base, int precision, size_t K, size_t rows, size_t planes, int box_rows) {
    const size_t dims[3] = {K, rows, planes};
    const int box[3] = {ST_BK, box_rows, precision ? 1 : 2};
    return encode_map(m, base, precision, 3, dims, box);
}
// 4-D map over gate-major weights: planes of [3][D][K]: box = 64 K-elements x box_units units x box_gates gates x PLANES planes
int make_gate_map(CUtensorMap* m, const void* base, int precision, size_t K, size_t D, size_t planes, int box_units, int box_gates) {
    const size_t dims[4] = {K, D, 3, planes};
    const int box[4] = {ST_BK, box_units, box_gates, precision ? 1 : 2};
    return encode_map(m, base, precision, 4, dims, box);
}

int launch_pack(PackJobs& jobs, int precision, cudaStream_t stream) {
    if (jobs.count == 0) return 0;
    for (int i = 0; i < jobs.count; ++i) {
        const PackJob& j = jobs.j[i];
        TG_REQUIRE(j.cols % 4 == 0 && j.ld % 4 == 0 && (reinterpret_cast<uintptr_t>(j.src) & 15) == 0, "pack16: unaligned weight matrix");
    }
    dim3 grid(64, jobs.count);
    if (precision) pack16_kernel<1><<<grid, 256, 0, stream>>>(jobs);
    else           pack16_kernel<0><<<grid, 256, 0, stream>>>(jobs);
    TG_LAUNCH_OK();
    return 0;
}

void pack_add(PackJobs& jobs, const float* src, int ld, int rows, int cols, void* hi) {
    PackJob& j = jobs.j[jobs.count++];
    j.src = src; j.ld = ld; j.rows = rows; j.cols = cols; j.hi = hi;
}

bool pdl_enabled() {
    static int enabled = -1;
    if (enabled < 0) {
        const char* e = getenv("TGGCN_STEP_PDL");
        enabled = (e != nullptr && e[0] == '0') ? 0 : 1;
    }
    return enabled != 0;
}

template <int PREC, int MT, int CG> int launch_step_t(const StepLaunch& L, int tiles, cudaStream_t stream) {
    using Cfg = StCfg<PREC, MT, CG>;
    if (int rc = ensure_smem((const void*)step_tc_kernel<PREC, MT, CG>, Cfg::SMEM_BYTES)) return rc;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(tiles * CG);
    cfg.blockDim = dim3(ST_THREADS);
    cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;          // CG = 2: the two CTAs of a tile on the two SMs of one TPC
    attr[0].val.clusterDim.x = CG; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;      // see pdl_wait in the kernel
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    // (only while both time directions' grids fit the GPU together: at CAD-120 B=256 — 192 CTAs per step — early dependents
    //  measured 5 % slower, at B=32 / 64 3-5 % faster)
    cfg.numAttrs = (pdl_enabled() && 2 * tiles * CG <= num_sms()) ? 2 : 1;
    TG_CUDA_OK(cudaLaunchKernelEx(&cfg, step_tc_kernel<PREC, MT, CG>, L));
    TG_LAUNCH_OK();
    return 0;
}

// Tile shape of a launch: mt = 128-row accumulators per CTA (1 or 2), cg = CTAs per tile (1, or 2 = tcgen05 cta_group::2 pair).
// The tensor maps of the launch must have been encoded for it: activation boxes of 128 * mt rows, weight boxes of 64 / cg rows.
struct StepShape { int mt, cg; };

int launch_step(StepLaunch& L, int precision, StepShape sh, cudaStream_t stream) {
    int begin = 0;
    for (int i = 0; i < L.count; ++i) {
        StepProblem& p = L.p[i];
        p.m_tiles = cdiv(p.rows, ST_BM * sh.mt * sh.cg);
        p.n_tiles = p.mode == ST_GRU ? p.D / ST_U : cdiv(p.D, ST_BN_RELU);
        p.tile_begin = begin;
        begin += p.m_tiles * p.n_tiles;
    }
    L.acc_scale = precision ? 1.0f : 1.0f / ST_W_SCALE;
    TG_REQUIRE(sh.mt == 1, "step kernel: the two-accumulator (256-row) tile is not built (measured no gain, profiles/r02_step_tile_height.txt)");
    if (sh.cg == 2) return precision ? launch_step_t<1, 1, 2>(L, begin, stream) : launch_step_t<0, 1, 2>(L, begin, stream);
    return precision ? launch_step_t<1, 1, 1>(L, begin, stream) : launch_step_t<0, 1, 1>(L, begin, stream);
}

// Default: single-CTA tiles.  TGGCN_STEP_CG=2 selects CTA pairs (tcgen05 cta_group::2): each SM then stages and reads half of the
// weight tile, the k-block period falls from 2463 to 2184 cycles (clock64 trace) — but a pair's epilogue ends later, and the whole
// forward measured 3.5 % slower (CAD-120 B=256: 27.7 vs 26.8 ms per 64 steps; profiles/r02_step_trace.txt).  The main loop is bound
// by shared-memory bandwidth: every tcgen05.mma re-reads its (128 + N) x 16 operand slice from shared memory, three split terms per
// k-step, while TMA writes the next stage; the period does not move with the batch size, the number of TMA copies per stage (8 -> 2)
// or the stage bytes (56 / 80 / 112 KB).  A two-accumulator 256-row tile (fewer L2 bytes, same shared-memory bytes) was measured at
// no gain and is no longer built.
StepShape choose_shape() {
    static int cg = 0;
    if (cg == 0) {
        const char* e = getenv("TGGCN_STEP_CG");
        cg = (e != nullptr && e[0] == '2') ? 2 : 1;
    }
    return StepShape{1, cg};
}

size_t plane_bytes(size_t rows, size_t K) { return rows * K * 2; }

// The forward and the backward time direction of a recurrence are independent chains (different weights, no exchange inside the
// loop): their step kernels go to two streams, so the GPU packs CTAs of both chains onto its SMs instead of running two half-empty
// waves per launch (a CAD-120 B=256 cell launch is 192 tiles on 148 SMs).  One non-blocking side stream and two events per device,
// created on first use; TGGCN_STEP_STREAMS=1 keeps everything on the caller's stream.
struct DirStreams { cudaStream_t side; cudaEvent_t fork, join; bool split; };
int get_dir_streams(DirStreams& out) {
    struct Entry { bool made; cudaStream_t s; cudaEvent_t f, j; };
    static Entry cache[64];
    static int enabled = -1;
    if (enabled < 0) {
        const char* e = getenv("TGGCN_STEP_STREAMS");
        enabled = (e != nullptr && e[0] == '1') ? 0 : 1;
    }
    out.split = enabled != 0;
    if (!out.split) return 0;
    int dev = 0;
    TG_CUDA_OK(cudaGetDevice(&dev));
    TG_REQUIRE(dev >= 0 && dev < 64, "device ordinal %d out of range", dev);
    Entry& e = cache[dev];
    if (!e.made) {
        TG_CUDA_OK(cudaStreamCreateWithFlags(&e.s, cudaStreamNonBlocking));
        TG_CUDA_OK(cudaEventCreateWithFlags(&e.f, cudaEventDisableTiming));
        TG_CUDA_OK(cudaEventCreateWithFlags(&e.j, cudaEventDisableTiming));
        e.made = true;
    }
    out.side = e.s; out.fork = e.f; out.join = e.j;
    return 0;
}

}  // namespace

int launch_bigru_big(BiGruParams& P, void* big_ws, int precision, cudaStream_t stream) {
    const int D = P.D, T = P.T, B = P.B;
    TG_REQUIRE(D % ST_U == 0 && D >= 2 * ST_BK, "bigru (large-batch path): hidden_size=%d must be a multiple of %d", D, ST_U);
    TG_REQUIRE(P.ngroups == 3, "bigru (large-batch path): expected the three groups of the forward");
    BigLayout BL;
    big_layout(B, P.g[0].E, P.g[1].E, D, 1, BL);          // the BiGRU offsets do not depend on hh
    uint8_t* ws = reinterpret_cast<uint8_t*>(big_ws);
    PackJobs jobs;
    jobs.count = 0;
    jobs.err = P.sync.error;
    for (int g = 0; g < 3; ++g)
        for (int dir = 0; dir < 2; ++dir)
            pack_add(jobs, P.g[g].whh[dir], D, 3 * D, D, ws + BL.whh_g[g] + (size_t)dir * 2 * plane_bytes(3 * D, D));
    if (int rc = launch_pack(jobs, precision, stream)) return rc;
    StepLaunch L;
    memset(&L, 0, sizeof(L));
    const StepShape shape = choose_shape();
    for (int g = 0; g < 3; ++g) {            // tensor maps: activation boxes of 128 * mt rows, weight boxes per tile shape
        const size_t rows = P.g[g].rows;
        if (int rc = make_map(&L.maps[g], ws + BL.ring_g[g], precision, D, rows, 8, ST_BM * shape.mt)) return rc;
        if (shape.cg == 1) {
            if (int rc = make_gate_map(&L.maps[3 + g], ws + BL.whh_g[g], precision, D, D, 4, ST_U, 3)) return rc;
        } else {
            if (int rc = make_gate_map(&L.maps[3 + g], ws + BL.whh_g[g], precision, D, D, 4, ST_U, 1)) return rc;
            if (int rc = make_gate_map(&L.maps[6 + g], ws + BL.whh_g[g], precision, D, D, 4, ST_U / 2, 1)) return rc;
        }
        // the state "before the first step" is zero: slot 1 is what step 0 reads
        TG_CUDA_OK(cudaMemsetAsync(ws + BL.ring_g[g] + 4 * plane_bytes(rows, D), 0, 4 * plane_bytes(rows, D), stream));
    }
    DirStreams ds;
    if (int rc = get_dir_streams(ds)) return rc;
    if (ds.split) {
        TG_CUDA_OK(cudaEventRecord(ds.fork, stream));
        TG_CUDA_OK(cudaStreamWaitEvent(ds.side, ds.fork, 0));
    }
    for (int s = 0; s < T; ++s) {
        const int slot_in = (s & 1) ^ 1, slot_out = s & 1;
        for (int dir = 0; dir < 2; ++dir) {
            if (ds.split || dir == 0) L.count = 0;
            for (int g = 0; g < 3; ++g) {
                StepProblem& q = L.p[L.count++];
                memset(&q, 0, sizeof(q));
                const BiGruGroup& G = P.g[g];
                q.mode = ST_GRU; q.rows = G.rows; q.E = G.E; q.D = D;
                q.nkb1 = 0; q.nkb2 = cdiv(D, ST_BK);
                q.a2_map = g; q.a2_plane = (slot_in * 2 + dir) * 2;
                q.b2_map = 3 + g; q.b2n_map = 6 + g; q.b2_plane = dir * 2;
                q.T = T; q.dir = dir; q.t = dir == 0 ? s : T - 1 - s; q.tprev = dir == 0 ? q.t - 1 : q.t + 1; q.first = s == 0;
                q.xg = G.gi; q.bhh = G.bhh[dir]; q.ugate = nullptr; q.hx = G.hfr; q.gsave = G.gates;
                q.ring_out = ws + BL.ring_g[g] + (size_t)(slot_out * 2 + dir) * 2 * plane_bytes(G.rows, D);
            }
            if (ds.split || dir == 1)
                if (int rc = launch_step(L, precision, shape, (ds.split && dir == 1) ? ds.side : stream)) return rc;
        }
    }
    if (ds.split) {
        TG_CUDA_OK(cudaEventRecord(ds.join, ds.side));
        TG_CUDA_OK(cudaStreamWaitEvent(stream, ds.join, 0));
    }
    return 0;
}

int launch_segment_big(SegParams& P, void* big_ws, int precision, int T_save, cudaStream_t stream) {
    const int D = P.D, T = P.T, B = P.B, H = P.H, O = P.O;
    TG_REQUIRE(D % ST_U == 0 && D >= 2 * ST_BK, "segment (large-batch path): hidden_size=%d must be a multiple of %d", D, ST_U);
    TG_REQUIRE(H <= AT_MAXE && O <= AT_MAXE, "segment (large-batch path): at most %d entities per type", AT_MAXE);
    TG_REQUIRE(O >= 2 && (!P.hh || H >= 2), "segment: need >=2 objects (and >=2 humans with humans->human messages)");
    P.nk_h = P.hh ? 2 : 1;
    const int nkh = P.nk_h;
    if (P.mg_T < 1) P.mg_T = 1;
    BigLayout BL;
    big_layout(B, H, O, D, P.hh, BL);
    uint8_t* ws = reinterpret_cast<uint8_t*>(big_ws);
    const size_t Rh = (size_t)B * H, Ro = (size_t)B * O;
    // ---- weights -> 16-bit operand planes ---------------------------------------------------------------------------------
    PackJobs jobs;
    jobs.count = 0;
    jobs.err = P.sync.error;
    for (int dir = 0; dir < 2; ++dir) {
        pack_add(jobs, P.wih_h[dir] + P.col_h, P.ldw_h, 3 * D, nkh * D, ws + BL.wih_h + (size_t)dir * 2 * plane_bytes(3 * D, nkh * D));
        pack_add(jobs, P.wih_o[dir] + P.col_o, P.ldw_o, 3 * D, 2 * D, ws + BL.wih_o + (size_t)dir * 2 * plane_bytes(3 * D, 2 * D));
        pack_add(jobs, P.whh_h[dir], D, 3 * D, D, ws + BL.whh_h + (size_t)dir * 2 * plane_bytes(3 * D, D));
        pack_add(jobs, P.whh_o[dir], D, 3 * D, D, ws + BL.whh_o + (size_t)dir * 2 * plane_bytes(3 * D, D));
    }
    for (int k = P.hh ? 0 : 1; k < 4; ++k) pack_add(jobs, P.wm[k], D, D, D, ws + BL.wm + (size_t)k * 2 * plane_bytes(D, D));
    if (int rc = launch_pack(jobs, precision, stream)) return rc;
    // ---- tensor maps ---------------------------------------------------------------------------------------------------------
    enum { M_RING_H = 0, M_RING_O, M_MG_H, M_MG_O, M_WIH_H, M_WIH_O, M_WHH_H, M_WHH_O, M_WM, M_WIH_H_N, M_WIH_O_N, M_WHH_H_N, M_WHH_O_N };
    StepLaunch LA, LB;
    memset(&LA, 0, sizeof(LA));
    memset(&LB, 0, sizeof(LB));
    const StepShape shape = choose_shape();
    auto encode_maps = [&](StepLaunch& L) -> int {       // activation boxes: 128 * mt rows; weight boxes: 64 / cg rows
        const int abox = ST_BM * shape.mt, wbox = ST_U / shape.cg;
        if (int rc = make_map(&L.maps[M_RING_H], ws + BL.ring_h, precision, D, Rh, 8, abox)) return rc;
        if (int rc = make_map(&L.maps[M_RING_O], ws + BL.ring_o, precision, D, Ro, 8, abox)) return rc;
        if (int rc = make_map(&L.maps[M_MG_H], ws + BL.mg_h, precision, (size_t)nkh * D, Rh, 4, abox)) return rc;
        if (int rc = make_map(&L.maps[M_MG_O], ws + BL.mg_o, precision, (size_t)2 * D, Ro, 4, abox)) return rc;
        const void* wbase[4] = {ws + BL.wih_h, ws + BL.wih_o, ws + BL.whh_h, ws + BL.whh_o};
        const size_t wk[4] = {(size_t)nkh * D, (size_t)2 * D, (size_t)D, (size_t)D};
        for (int i = 0; i < 4; ++i) {
            if (shape.cg == 1) {
                if (int rc = make_gate_map(&L.maps[M_WIH_H + i], wbase[i], precision, wk[i], D, 4, ST_U, 3)) return rc;
            } else {
                if (int rc = make_gate_map(&L.maps[M_WIH_H + i], wbase[i], precision, wk[i], D, 4, ST_U, 1)) return rc;
                if (int rc = make_gate_map(&L.maps[M_WIH_H_N + i], wbase[i], precision, wk[i], D, 4, ST_U / 2, 1)) return rc;
            }
        }
        (void)wbox;
        return make_map(&L.maps[M_WM], ws + BL.wm, precision, D, D, 8, ST_BN_RELU / shape.cg);
    };
    TG_CUDA_OK(cudaMemsetAsync(ws + BL.ring_h + 4 * plane_bytes(Rh, D), 0, 4 * plane_bytes(Rh, D), stream));
    TG_CUDA_OK(cudaMemsetAsync(ws + BL.ring_o + 4 * plane_bytes(Ro, D), 0, 4 * plane_bytes(Ro, D), stream));

    // per-step message buffers: the save buffers of the backward when training, a scratch region otherwise
    const bool saving = P.smsg[1] != nullptr;
    float* msg_scratch = reinterpret_cast<float*>(ws + BL.msg);
    const size_t at_smem = sizeof(float) * (size_t)(H + O) * D;
    TG_REQUIRE(at_smem <= 200 * 1024, "segment (large-batch path): %zu bytes of shared memory for the attention kernel", at_smem);
    if (precision) { if (int rc = ensure_smem((const void*)seg_attend_kernel<1>, at_smem)) return rc; }
    else           { if (int rc = ensure_smem((const void*)seg_attend_kernel<0>, at_smem)) return rc; }

    if (int rc = encode_maps(LA)) return rc;
    memcpy(LB.maps, LA.maps, sizeof(LA.maps));
    DirStreams ds;
    if (int rc = get_dir_streams(ds)) return rc;
    if (ds.split) {
        TG_CUDA_OK(cudaEventRecord(ds.fork, stream));
        TG_CUDA_OK(cudaStreamWaitEvent(ds.side, ds.fork, 0));
    }
    for (int s = 0; s < T; ++s) {
        const int slot_in = (s & 1) ^ 1, slot_out = s & 1;
        // the two time directions are independent chains of (messages -> attention -> cells): one stream each
        for (int pass = 0; pass < (ds.split ? 2 : 1); ++pass) {
            const int dir_lo = ds.split ? pass : 0, dir_hi = ds.split ? pass + 1 : 2;
            cudaStream_t st = (ds.split && pass == 1) ? ds.side : stream;
            AttendParams A;
            memset(&A, 0, sizeof(A));
            // ---- phase A1: per-sender messages msg = ReLU(W_k s_prev + b_k) for every kind ----------------------------------------
            LA.count = 0;
            for (int dir = dir_lo; dir < dir_hi; ++dir) {
                const int t = dir == 0 ? s : T - 1 - s;
                size_t scratch_off = (size_t)dir * 2 * (Rh + Ro) * D;
                for (int k = 0; k < 4; ++k) {
                    const bool send_h = (k == 0 || k == 2);
                    const int Es = send_h ? H : O;
                    const size_t Rs = send_h ? Rh : Ro;
                    float* out;
                    long long bstride;
                    if (saving) { out = P.smsg[k] != nullptr ? P.smsg[k] + (((size_t)dir * B) * T + t) * Es * D : nullptr; bstride = (long long)T * Es * D; }
                    else        { out = msg_scratch + scratch_off; bstride = (long long)Es * D; }
                    scratch_off += Rs * D;
                    A.msg[dir][k] = out; A.msg_bstride[k] = bstride;
                    if (k == 0 && !P.hh) continue;
                    StepProblem& q = LA.p[LA.count++];
                    memset(&q, 0, sizeof(q));
                    q.mode = ST_RELU; q.rows = (int)Rs; q.E = Es; q.D = D;
                    q.nkb1 = 0; q.nkb2 = cdiv(D, ST_BK);
                    q.a2_map = send_h ? M_RING_H : M_RING_O; q.a2_plane = (slot_in * 2 + dir) * 2;
                    q.b2_map = M_WM; q.b2_plane = k * 2;
                    q.bias = P.bm[k]; q.out = out; q.out_bstride = bstride;
                }
            }
            if (int rc = launch_step(LA, precision, shape, st)) return rc;
            // ---- phase A2: attention over the previous states + aggregation -> operand rows of the cell GEMM --------------------
            A.B = B; A.T = T; A.H = H; A.O = O; A.D = D; A.hh = P.hh; A.nk_h = nkh; A.mean_pool = P.mean_pool; A.att_noscale = P.att_noscale; A.first = s == 0; A.s = s;
            for (int i = 0; i < 3; ++i) A.dist[i] = P.dist[i];
            A.dir_base = dir_lo;
            A.hx_h = P.hx_h; A.hx_o = P.hx_o; A.om = P.om;
            A.mg16_h = ws + BL.mg_h; A.mg16_o = ws + BL.mg_o;
            A.mg32_h = P.mg_T > 1 ? P.mg_h : nullptr; A.mg32_o = P.mg_T > 1 ? P.mg_o : nullptr; A.mg_T = P.mg_T;
            for (int k = 0; k < 4; ++k) A.salpha[k] = P.salpha[k];
            A.att_f = P.att_f; A.att_b = P.att_b; A.err = P.sync.error;
            {
                cudaLaunchConfig_t cfg;
                memset(&cfg, 0, sizeof(cfg));
                cfg.gridDim = dim3(B, dir_hi - dir_lo);
                cfg.blockDim = dim3(AT_THREADS);
                cfg.dynamicSmemBytes = at_smem;
                cfg.stream = st;
                cudaLaunchAttribute attr[1];
                attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
                attr[0].val.programmaticStreamSerializationAllowed = 1;
                cfg.attrs = attr;
                cfg.numAttrs = (pdl_enabled() && (int)(Rh + Ro) <= 1024) ? 1 : 0;
                if (precision) TG_CUDA_OK(cudaLaunchKernelEx(&cfg, seg_attend_kernel<1>, A));
                else           TG_CUDA_OK(cudaLaunchKernelEx(&cfg, seg_attend_kernel<0>, A));
            }
            TG_LAUNCH_OK();
            // ---- phase B: gated GRU cells -----------------------------------------------------------------------------------------
            LB.count = 0;
            for (int dir = dir_lo; dir < dir_hi; ++dir)
                for (int type = 0; type < 2; ++type) {
                    const bool is_h = type == 0;
                    StepProblem& q = LB.p[LB.count++];
                    memset(&q, 0, sizeof(q));
                    const int E = is_h ? H : O, nk = is_h ? nkh : 2;
                    const size_t R = is_h ? Rh : Ro;
                    q.mode = ST_GRU; q.rows = (int)R; q.E = E; q.D = D;
                    q.nkb1 = cdiv(nk * D, ST_BK); q.nkb2 = cdiv(D, ST_BK);
                    q.a1_map = is_h ? M_MG_H : M_MG_O; q.a1_plane = dir * 2;
                    q.a2_map = is_h ? M_RING_H : M_RING_O; q.a2_plane = (slot_in * 2 + dir) * 2;
                    q.b1_map = is_h ? M_WIH_H : M_WIH_O; q.b1n_map = is_h ? M_WIH_H_N : M_WIH_O_N; q.b1_plane = dir * 2;
                    q.b2_map = is_h ? M_WHH_H : M_WHH_O; q.b2n_map = is_h ? M_WHH_H_N : M_WHH_O_N; q.b2_plane = dir * 2;
                    q.T = T; q.dir = dir; q.t = dir == 0 ? s : T - 1 - s; q.tprev = dir == 0 ? q.t - 1 : q.t + 1; q.first = s == 0;
                    q.xg = is_h ? P.gs_h : P.gs_o; q.bhh = is_h ? P.bhh_h[dir] : P.bhh_o[dir];
                    q.ugate = is_h ? P.u_h : P.u_o; q.hx = is_h ? P.hx_h : P.hx_o; q.gsave = is_h ? P.sgates_h : P.sgates_o;
                    q.ring_out = ws + (is_h ? BL.ring_h : BL.ring_o) + (size_t)(slot_out * 2 + dir) * 2 * plane_bytes(R, D);
                }
            if (int rc = launch_step(LB, precision, shape, st)) return rc;
        }
    }
    if (ds.split) {
        TG_CUDA_OK(cudaEventRecord(ds.join, ds.side));
        TG_CUDA_OK(cudaStreamWaitEvent(stream, ds.join, 0));
    }
    (void)T_save;
    return 0;
}

}  // namespace tg
