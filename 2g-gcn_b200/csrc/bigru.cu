// Frame-level bidirectional GRUs (K-C): the recurrences of human_bd_rnn / object_bd_rnn /
// geometry_bd_rnn (vhoi/models.py:649-651, :983-1002) for every entity and both directions in ONE
// persistent cooperative kernel.  Input pre-activations W_ih x + b_ih were hoisted into a batched
// projection; each step computes W_hh h for all rows, the gate math, and publishes h through L2,
// followed by one grid barrier.  Entities fold into the batch (same weights), padding is processed
// like the reference does (no packing).
#include "recurrent.cuh"
#include "bigru.h"
#include "step_tc.cuh"

namespace tg {

template <int NT, int STAGES>
__device__ __forceinline__ void bigru_tile(const BiGruParams& P, const BiGruGroup& G, int local, int s, float* smem,
                                           const float** tab) {
    constexpr int RBT = 8 * NT, NPAIR = NT / 2;
    const int D = P.D, T = P.T;
    // local tile index -> (dir, row block, unit block)
    const int per_dir = G.n_rb * G.n_ub;
    const int dir = local / per_dir;
    const int rem = local - dir * per_dir;
    const int rb = rem / G.n_ub, ub = rem - rb * G.n_ub;
    const int row0 = rb * RBT, unit0 = ub * REC_J;
    const int t = dir == 0 ? s : T - 1 - s;
    const int tprev = dir == 0 ? t - 1 : t + 1;
    const int tid = threadIdx.x;

    __syncthreads();   // previous tile may still be reading the pointer table
    if (tid < 3 * REC_J) {
        const int g = tid / REC_J, unit = unit0 + tid % REC_J;
        tab[tid] = unit < D ? G.whh[dir] + (size_t)(g * D + unit) * D : nullptr;
    } else if (tid < 3 * REC_J + RBT) {
        const int r = row0 + tid - 3 * REC_J;
        const float* ptr = nullptr;
        if (r < G.rows && s > 0) {
            const int b = r / G.E, e = r - b * G.E;
            ptr = G.hfr + ((size_t)(b * T + tprev) * G.E + e) * 2 * D + dir * D;
        }
        tab[tid] = ptr;
    }
    // epilogue operands are fetched now so that their latency overlaps the K loop
    const int unit = unit0 + (tid & 15);
    float xg[NPAIR][3], hprev[NPAIR], bh[3] = {0.f, 0.f, 0.f};
    bool valid[NPAIR];
    size_t orow[NPAIR];
    float* gsave[NPAIR];
    if (unit < D) {
        bh[0] = __ldg(G.bhh[dir] + unit); bh[1] = __ldg(G.bhh[dir] + D + unit); bh[2] = __ldg(G.bhh[dir] + 2 * D + unit);
    }
#pragma unroll
    for (int p = 0; p < NPAIR; ++p) {
        const int r = row0 + (tid >> 4) + 16 * p;
        valid[p] = unit < D && r < G.rows && ((tid >> 4) + 16 * p) < RBT;
        xg[p][0] = xg[p][1] = xg[p][2] = hprev[p] = 0.0f;
        orow[p] = 0;
        gsave[p] = nullptr;
        if (valid[p]) {
            const int b = r / G.E, e = r - b * G.E;
            const float* gi = G.gi + (((size_t)(b * T + t) * G.E + e) * 2 + dir) * 3 * D;
            xg[p][0] = __ldg(gi + unit); xg[p][1] = __ldg(gi + D + unit); xg[p][2] = __ldg(gi + 2 * D + unit);
            if (s > 0) hprev[p] = ld_cg(G.hfr + ((size_t)(b * T + tprev) * G.E + e) * 2 * D + dir * D + unit);
            orow[p] = ((size_t)(b * T + t) * G.E + e) * 2 * D + dir * D + unit;
            gsave[p] = G.gates != nullptr ? G.gates + (((size_t)(b * T + t) * G.E + e) * 2 + dir) * 4 * D + unit : nullptr;
        }
    }
    __syncthreads();

    float acc[3][NPAIR];
    tile_accumulate<3, NT, STAGES>(acc, tab, tab, s > 0 ? D : 0, 0, 0u, 0u, G.whh[dir], smem);

#pragma unroll
    for (int p = 0; p < NPAIR; ++p)
        if (valid[p])
            G.hfr[orow[p]] = gru_update(xg[p][0], xg[p][1], xg[p][2], acc[0][p] + bh[0], acc[1][p] + bh[1], acc[2][p] + bh[2], hprev[p],
                                        gsave[p], D);
}

__global__ void __launch_bounds__(REC_THREADS, 1) bigru_kernel(const BiGruParams P, int s_begin, int s_end, int persistent) {
    extern __shared__ __align__(16) float smem[];
    __shared__ const float* tab[3 * REC_J + 32];
    __shared__ int s_fail;
    if (threadIdx.x == 0) s_fail = 0;
    unsigned int epoch = 0;
    for (int s = s_begin; s < s_end; ++s) {
        for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
            int gi = 0;
#pragma unroll 1
            for (int i = 1; i < P.ngroups; ++i)
                if (tile >= P.g[i].tile_begin) gi = i;
            const BiGruGroup& G = P.g[gi];
            if (G.cfg == 4) bigru_tile<4, 3>(P, G, tile - G.tile_begin, s, smem, tab);
            else            bigru_tile<2, 4>(P, G, tile - G.tile_begin, s, smem, tab);
        }
        if (persistent && s + 1 < s_end) {
            if (!grid_barrier(P.sync, epoch, gridDim.x, &s_fail)) return;
        }
    }
}

static void plan_tiles(BiGruParams& P) {
    int begin = 0;
    for (int i = 0; i < P.ngroups; ++i) {
        BiGruGroup& G = P.g[i];
        G.cfg = G.rows > 16 ? 4 : 2;          // n8 tiles of rows per tile
        G.jeff = REC_J;
        G.n_rb = cdiv(G.rows, 8 * G.cfg);
        G.n_ub = cdiv(P.D, REC_J);
        G.tile_begin = begin;
        begin += 2 * G.n_rb * G.n_ub;
    }
    P.total_tiles = begin;
}

int launch_bigru(BiGruParams& P, int persistent, cudaStream_t stream) {
    if (P.big_ws != nullptr) return launch_bigru_big(P, P.big_ws, P.precision, stream);
    TG_REQUIRE(P.D % 16 == 0, "bigru: hidden_size=%d must be a multiple of 16", P.D);
    const int fa = tile_smem_floats(3, 4, 3), fb = tile_smem_floats(3, 2, 4);
    const size_t smem = sizeof(float) * (size_t)(fa > fb ? fa : fb);
    auto kern = bigru_kernel;
    if (int rc = ensure_smem((const void*)kern, smem)) return rc;
    int per_sm = 0;
    TG_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, REC_THREADS, smem));
    TG_REQUIRE(per_sm >= 1, "bigru: kernel does not fit on an SM (smem %zu)", smem);
    const int capacity = per_sm * num_sms();
    plan_tiles(P);
    if (persistent) {
        int rc = launch_bigru_cluster(P, stream);             // independent recurrences on 16-CTA clusters, state exchanged through DSMEM
        if (rc >= 0) return rc;
        rc = launch_bigru_resident(P, stream);                // recurrent weights resident in shared memory when the shape allows
        if (rc >= 0) return rc;
        const int grid = P.total_tiles < capacity ? P.total_tiles : capacity;
        TG_CUDA_OK(cudaMemsetAsync(P.sync.counter, 0, sizeof(unsigned int), stream));      // the error word belongs to the caller (tggcn_forward zeroes it once)
        int s0 = 0, s1 = P.T, pers = 1;
        void* args[] = {(void*)&P, (void*)&s0, (void*)&s1, (void*)&pers};
        TG_CUDA_OK(cudaLaunchCooperativeKernel((const void*)kern, dim3(grid), dim3(REC_THREADS), args, smem, stream));
        ++g_launches;
    } else {
        for (int s = 0; s < P.T; ++s) {
            kern<<<P.total_tiles, REC_THREADS, smem, stream>>>(P, s, s + 1, 0);
            TG_LAUNCH_OK();
        }
    }
    return 0;
}

}  // namespace tg
