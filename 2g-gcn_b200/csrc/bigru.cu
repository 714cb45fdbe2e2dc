// Frame-level bidirectional GRUs (K-C): the recurrences of human_bd_rnn / object_bd_rnn /
// geometry_bd_rnn (vhoi/models.py:649-651, :983-1002) for every entity and both directions in ONE
// persistent cooperative kernel.  Input pre-activations W_ih x + b_ih were hoisted into a batched
// projection; each step computes W_hh h for all rows, the gate math, and publishes h through L2,
// followed by one grid barrier.  Entities fold into the batch (same weights), padding is processed
// like the reference does (no packing).
#include "recurrent.cuh"
#include "bigru.h"

namespace tg {

template <int NRG, int KC>
__device__ __forceinline__ void bigru_tile(const BiGruParams& P, const BiGruGroup& G, int local, int s, float* smem,
                                           const float** xrows) {
    constexpr int RB = TileGeom<NRG>::RB, J = TileGeom<NRG>::J;
    const int D = P.D, T = P.T;
    // local tile index -> (dir, row block, unit block)
    const int per_dir = G.n_rb * G.n_ub;
    const int dir = local / per_dir;
    const int rem = local - dir * per_dir;
    const int rb = rem / G.n_ub, ub = rem - rb * G.n_ub;
    const int row0 = rb * RB;
    const int unit0 = ub * G.jeff;
    const int unit_end = min(unit0 + G.jeff, D);
    const int t = dir == 0 ? s : T - 1 - s;
    const int tprev = dir == 0 ? t - 1 : t + 1;
    const int tid = threadIdx.x;

    __syncthreads();   // previous tile may still be reading xrows / smem
    if (tid < RB) {
        const int r = row0 + tid;
        const float* ptr = nullptr;
        if (r < G.rows && s > 0) {
            const int b = r / G.E, e = r - b * G.E;
            ptr = G.hfr + ((size_t)(b * T + tprev) * G.E + e) * 2 * D + dir * D;
        }
        xrows[tid] = ptr;
    }
    __syncthreads();

    float acc[3][2] = {{0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}};
    if (s > 0)
        tile_accumulate<3, NRG, KC, RB>(acc, G.whh[dir], D, D, unit0, unit_end, 3 * D, xrows, D, smem, NoHook());

    const int j = (tid & 15) + 16 * (tid / (16 * NRG));
    const int rg = (tid >> 4) % NRG;
    const int unit = unit0 + j;
    if (unit < unit_end) {
        const float* bhh = G.bhh[dir];
        const float br = __ldg(bhh + unit), bz = __ldg(bhh + D + unit), bn = __ldg(bhh + 2 * D + unit);
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
            const int r = row0 + 2 * rg + rr;
            if (r >= G.rows) continue;
            const int b = r / G.E, e = r - b * G.E;
            const float* gi = G.gi + (((size_t)(b * T + t) * G.E + e) * 2 + dir) * 3 * D;
            const float hprev = s > 0 ? ld_cg(G.hfr + ((size_t)(b * T + tprev) * G.E + e) * 2 * D + dir * D + unit) : 0.0f;
            const float h = gru_update(__ldg(gi + unit), __ldg(gi + D + unit), __ldg(gi + 2 * D + unit),
                                       acc[0][rr] + br, acc[1][rr] + bz, acc[2][rr] + bn, hprev);
            G.hfr[((size_t)(b * T + t) * G.E + e) * 2 * D + dir * D + unit] = h;
        }
    }
}

template <int KC>
__global__ void __launch_bounds__(REC_THREADS) bigru_kernel(const BiGruParams P, int s_begin, int s_end, int persistent) {
    extern __shared__ __align__(16) float smem[];
    __shared__ const float* xrows[32];
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
            const int local = tile - G.tile_begin;
            switch (G.nrg) {
                case 4:  bigru_tile<4, KC>(P, G, local, s, smem, xrows); break;
                case 8:  bigru_tile<8, KC>(P, G, local, s, smem, xrows); break;
                default: bigru_tile<16, KC>(P, G, local, s, smem, xrows); break;
            }
        }
        if (persistent && s + 1 < s_end) {
            if (!grid_barrier(P.sync, epoch, gridDim.x, &s_fail)) return;
        }
    }
}

static int pick_nrg(int rows) { return rows <= 8 ? 4 : (rows <= 16 ? 8 : 16); }

size_t bigru_smem_bytes(int KC) {
    // worst case over NRG in {4,8,16}: NG*J + RB rows of (KC+4) floats, double buffered
    size_t m = 0;
    const int nrgs[3] = {4, 8, 16};
    for (int i = 0; i < 3; ++i) {
        const size_t rows = 3 * (REC_THREADS / nrgs[i]) + 2 * nrgs[i];
        const size_t b = 2 * rows * (KC + 4) * sizeof(float);
        if (b > m) m = b;
    }
    return m;
}

// Choose the per-group unit-block sizes so that the whole step is one wave of nearly equal tiles.
static void plan_tiles(BiGruParams& P, int capacity) {
    int best_tau = 1 << 30;
    for (int tau = 8; tau <= 64 * 64; ++tau) {
        int tiles = 0;
        for (int i = 0; i < P.ngroups; ++i) {
            BiGruGroup& G = P.g[i];
            const int RB = 2 * G.nrg, J = REC_THREADS / G.nrg;
            int jeff = tau / RB; if (jeff < 1) jeff = 1; if (jeff > J) jeff = J; if (jeff > P.D) jeff = P.D;
            tiles += 2 * cdiv(G.rows, RB) * cdiv(P.D, jeff);
        }
        if (tiles <= capacity) { best_tau = tau; break; }
    }
    int begin = 0;
    for (int i = 0; i < P.ngroups; ++i) {
        BiGruGroup& G = P.g[i];
        const int RB = 2 * G.nrg, J = REC_THREADS / G.nrg;
        int jeff = best_tau == (1 << 30) ? J : best_tau / RB;
        if (jeff < 1) jeff = 1; if (jeff > J) jeff = J; if (jeff > P.D) jeff = P.D;
        G.jeff = jeff;
        G.n_rb = cdiv(G.rows, RB);
        G.n_ub = cdiv(P.D, jeff);
        G.tile_begin = begin;
        begin += 2 * G.n_rb * G.n_ub;
    }
    P.total_tiles = begin;
}

int launch_bigru(BiGruParams& P, int persistent, cudaStream_t stream) {
    TG_REQUIRE(P.D % 16 == 0, "bigru: hidden_size=%d must be a multiple of 16", P.D);
    const int KC = (P.D % 32 == 0) ? 32 : 16;
    for (int i = 0; i < P.ngroups; ++i) P.g[i].nrg = pick_nrg(P.g[i].rows);
    const size_t smem = bigru_smem_bytes(KC);
    auto kern = KC == 32 ? bigru_kernel<32> : bigru_kernel<16>;
    TG_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    TG_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, REC_THREADS, smem));
    TG_REQUIRE(per_sm >= 1, "bigru: kernel does not fit on an SM (smem %zu)", smem);
    const int capacity = per_sm * num_sms();
    plan_tiles(P, num_sms());   // one tile per SM per step when possible
    if (persistent) {
        const int grid = P.total_tiles < capacity ? P.total_tiles : capacity;
        TG_CUDA_OK(cudaMemsetAsync(P.sync.counter, 0, 2 * sizeof(unsigned int), stream));
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
