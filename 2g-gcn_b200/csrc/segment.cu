// Segment-level recurrent graph (K-F): the loop of vhoi/models.py:785-880 — per step, attention
// messages over the PREVIOUS step's segment states of all entities (:1051,:1145,:1239,:1334), then a
// gated GRUCell update h <- u*GRUCell([x_frame, mg_a, mg_b], h) + (1-u)*h (:1535-1564) for every human
// and object, forward and backward direction in lock-step — as ONE persistent cooperative kernel.
//
// Each step has two phases separated by grid barriers:
//   A  message tiles : msg = ReLU(W s_prev + b) for a slice of output units of one message kind, the
//                      scaled-dot-product attention weights of the receivers (accumulated from the same
//                      K-chunks), and the aggregated message mg = sum_s alpha_s msg_s  -> L2
//   B  cell tiles    : W_ih[:, seg cols] mg + (hoisted frame part) , W_hh h_prev, gate math, blend with
//                      the hard gate u, publish the new state (which is also the output row)  -> L2
// The frame-part of W_ih x (3D/4D of the 5D/6D input columns) was hoisted into a batched projection.
#include "recurrent.cuh"
#include "bigru.h"

namespace tg {

constexpr int MSG_NRG = 16;                 // message tiles: 32 sender rows x 48 units
constexpr int MSG_UNITS = 3 * (REC_THREADS / MSG_NRG);
constexpr int MSG_LDM = MSG_UNITS + 1;
constexpr int MSG_MAXPAIRS = 512;

struct SegShared {
    const float* xrows[64];
    float msg[32 * MSG_LDM];
    float logit[MSG_MAXPAIRS];
    float alpha[MSG_MAXPAIRS];
    int s_fail;
};

struct PairHook {
    int nmine;
    int rowR[2], rowS[2];
    float* lg;
    int KC;
    __device__ __forceinline__ void operator()(const float* Xs, int LD) const {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            if (i < nmine) {
                const float* a = Xs + rowR[i] * LD;
                const float* b = Xs + rowS[i] * LD;
                float acc = lg[i];
                for (int k = 0; k < KC; k += 4) {
                    const float4 x = *reinterpret_cast<const float4*>(a + k);
                    const float4 y = *reinterpret_cast<const float4*>(b + k);
                    acc = fmaf(x.x, y.x, acc); acc = fmaf(x.y, y.y, acc);
                    acc = fmaf(x.z, y.z, acc); acc = fmaf(x.w, y.w, acc);
                }
                lg[i] = acc;
            }
        }
    }
};

// ---- phase A ------------------------------------------------------------------------------------
template <int KC>
__device__ __forceinline__ void seg_message_tile(const SegParams& P, int tile, int s, float* smem, SegShared& sh) {
    const int D = P.D, T = P.T, B = P.B, H = P.H, O = P.O;
    const int dir = tile / P.msg_tiles_dir;
    int rem = tile - dir * P.msg_tiles_dir;
    int kind = 0;
#pragma unroll
    for (int k = 1; k < 4; ++k)
        if (rem >= P.msg_tile_begin[k]) kind = k;
    rem -= P.msg_tile_begin[kind];
    const int nub = P.msg_tiles_kind[kind] / P.n_vb;
    const int vb = rem / nub, ub = rem - vb * nub;
    const bool send_h = (kind == 0 || kind == 2);       // sender type: humans for hh, ho
    const bool recv_h = (kind == 0 || kind == 1);       // receiver type: humans for hh, oh
    const int Es = send_h ? H : O, Er = recv_h ? H : O;
    const bool same = (send_h == recv_h);
    const int b0 = vb * P.bbv, nb = min(P.bbv, B - b0);
    const int unit0 = ub * MSG_UNITS;
    const int t = dir == 0 ? s : T - 1 - s;
    const int tprev = dir == 0 ? t - 1 : t + 1;
    const int tid = threadIdx.x;

    __syncthreads();
    if (tid < 64) {
        const float* ptr = nullptr;
        if (s > 0) {
            if (tid < 32) {
                const int bl = tid / Es, e = tid - bl * Es;
                if (bl < nb) {
                    const float* base = send_h ? P.hx_h : P.hx_o;
                    ptr = base + ((size_t)((b0 + bl) * T + tprev) * Es + e) * 2 * D + dir * D;
                }
            } else if (!same) {
                const int l = tid - 32, bl = l / Er, e = l - bl * Er;
                if (bl < nb) {
                    const float* base = recv_h ? P.hx_h : P.hx_o;
                    ptr = base + ((size_t)((b0 + bl) * T + tprev) * Er + e) * 2 * D + dir * D;
                }
            }
        }
        sh.xrows[tid] = ptr;
    }
    // attention pairs owned by this thread: p = (bl*Er + r)*Es + sdr
    const int npairs = nb * Er * Es;
    float lg[2] = {0.f, 0.f};
    PairHook hook;
    hook.lg = lg; hook.KC = KC; hook.nmine = 0;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int p = tid + i * REC_THREADS;
        hook.rowR[i] = 0; hook.rowS[i] = 0;
        if (p < npairs) {
            const int sdr = p % Es, br = p / Es;
            const int r = br % Er, bl = br / Er;
            hook.rowS[i] = bl * Es + sdr;
            hook.rowR[i] = same ? bl * Es + r : 32 + bl * Er + r;
            hook.nmine = i + 1;
        }
    }
    __syncthreads();

    float acc[3][2] = {{0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}};
    if (s > 0)
        tile_accumulate<3, MSG_NRG, KC, 64>(acc, P.wm[kind], D, REC_THREADS / MSG_NRG, unit0, D, D, sh.xrows, D, smem, hook);

    // messages of this unit slice for every sender row
    {
        const int j = tid & 15, rg = tid >> 4;
#pragma unroll
        for (int g = 0; g < 3; ++g) {
            const int c = g * 16 + j, u = unit0 + c;
            const float bias = u < D ? __ldg(P.bm[kind] + u) : 0.0f;
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) sh.msg[(2 * rg + rr) * MSG_LDM + c] = fmaxf(acc[g][rr] + bias, 0.0f);
        }
    }
    const float scale = 1.0f / sqrtf((float)D);
#pragma unroll
    for (int i = 0; i < 2; ++i)
        if (i < hook.nmine) sh.logit[tid + i * REC_THREADS] = lg[i] * scale;
    __syncthreads();
    // masked softmax over the senders of each receiver (vhoi/models.py:1750-1753)
    if (tid < nb * Er) {
        const int bl = tid / Er, r = tid - bl * Er;
        const int b = b0 + bl;
        const float* lrow = sh.logit + tid * Es;
        float m = -INFINITY;
        for (int sdr = 0; sdr < Es; ++sdr) {
            bool ok = !(same && sdr == r);
            if (!send_h) ok = ok && (P.om[b * O + sdr] != 0.0f);
            if (ok) m = fmaxf(m, lrow[sdr]);
        }
        float sum = 0.0f;
        for (int sdr = 0; sdr < Es; ++sdr) {
            bool ok = !(same && sdr == r);
            if (!send_h) ok = ok && (P.om[b * O + sdr] != 0.0f);
            const float ex = ok ? expf(lrow[sdr] - m) : 0.0f;
            sh.alpha[tid * Es + sdr] = ex;
            sum += ex;
        }
        const float inv = sum > 0.0f ? 1.0f / sum : 0.0f;
        for (int sdr = 0; sdr < Es; ++sdr) {
            const float a = sh.alpha[tid * Es + sdr] * inv;
            sh.alpha[tid * Es + sdr] = a;
            if (kind == 1 && ub == 0) {
                float* att = dir == 0 ? P.att_f : P.att_b;
                if (att != nullptr) att[((size_t)(b * H + r) * T + t) * O + sdr] = a;
            }
        }
    }
    __syncthreads();
    // aggregated message for every receiver of the block
    const int nk_r = recv_h ? P.nk_h : 2;
    const int slot = recv_h ? (kind == 0 ? 0 : P.nk_h - 1) : kind - 2;
    float* mg = recv_h ? P.mg_h : P.mg_o;
    for (int idx = tid; idx < nb * Er * MSG_UNITS; idx += REC_THREADS) {
        const int c = idx % MSG_UNITS, br = idx / MSG_UNITS;
        const int u = unit0 + c;
        if (u >= D) continue;
        const int bl = br / Er;
        const float* al = sh.alpha + br * Es;
        const float* ms = sh.msg + (bl * Es) * MSG_LDM + c;
        float v = 0.0f;
        for (int sdr = 0; sdr < Es; ++sdr) v = fmaf(al[sdr], ms[sdr * MSG_LDM], v);
        const int r = br - bl * Er;
        mg[(((size_t)dir * B + b0 + bl) * Er + r) * nk_r * D + slot * D + u] = v;
    }
}

// ---- phase B ------------------------------------------------------------------------------------
template <int NRG, int KC>
__device__ __forceinline__ void seg_cell_tile(const SegParams& P, bool is_h, int dir, int rb, int ub, int s, float* smem,
                                              SegShared& sh) {
    constexpr int RB = TileGeom<NRG>::RB;
    const int D = P.D, T = P.T, B = P.B;
    const int E = is_h ? P.H : P.O;
    const int rows = B * E;
    const int jeff = is_h ? P.jeff_h : P.jeff_o;
    const int row0 = rb * RB, unit0 = ub * jeff, unit_end = min(unit0 + jeff, D);
    const int t = dir == 0 ? s : T - 1 - s;
    const int tprev = dir == 0 ? t - 1 : t + 1;
    const int tid = threadIdx.x;
    const int nk = is_h ? P.nk_h : 2;
    float* hx = is_h ? P.hx_h : P.hx_o;
    const float* mgbase = is_h ? P.mg_h : P.mg_o;

    __syncthreads();
    if (tid < RB) {
        const int r = row0 + tid;
        sh.xrows[tid] = r < rows ? mgbase + ((size_t)dir * rows + r) * nk * D : nullptr;
    }
    __syncthreads();
    float acc_i[3][2] = {{0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}};
    float acc_h[3][2] = {{0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}};
    {
        const float* W = (is_h ? P.wih_h[dir] + P.col_h : P.wih_o[dir] + P.col_o);
        const int ldw = is_h ? P.ldw_h : P.ldw_o;
        tile_accumulate<3, NRG, KC, RB>(acc_i, W, ldw, D, unit0, unit_end, 3 * D, sh.xrows, nk * D, smem, NoHook());
    }
    if (s > 0) {
        if (tid < RB) {
            const int r = row0 + tid;
            const float* ptr = nullptr;
            if (r < rows) {
                const int b = r / E, e = r - b * E;
                ptr = hx + ((size_t)(b * T + tprev) * E + e) * 2 * D + dir * D;
            }
            sh.xrows[tid] = ptr;
        }
        __syncthreads();
        tile_accumulate<3, NRG, KC, RB>(acc_h, is_h ? P.whh_h[dir] : P.whh_o[dir], D, D, unit0, unit_end, 3 * D,
                                        sh.xrows, D, smem, NoHook());
    }
    const int j = (tid & 15) + 16 * (tid / (16 * NRG));
    const int rg = (tid >> 4) % NRG;
    const int unit = unit0 + j;
    if (unit < unit_end) {
        const float* bhh = is_h ? P.bhh_h[dir] : P.bhh_o[dir];
        const float br = __ldg(bhh + unit), bz = __ldg(bhh + D + unit), bn = __ldg(bhh + 2 * D + unit);
        const float* gsb = is_h ? P.gs_h : P.gs_o;
        const float* ub_ = is_h ? P.u_h : P.u_o;
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
            const int r = row0 + 2 * rg + rr;
            if (r >= rows) continue;
            const int b = r / E, e = r - b * E;
            const size_t fe = (size_t)(b * T + t) * E + e;
            const float* gs = gsb + (fe * 2 + dir) * 3 * D;
            const float hprev = s > 0 ? ld_cg(hx + ((size_t)(b * T + tprev) * E + e) * 2 * D + dir * D + unit) : 0.0f;
            const float hnew = gru_update(__ldg(gs + unit) + acc_i[0][rr], __ldg(gs + D + unit) + acc_i[1][rr],
                                          __ldg(gs + 2 * D + unit) + acc_i[2][rr], acc_h[0][rr] + br, acc_h[1][rr] + bz,
                                          acc_h[2][rr] + bn, hprev);
            const float u = __ldg(ub_ + fe);
            hx[fe * 2 * D + dir * D + unit] = u * hnew + (1.0f - u) * hprev;
        }
    }
}

template <int KC>
__device__ __forceinline__ void seg_cell_dispatch(const SegParams& P, int tile, int s, float* smem, SegShared& sh) {
    const int dir = tile / P.cell_tiles_dir;
    int rem = tile - dir * P.cell_tiles_dir;
    const bool is_h = rem < P.cell_tiles_h_dir;
    if (!is_h) rem -= P.cell_tiles_h_dir;
    const int nub = is_h ? P.nub_h : P.nub_o;
    const int rb = rem / nub, ub = rem - rb * nub;
    const int nrg = is_h ? P.nrg_h : P.nrg_o;
    switch (nrg) {
        case 4:  seg_cell_tile<4, KC>(P, is_h, dir, rb, ub, s, smem, sh); break;
        case 8:  seg_cell_tile<8, KC>(P, is_h, dir, rb, ub, s, smem, sh); break;
        default: seg_cell_tile<16, KC>(P, is_h, dir, rb, ub, s, smem, sh); break;
    }
}

// phases: bit 0 = A (messages), bit 1 = B (cells)
template <int KC>
__global__ void __launch_bounds__(REC_THREADS) segment_kernel(const SegParams P, int s_begin, int s_end, int phases,
                                                             int persistent) {
    extern __shared__ __align__(16) float smem[];
    __shared__ SegShared sh;
    if (threadIdx.x == 0) sh.s_fail = 0;
    unsigned int epoch = 0;
    for (int s = s_begin; s < s_end; ++s) {
        if (phases & 1) {
            for (int tile = blockIdx.x; tile < P.tilesA; tile += gridDim.x) seg_message_tile<KC>(P, tile, s, smem, sh);
            if (persistent && !grid_barrier(P.sync, epoch, gridDim.x, &sh.s_fail)) return;
        }
        if (phases & 2) {
            for (int tile = blockIdx.x; tile < P.tilesB; tile += gridDim.x) seg_cell_dispatch<KC>(P, tile, s, smem, sh);
            if (persistent && s + 1 < s_end && !grid_barrier(P.sync, epoch, gridDim.x, &sh.s_fail)) return;
        }
    }
}

static int pick_nrg(int rows) { return rows <= 8 ? 4 : (rows <= 16 ? 8 : 16); }

static size_t segment_smem_bytes(int KC) {
    size_t m = (size_t)tile_smem_floats<3, MSG_NRG, 32, 64>();   // message tiles (KC<=32)
    const int nrgs[3] = {4, 8, 16};
    for (int i = 0; i < 3; ++i) {
        const size_t f = 2 * (size_t)(3 * (REC_THREADS / nrgs[i]) + 2 * nrgs[i]) * (KC + 4);
        if (f > m) m = f;
    }
    return m * sizeof(float);
}

int launch_segment(SegParams& P, int persistent, cudaStream_t stream) {
    const int D = P.D, B = P.B, H = P.H, O = P.O;
    TG_REQUIRE(D % 16 == 0, "segment: hidden_size=%d must be a multiple of 16", D);
    TG_REQUIRE(O >= 2, "segment: objects->object messages need at least 2 object slots (got %d)", O);
    TG_REQUIRE(!P.hh || H >= 2, "segment: humans->human messages need at least 2 humans (got %d)", H);
    const int maxE = H > O ? H : O;
    TG_REQUIRE(maxE <= 16, "segment: at most 16 entities per type supported (got %d)", maxE);
    const int KC = (D % 32 == 0) ? 32 : 16;
    P.nk_h = P.hh ? 2 : 1;
    P.bbv = 32 / maxE;
    P.n_vb = cdiv(B, P.bbv);
    TG_REQUIRE(P.bbv * maxE * maxE <= MSG_MAXPAIRS, "segment: too many attention pairs per tile");
    const int nub_msg = cdiv(D, MSG_UNITS);
    int begin = 0;
    for (int k = 0; k < 4; ++k) {
        P.msg_tile_begin[k] = begin;
        P.msg_tiles_kind[k] = (k == 0 && !P.hh) ? 0 : P.n_vb * nub_msg;
        begin += P.msg_tiles_kind[k];
    }
    // kinds with zero tiles must never win the "rem >= begin" search: give them an unreachable begin
    if (!P.hh) P.msg_tile_begin[0] = 0;
    P.msg_tile_begin[4] = begin;
    P.msg_tiles_dir = begin;
    P.tilesA = 2 * begin;

    auto kern = KC == 32 ? segment_kernel<32> : segment_kernel<16>;
    const size_t smem = segment_smem_bytes(KC);
    TG_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    TG_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, REC_THREADS, smem));
    TG_REQUIRE(per_sm >= 1, "segment: kernel does not fit on an SM (smem %zu)", smem);
    const int sms = num_sms();
    const int capacity = per_sm * sms;

    // cell tiles: equalise cost (rows x units) across the human and object cells, one wave per step
    P.nrg_h = pick_nrg(B * H);
    P.nrg_o = pick_nrg(B * O);
    const int RBh = 2 * P.nrg_h, RBo = 2 * P.nrg_o, Jh = REC_THREADS / P.nrg_h, Jo = REC_THREADS / P.nrg_o;
    auto clampj = [&](int v, int J) { if (v < 1) v = 1; if (v > J) v = J; if (v > D) v = D; return v; };
    int jh = clampj(Jh, Jh), jo = clampj(Jo, Jo);
    for (int tau = 8; tau <= 64 * 64; ++tau) {
        const int a = clampj(tau / RBh, Jh), b = clampj(tau / RBo, Jo);
        const int tiles = 2 * (cdiv(B * H, RBh) * cdiv(D, a) + cdiv(B * O, RBo) * cdiv(D, b));
        if (tiles <= sms) { jh = a; jo = b; break; }
    }
    P.jeff_h = jh; P.jeff_o = jo;
    P.nrb_h = cdiv(B * H, RBh); P.nub_h = cdiv(D, jh);
    P.nrb_o = cdiv(B * O, RBo); P.nub_o = cdiv(D, jo);
    P.cell_tiles_h_dir = P.nrb_h * P.nub_h;
    P.cell_tiles_dir = P.cell_tiles_h_dir + P.nrb_o * P.nub_o;
    P.tilesB = 2 * P.cell_tiles_dir;

    if (persistent) {
        int grid = P.tilesA > P.tilesB ? P.tilesA : P.tilesB;
        if (grid > capacity) grid = capacity;
        TG_CUDA_OK(cudaMemsetAsync(P.sync.counter, 0, 2 * sizeof(unsigned int), stream));
        int s0 = 0, s1 = P.T, phases = 3, pers = 1;
        void* args[] = {(void*)&P, (void*)&s0, (void*)&s1, (void*)&phases, (void*)&pers};
        TG_CUDA_OK(cudaLaunchCooperativeKernel((const void*)kern, dim3(grid), dim3(REC_THREADS), args, smem, stream));
        ++g_launches;
    } else {
        for (int s = 0; s < P.T; ++s) {
            kern<<<P.tilesA, REC_THREADS, smem, stream>>>(P, s, s + 1, 1, 0);
            TG_LAUNCH_OK();
            kern<<<P.tilesB, REC_THREADS, smem, stream>>>(P, s, s + 1, 2, 0);
            TG_LAUNCH_OK();
        }
    }
    return 0;
}

}  // namespace tg
