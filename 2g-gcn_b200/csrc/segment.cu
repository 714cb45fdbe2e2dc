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

constexpr int MSG_NG = 4;                    // message tiles: 16 sender rows x (4 x 16) units ...
constexpr int MSG_NGL = MSG_NG + 1;          // ... plus one group whose "weight rows" are the receivers' states (logits)
constexpr int MSG_ROWS = REC_RB;             // sender rows per message tile (whole videos)
constexpr int MSG_UNITS = MSG_NG * REC_J;
constexpr int MSG_LDM = MSG_UNITS + 1;
constexpr int MSG_MAXPAIRS = REC_RB * REC_J;

struct SegShared {
    const float* wrows[MSG_NGL * REC_J];
    const float* xrows[REC_RB];
    float msg[MSG_ROWS * MSG_LDM];
    float logit[MSG_MAXPAIRS];
    float alpha[MSG_MAXPAIRS];
    int s_fail;
};

// ---- phase A ------------------------------------------------------------------------------------
__device__ __forceinline__ void seg_message_tile(const SegParams& P, int tile, int s, float* smem, SegShared& sh) {
    const int D = P.D, T = P.T, B = P.B, H = P.H, O = P.O;
    const int dir = tile / P.msg_tiles_dir;
    int rem = tile - dir * P.msg_tiles_dir;
    int kind = 0;
#pragma unroll
    for (int k = 1; k < 4; ++k)
        if (rem >= P.msg_tile_begin[k]) kind = k;
    rem -= P.msg_tile_begin[kind];
    const int nub = P.msg_tiles_kind[kind] / P.n_vb[kind];
    const int vb = rem / nub, ub = rem - vb * nub;
    const bool send_h = (kind == 0 || kind == 2);       // sender type: humans for hh, ho
    const bool recv_h = (kind == 0 || kind == 1);       // receiver type: humans for hh, oh
    const int Es = send_h ? H : O, Er = recv_h ? H : O;
    const bool same = (send_h == recv_h);
    const int b0 = vb * P.bbv[kind], nb = min(P.bbv[kind], B - b0);
    const int unit0 = ub * MSG_UNITS;
    const int t = dir == 0 ? s : T - 1 - s;
    const int tprev = dir == 0 ? t - 1 : t + 1;
    const int tid = threadIdx.x;

    __syncthreads();
    if (tid < MSG_UNITS) {                                   // message MLP rows of this unit slice
        const int unit = unit0 + tid;
        sh.wrows[tid] = unit < D ? P.wm[kind] + (size_t)unit * D : nullptr;
    } else if (tid < MSG_UNITS + REC_J) {                    // receivers' previous states: the logit "weights"
        const int l = tid - MSG_UNITS, bl = l / Er, e = l - bl * Er;
        const float* ptr = nullptr;
        if (s > 0 && bl < nb) {
            const float* base = recv_h ? P.hx_h : P.hx_o;
            ptr = base + ((size_t)((b0 + bl) * T + tprev) * Er + e) * 2 * D + dir * D;
        }
        sh.wrows[tid] = ptr;
    } else if (tid < MSG_UNITS + REC_J + MSG_ROWS) {         // senders' previous states
        const int l = tid - MSG_UNITS - REC_J, bl = l / Es, e = l - bl * Es;
        const float* ptr = nullptr;
        if (s > 0 && bl < nb) {
            const float* base = send_h ? P.hx_h : P.hx_o;
            ptr = base + ((size_t)((b0 + bl) * T + tprev) * Es + e) * 2 * D + dir * D;
        }
        sh.xrows[l] = ptr;
    }
    __syncthreads();

    float acc[MSG_NGL];
    tile_accumulate<MSG_NGL, 3>(acc, sh.wrows, sh.xrows, s > 0 ? D : 0, smem);

    // thread pair: unit (or receiver) index = tid % 16, sender row = tid / 16
    if (tid < MSG_ROWS * REC_J) {
        const int j = tid & 15, row = tid >> 4;
#pragma unroll
        for (int g = 0; g < MSG_NG; ++g) {
            const int c = g * REC_J + j, u = unit0 + c;
            const float bias = u < D ? __ldg(P.bm[kind] + u) : 0.0f;
            sh.msg[row * MSG_LDM + c] = fmaxf(acc[g] + bias, 0.0f);
        }
        // logit of (receiver j, sender row) when both belong to the same video of the block
        const int blr = j / Er, r = j - blr * Er, bls = row / Es, sdr = row - bls * Es;
        if (blr == bls && blr < nb) sh.logit[(blr * Er + r) * Es + sdr] = acc[MSG_NG] * (1.0f / sqrtf((float)D));
    }
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
__device__ __forceinline__ void seg_cell_tile(const SegParams& P, int tile, int s, float* smem, SegShared& sh) {
    const int D = P.D, T = P.T, B = P.B;
    const int dir = tile / P.cell_tiles_dir;
    int rem = tile - dir * P.cell_tiles_dir;
    const bool is_h = rem < P.cell_tiles_h_dir;
    if (!is_h) rem -= P.cell_tiles_h_dir;
    const int nub = is_h ? P.nub_h : P.nub_o;
    const int rb = rem / nub, ub = rem - rb * nub;
    const int E = is_h ? P.H : P.O;
    const int rows = B * E;
    const int row0 = rb * REC_RB, unit0 = ub * REC_J;
    const int t = dir == 0 ? s : T - 1 - s;
    const int tprev = dir == 0 ? t - 1 : t + 1;
    const int tid = threadIdx.x;
    const int nk = is_h ? P.nk_h : 2;
    float* hx = is_h ? P.hx_h : P.hx_o;
    const float* mgbase = is_h ? P.mg_h : P.mg_o;

    __syncthreads();
    if (tid < 3 * REC_J) {           // segment-message columns of W_ih
        const int g = tid / REC_J, unit = unit0 + tid % REC_J;
        const float* W = is_h ? P.wih_h[dir] + P.col_h : P.wih_o[dir] + P.col_o;
        const int ldw = is_h ? P.ldw_h : P.ldw_o;
        sh.wrows[tid] = unit < D ? W + (size_t)(g * D + unit) * ldw : nullptr;
    } else if (tid < 3 * REC_J + REC_RB) {
        const int r = row0 + tid - 3 * REC_J;
        sh.xrows[tid - 3 * REC_J] = r < rows ? mgbase + ((size_t)dir * rows + r) * nk * D : nullptr;
    }
    __syncthreads();
    float acc_i[3], acc_h[3];
    tile_accumulate<3, 4>(acc_i, sh.wrows, sh.xrows, nk * D, smem);
    if (tid < 3 * REC_J) {           // W_hh
        const int g = tid / REC_J, unit = unit0 + tid % REC_J;
        const float* W = is_h ? P.whh_h[dir] : P.whh_o[dir];
        sh.wrows[tid] = unit < D ? W + (size_t)(g * D + unit) * D : nullptr;
    } else if (tid < 3 * REC_J + REC_RB) {
        const int r = row0 + tid - 3 * REC_J;
        const float* ptr = nullptr;
        if (r < rows && s > 0) {
            const int b = r / E, e = r - b * E;
            ptr = hx + ((size_t)(b * T + tprev) * E + e) * 2 * D + dir * D;
        }
        sh.xrows[tid - 3 * REC_J] = ptr;
    }
    __syncthreads();
    tile_accumulate<3, 4>(acc_h, sh.wrows, sh.xrows, s > 0 ? D : 0, smem);

    const int unit = unit0 + (tid & 15), lr = tid >> 4, r = row0 + lr;
    if (lr < REC_RB && unit < D && r < rows) {
        const float* bhh = is_h ? P.bhh_h[dir] : P.bhh_o[dir];
        const float* gsb = is_h ? P.gs_h : P.gs_o;
        const float* ub_ = is_h ? P.u_h : P.u_o;
        const int b = r / E, e = r - b * E;
        const size_t fe = (size_t)(b * T + t) * E + e;
        const float* gs = gsb + (fe * 2 + dir) * 3 * D;
        const float hprev = s > 0 ? ld_cg(hx + ((size_t)(b * T + tprev) * E + e) * 2 * D + dir * D + unit) : 0.0f;
        const float hnew = gru_update(__ldg(gs + unit) + acc_i[0], __ldg(gs + D + unit) + acc_i[1],
                                      __ldg(gs + 2 * D + unit) + acc_i[2], acc_h[0] + __ldg(bhh + unit),
                                      acc_h[1] + __ldg(bhh + D + unit), acc_h[2] + __ldg(bhh + 2 * D + unit), hprev);
        const float u = __ldg(ub_ + fe);
        hx[fe * 2 * D + dir * D + unit] = u * hnew + (1.0f - u) * hprev;
    }
}

// phases: bit 0 = A (messages), bit 1 = B (cells)
__global__ void __launch_bounds__(REC_THREADS, 1) segment_kernel(const SegParams P, int s_begin, int s_end, int phases,
                                                                int persistent) {
    extern __shared__ __align__(16) float smem[];
    __shared__ SegShared sh;
    if (threadIdx.x == 0) sh.s_fail = 0;
    unsigned int epoch = 0;
    for (int s = s_begin; s < s_end; ++s) {
        if (phases & 1) {
            for (int tile = blockIdx.x; tile < P.tilesA; tile += gridDim.x) seg_message_tile(P, tile, s, smem, sh);
            if (persistent && !grid_barrier(P.sync, epoch, gridDim.x, &sh.s_fail)) return;
        }
        if (phases & 2) {
            for (int tile = blockIdx.x; tile < P.tilesB; tile += gridDim.x) seg_cell_tile(P, tile, s, smem, sh);
            if (persistent && s + 1 < s_end && !grid_barrier(P.sync, epoch, gridDim.x, &sh.s_fail)) return;
        }
    }
}

int launch_segment(SegParams& P, int persistent, cudaStream_t stream) {
    const int D = P.D, B = P.B, H = P.H, O = P.O;
    TG_REQUIRE(D % 16 == 0, "segment: hidden_size=%d must be a multiple of 16", D);
    TG_REQUIRE(O >= 2, "segment: objects->object messages need at least 2 object slots (got %d)", O);
    TG_REQUIRE(!P.hh || H >= 2, "segment: humans->human messages need at least 2 humans (got %d)", H);
    const int maxE = H > O ? H : O;
    TG_REQUIRE(maxE <= 16, "segment: at most 16 entities per type supported (got %d)", maxE);
    P.nk_h = P.hh ? 2 : 1;
    const int nub_msg = cdiv(D, MSG_UNITS);
    int begin = 0;
    for (int k = 0; k < 4; ++k) {
        const int Es = (k == 0 || k == 2) ? H : O, Er = (k == 0 || k == 1) ? H : O;
        const int me = Es > Er ? Es : Er;
        P.bbv[k] = MSG_ROWS / me;                       // whole videos per message tile (senders and receivers fit)
        P.n_vb[k] = cdiv(B, P.bbv[k]);
        P.msg_tile_begin[k] = begin;
        P.msg_tiles_kind[k] = (k == 0 && !P.hh) ? 0 : P.n_vb[k] * nub_msg;
        begin += P.msg_tiles_kind[k];
    }
    if (!P.hh) P.msg_tile_begin[0] = 0;
    P.msg_tile_begin[4] = begin;
    P.msg_tiles_dir = begin;
    P.tilesA = 2 * begin;

    auto kern = segment_kernel;
    const int fa = tile_smem_floats(3, 4), fb = tile_smem_floats(MSG_NGL, 3);
    const size_t smem = sizeof(float) * (size_t)(fa > fb ? fa : fb);
    static bool configured = false;
    if (!configured) {
        TG_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    int per_sm = 0;
    TG_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, REC_THREADS, smem));
    TG_REQUIRE(per_sm >= 1, "segment: kernel does not fit on an SM (smem %zu)", smem);
    const int capacity = per_sm * num_sms();

    P.cfg_h = P.cfg_o = 1;
    P.jeff_h = P.jeff_o = REC_J;
    P.nrb_h = cdiv(B * H, REC_RB); P.nub_h = cdiv(D, REC_J);
    P.nrb_o = cdiv(B * O, REC_RB); P.nub_o = cdiv(D, REC_J);
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
