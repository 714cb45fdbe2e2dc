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
#include <stdlib.h>
#include "recurrent.cuh"
#include "recurrent_res.cuh"
#include "bigru.h"
#include "step_tc.cuh"

namespace tg {

constexpr int MSG_NG = 4;                    // message tiles: 16 sender rows x (4 x 16) units ...
constexpr int MSG_NGL = MSG_NG + 1;          // ... plus one group whose "weight rows" are the receivers' states (logits)
constexpr int MSG_ROWS = 16;                 // sender rows per message tile (whole videos)
constexpr int MSG_UNITS = MSG_NG * REC_J;
constexpr int MSG_LDM = MSG_UNITS + 1;
constexpr int MSG_MAXPAIRS = MSG_ROWS * REC_J;

struct SegShared {
    const float* tab1[MSG_NGL * REC_J + 32];     // weight rows then activation rows, K segment 1
    const float* tab2[MSG_NGL * REC_J + 32];     // K segment 2 (cell tiles: W_hh against the previous state)
    float om[MSG_ROWS];                          // objects_mask of the sender rows of a message tile
    float msg[MSG_ROWS * MSG_LDM];
    float logit[MSG_MAXPAIRS];
    float alpha[MSG_MAXPAIRS];
    // step plan of the message tile's epilogue, written by the set-up part (no integer divisions on the compute path)
    int pair[MSG_MAXPAIRS];                      // (receiver, sender) pair of thread tid: receiver | sender << 8 | allowed << 16, or -1
    int lidx[MSG_MAXPAIRS];                      // epilogue thread (receiver tid % 16, sender row tid / 16): index into logit[], or -1
    unsigned int rcv_mask[MSG_ROWS];             // allowed senders of a receiver (bit q)
    int rcv_msg[MSG_ROWS];                       // offset of the receiver's video's first sender row in msg[]
    float* rcv_mg[MSG_ROWS];                     // where the receiver's aggregated message slice goes
    float* rcv_att[MSG_ROWS];                    // attention outputs of the receiver (inspect / saved for the backward), or null
    float* rcv_sal[MSG_ROWS];
    float* row_smsg[MSG_ROWS];                   // saved post-ReLU message slice of a sender row, or null
    int s_fail;
};

// ---- phase A ------------------------------------------------------------------------------------
// MODE: 0 = streaming mma.sync tiles, 2 = streaming message tiles + cell tiles with
// on-chip resident weights (recurrent_res.cuh)
// part: bit 0 = set-up (pointer tables; depends on nothing another CTA writes during the step), bit 1 = compute.  The persistent
// resident variant runs the set-up of the NEXT phase between grid_arrive and grid_wait; everything else passes part = 3.
// vbi: iteration over the video blocks a resident tile owns (video block = first block + vbi * groups); returns false
// (uniformly) when the tile has no such block.
template <int MODE>
__device__ __forceinline__ bool seg_message_tile(const SegParams& P, int tile, int s, float* smem, SegShared& sh,
                                                 uint4* wmsg, int& msg_ready, int part, int vbi = 0) {
    const int D = P.D, T = P.T, B = P.B, H = P.H, O = P.O;
    const int dir = tile / P.msg_tiles_dir;
    int rem = tile - dir * P.msg_tiles_dir;
    int kind = 0;
#pragma unroll
    for (int k = 1; k < 4; ++k)
        if (rem >= P.msg_tile_begin[k]) kind = k;
    rem -= P.msg_tile_begin[kind];
    const int nub = P.msg_tiles_kind[kind] / P.msg_g[kind];
    const int vbg = rem / nub, ub = rem - vbg * nub;
    const int vb = vbg + vbi * P.msg_g[kind];
    if (vb >= P.n_vb[kind]) return false;
    const bool send_h = (kind == 0 || kind == 2);       // sender type: humans for hh, ho
    const bool recv_h = (kind == 0 || kind == 1);       // receiver type: humans for hh, oh
    const int Es = send_h ? H : O, Er = recv_h ? H : O;
    const bool same = (send_h == recv_h);
    const int b0 = vb * P.bbv[kind], nb = min(P.bbv[kind], B - b0);
    const int unit0 = ub * MSG_UNITS;
    const int t = dir == 0 ? s : T - 1 - s;
    const int tprev = dir == 0 ? t - 1 : t + 1;
    const int tid = threadIdx.x;

    if (part == 3) __syncthreads();
    if (!(part & 1)) {
    } else if (tid < MSG_UNITS) {                            // message MLP rows of this unit slice
        const int unit = unit0 + tid;
        sh.tab1[tid] = unit < D ? P.wm[kind] + (size_t)unit * D : nullptr;
    } else if (tid < MSG_UNITS + REC_J) {                    // receivers' previous states: the logit "weights"
        const int l = tid - MSG_UNITS, bl = l / Er, e = l - bl * Er;
        const float* ptr = nullptr;
        if (s > 0 && bl < nb) {
            const float* base = recv_h ? P.hx_h : P.hx_o;
            ptr = base + ((size_t)((b0 + bl) * T + tprev) * Er + e) * 2 * D + dir * D;
        }
        sh.tab1[tid] = ptr;
    } else if (tid < MSG_UNITS + REC_J + MSG_ROWS) {         // senders' previous states
        const int l = tid - MSG_UNITS - REC_J, bl = l / Es, e = l - bl * Es;
        const float* ptr = nullptr;
        if (s > 0 && bl < nb) {
            const float* base = send_h ? P.hx_h : P.hx_o;
            ptr = base + ((size_t)((b0 + bl) * T + tprev) * Es + e) * 2 * D + dir * D;
        }
        sh.tab1[tid] = ptr;
        sh.om[l] = (!send_h && bl < nb) ? __ldg(P.om + (b0 + bl) * O + e) : 1.0f;
    }
    if (part & 1) {
        const int nk_r = recv_h ? P.nk_h : 2;
        const int slot = recv_h ? (kind == 0 ? 0 : P.nk_h - 1) : kind - 2;
        if (tid < MSG_MAXPAIRS) {
            int info = -1;
            if (tid < nb * Er * Es) {
                const int br = tid / Es, sdr = tid - br * Es, bl = br / Er, r = br - bl * Er;
                const bool ok = !(same && sdr == r) && (send_h || __ldg(P.om + (b0 + bl) * O + sdr) != 0.0f);
                info = br | (sdr << 8) | ((int)ok << 16);
            }
            sh.pair[tid] = info;
            const int j = tid & 15, row = tid >> 4;
            const int blr = j / Er, r = j - blr * Er, bls = row / Es, sdr = row - bls * Es;
            sh.lidx[tid] = (blr == bls && blr < nb) ? (blr * Er + r) * Es + sdr : -1;
        }
        if (tid < MSG_ROWS) {                                    // receiver tid of the block
            const int br = tid, bl = br / Er, r = br - bl * Er, b = b0 + bl;
            unsigned int mask = 0;
            float* mgp = nullptr;
            float* attp = nullptr;
            float* salp = nullptr;
            if (br < nb * Er) {
                for (int q = 0; q < Es; ++q)
                    if (!(same && q == r) && (send_h || __ldg(P.om + b * O + q) != 0.0f)) mask |= 1u << q;
                float* mg = recv_h ? P.mg_h : P.mg_o;
                mgp = mg + ((((size_t)dir * B + b) * P.mg_T + (P.mg_T > 1 ? t : 0)) * Er + r) * nk_r * D + slot * D + unit0;
                if (ub == 0) {
                    float* att = dir == 0 ? P.att_f : P.att_b;
                    if (kind == 1 && att != nullptr) attp = att + ((size_t)(b * H + r) * T + t) * O;
                    if (P.salpha[kind] != nullptr) salp = P.salpha[kind] + ((((size_t)dir * B + b) * T + t) * Er + r) * Es;
                }
            }
            sh.rcv_mask[br] = mask;
            sh.rcv_msg[br] = bl * Es * MSG_LDM;
            sh.rcv_mg[br] = mgp;
            sh.rcv_att[br] = attp;
            sh.rcv_sal[br] = salp;
        } else if (tid >= 32 && tid < 32 + MSG_ROWS) {           // sender row tid - 32
            const int row = tid - 32, bls = row / Es, sdr = row - bls * Es;
            float* ptr = nullptr;
            if (P.smsg[kind] != nullptr && bls < nb) ptr = P.smsg[kind] + ((((size_t)dir * B + b0 + bls) * T + t) * Es + sdr) * D + unit0;
            sh.row_smsg[row] = ptr;
        }
    }
    if (!(part & 2)) return true;
    // bias of this thread's message columns, fetched before the K loop
    float bias[MSG_NG];
#pragma unroll
    for (int g = 0; g < MSG_NG; ++g) {
        const int u = unit0 + g * REC_J + (tid & 15);
        bias[g] = u < D ? __ldg(P.bm[kind] + u) : 0.0f;
    }
    if (part == 3) __syncthreads();

    float acc[MSG_NGL][1];
    if (MODE == 2 && P.res_msg) {
        if (!msg_ready) { res_fill_msg(wmsg, sh.tab1, D, P.sync.error); msg_ready = 1; }   // first step: this CTA's message weights go on chip
        tile_accumulate_msg_res(acc, sh.tab1 + MSG_UNITS, s > 0 ? D : 0, wmsg, P.wm[kind], smem);
    } else {
        tile_accumulate<MSG_NGL, 2, 3>(acc, sh.tab1, sh.tab1, s > 0 ? D : 0, 0, 0u, 0u, P.wm[kind], smem);
    }

    // thread pair: unit (or receiver) index = tid % 16, sender row = tid / 16
    if (tid < MSG_ROWS * REC_J) {
        const int j = tid & 15, row = tid >> 4;
        float* sv = sh.row_smsg[row];
#pragma unroll
        for (int g = 0; g < MSG_NG; ++g) {
            const int c = g * REC_J + j;
            const float mv = fmaxf(acc[g][0] + bias[g], 0.0f);
            sh.msg[row * MSG_LDM + c] = mv;
            if (sv != nullptr && unit0 + c < D) sv[c] = mv;
        }
        // logit of (receiver j, sender row) when both belong to the same video of the block
        const int li = sh.lidx[tid];
        if (li >= 0) {
            float lg = acc[MSG_NG][0] * (P.att_noscale ? 1.0f : 1.0f / sqrtf((float)D));
            if (!P.mean_pool && dist_of_kind(P.dist, kind) != nullptr) {     // distance-based weights (models.py:1757-1775)
                const int sdr = li % Es, br = li / Es, bl = br / Er, r = br - bl * Er;
                bool dv;
                dist_logit(P.dist, kind, (size_t)(b0 + bl) * T + t, H, O, r, sdr, lg, dv);
                if (!dv) lg = -INFINITY;                                     // a zero distance masks the sender
            }
            sh.logit[li] = lg;
        }
    }
    __syncthreads();
    // masked softmax over the senders of each receiver (vhoi/models.py:1750-1753), one thread per (receiver, sender) pair:
    // un-normalised weight exp(l - max) into shared memory; the pair threads of unit block 0 also publish the normalised weights
    if (tid < MSG_MAXPAIRS && sh.pair[tid] >= 0) {
        const int info = sh.pair[tid], br = info & 255, sdr = (info >> 8) & 255;
        const unsigned int mask = sh.rcv_mask[br];
        const float* lrow = sh.logit + br * Es;
        float m = -INFINITY;
        for (int q = 0; q < Es; ++q)
            if ((mask >> q) & 1u) m = fmaxf(m, lrow[q]);
        const float ex = ((info >> 16) && (P.mean_pool || lrow[sdr] > -INFINITY)) ? (P.mean_pool ? 1.0f : expf(lrow[sdr] - m)) : 0.0f;   // 'mp': weight 1 / #valid senders
        sh.alpha[tid] = ex;
        float* attp = sh.rcv_att[br];
        float* salp = sh.rcv_sal[br];
        if (attp != nullptr || salp != nullptr) {            // the sum over this receiver's senders in sender order
            float sum = 0.0f;
            for (int q = 0; q < Es; ++q) sum += (((mask >> q) & 1u) && (P.mean_pool || lrow[q] > -INFINITY)) ? (P.mean_pool ? 1.0f : expf(lrow[q] - m)) : 0.0f;
            const float a = ex * (sum > 0.0f ? 1.0f / sum : 0.0f);
            if (attp != nullptr) attp[sdr] = a;
            if (salp != nullptr) salp[sdr] = a;
        }
    }
    __syncthreads();
    // aggregated message for every receiver of the block: unit column tid % 64, receivers tid / 64 + 4 i
    if (tid < REC_THREADS) {
        const int c = tid & (MSG_UNITS - 1);
        const bool uok = unit0 + c < D;
        for (int br = tid / MSG_UNITS; br < nb * Er; br += REC_THREADS / MSG_UNITS) {
            const float* al = sh.alpha + br * Es;
            const float* ms = sh.msg + sh.rcv_msg[br] + c;
            float sum = 0.0f;
            for (int sdr = 0; sdr < Es; ++sdr) sum += al[sdr];
            const float inv = sum > 0.0f ? 1.0f / sum : 0.0f;
            float v = 0.0f;
            for (int sdr = 0; sdr < Es; ++sdr) v = fmaf(al[sdr] * inv, ms[sdr * MSG_LDM], v);
            if (MODE == 2 && !(v < RES_F16_MAX)) atomicOr(P.sync.error, 2u);      // operand of the fp16-split cell tile out of range
            if (uok) sh.rcv_mg[br][c] = v;
        }
    }
    return true;
}

// ---- phase B ------------------------------------------------------------------------------------
// One pipeline over the concatenated K range [segment-message columns of W_ih | W_hh] with four weight groups:
// r and z accumulate over both segments, n_i only over the first, n_h only over the second (GRU needs them apart).
// Epilogue operands of a cell tile, fetched by the set-up part and consumed after the K loop.
struct CellPre {
    float bh[3], xg[2][3], hprev[2], ug[2];
    bool valid[2];
    size_t orow[2];
    float* gsave[2];
};

// part: as for seg_message_tile (bit 0 = pointer tables + epilogue operands, bit 1 = K loop + gate math).
template <int NT, int MODE>
__device__ __forceinline__ void seg_cell_tile(const SegParams& P, bool is_h, int dir, int rb, int ub, int s, float* smem,
                                              SegShared& sh, ResState& res, int part, CellPre& pre) {
    constexpr int RBT = 8 * NT, NPAIR = NT / 2, WR = 4 * REC_J;
    const int D = P.D, T = P.T, B = P.B;
    const int E = is_h ? P.H : P.O;
    const int rows = B * E;
    const int row0 = rb * RBT, unit0 = ub * REC_J;
    const int t = dir == 0 ? s : T - 1 - s;
    const int tprev = dir == 0 ? t - 1 : t + 1;
    const int tid = threadIdx.x;
    const int nk = is_h ? P.nk_h : 2;
    float* hx = is_h ? P.hx_h : P.hx_o;
    const float* Wh = is_h ? P.whh_h[dir] : P.whh_o[dir];

    if (part == 3) __syncthreads();
    if (part & 1) {
        const float* mgbase = is_h ? P.mg_h : P.mg_o;
        const float* Wi = is_h ? P.wih_h[dir] + P.col_h : P.wih_o[dir] + P.col_o;
        const int ldw = is_h ? P.ldw_h : P.ldw_o;
        if (tid < WR) {
            if (MODE != 2 || !res.ready) {                           // resident weights: the row pointers are only needed for the fill
                const int g = tid / REC_J, unit = unit0 + tid % REC_J;   // groups: 0 r, 1 z, 2 n (input part), 3 n (hidden part)
                const int gate = g < 3 ? g : 2;
                const bool ok = unit < D;
                sh.tab1[tid] = (ok && g < 3) ? Wi + (size_t)(gate * D + unit) * ldw : nullptr;
                sh.tab2[tid] = (ok && g != 2) ? Wh + (size_t)(gate * D + unit) * D : nullptr;
            }
        } else if (tid < WR + RBT) {
            const int r = row0 + tid - WR;
            const float* p1 = nullptr;
            const float* p2 = nullptr;
            if (r < rows) {
                const int b = r / E, e = r - b * E;
                p1 = mgbase + ((((size_t)dir * B + b) * P.mg_T + (P.mg_T > 1 ? t : 0)) * E + e) * nk * D;
                if (s > 0) {
                    p2 = hx + ((size_t)(b * T + tprev) * E + e) * 2 * D + dir * D;
                }
            }
            sh.tab1[tid] = p1;
            sh.tab2[tid] = p2;
        }
        // epilogue operands, fetched before the K loop
        const int unit = unit0 + (tid & 15);
        const float* bhh = is_h ? P.bhh_h[dir] : P.bhh_o[dir];
        const float* gsb = is_h ? P.gs_h : P.gs_o;
        const float* ub_ = is_h ? P.u_h : P.u_o;
        float* sg = is_h ? P.sgates_h : P.sgates_o;
        pre.bh[0] = pre.bh[1] = pre.bh[2] = 0.f;
        if (unit < D) { pre.bh[0] = __ldg(bhh + unit); pre.bh[1] = __ldg(bhh + D + unit); pre.bh[2] = __ldg(bhh + 2 * D + unit); }
#pragma unroll
        for (int p = 0; p < NPAIR; ++p) {
            const int lr = (tid >> 4) + 16 * p, r = row0 + lr;
            pre.valid[p] = unit < D && r < rows && lr < RBT && tid < REC_THREADS;
            pre.xg[p][0] = pre.xg[p][1] = pre.xg[p][2] = pre.hprev[p] = pre.ug[p] = 0.0f;
            pre.orow[p] = 0;
            pre.gsave[p] = nullptr;
            if (pre.valid[p]) {
                const int b = r / E, e = r - b * E;
                const size_t fe = (size_t)(b * T + t) * E + e;
                const float* gs = gsb + (fe * 2 + dir) * 3 * D;
                pre.xg[p][0] = __ldg(gs + unit); pre.xg[p][1] = __ldg(gs + D + unit); pre.xg[p][2] = __ldg(gs + 2 * D + unit);
                pre.ug[p] = __ldg(ub_ + fe);
                if (s > 0) pre.hprev[p] = ld_cg(hx + ((size_t)(b * T + tprev) * E + e) * 2 * D + dir * D + unit);
                pre.orow[p] = fe * 2 * D + dir * D + unit;
                if (sg != nullptr) pre.gsave[p] = sg + (fe * 2 + dir) * 4 * D + unit;
            }
        }
    }
    if (!(part & 2)) return;
    if (part == 3) __syncthreads();

    float acc[4][NPAIR];
    if (MODE == 2) {
        if (!res.ready) res_fill_cell(res, sh.tab1, sh.tab2, nk * D, D, P.sync.error);   // first step: this CTA's weight fragments go on chip
        tile_accumulate_res<NT>(acc, sh.tab1 + WR, sh.tab2 + WR, nk * D, D, res, Wh, smem);   // s == 0: null state rows = zeros
    } else {
        tile_accumulate<4, NT, 3>(acc, sh.tab1, sh.tab2, nk * D, s > 0 ? D : 0, 1u << 3, 1u << 2, Wh, smem);
    }

#pragma unroll
    for (int p = 0; p < NPAIR; ++p) {
        if (!pre.valid[p]) continue;
        const float hnew = gru_update(pre.xg[p][0] + acc[0][p], pre.xg[p][1] + acc[1][p], pre.xg[p][2] + acc[2][p], pre.bh[0], pre.bh[1],
                                      acc[3][p] + pre.bh[2], pre.hprev[p], pre.gsave[p], D);
        hx[pre.orow[p]] = pre.ug[p] * hnew + (1.0f - pre.ug[p]) * pre.hprev[p];
    }
}

// rbi: with res_multi a tile is one (direction, entity type, row-block group, unit block) and owns every cell_g-th row block —
// the weights stay resident while rbi walks over them; returns false (uniformly) when there is no such row block.
template <int MODE>
__device__ __forceinline__ bool seg_cell_dispatch(const SegParams& P, int tile, int s, float* smem, SegShared& sh,
                                                  ResState& res, int part, CellPre& pre, int rbi = 0) {
    const int dir = tile / P.cell_tiles_dir;
    int rem = tile - dir * P.cell_tiles_dir;
    const bool is_h = rem < P.cell_tiles_h_dir;
    if (!is_h) rem -= P.cell_tiles_h_dir;
    const int nub = is_h ? P.nub_h : P.nub_o;
    const int rbg = rem / nub, ub = rem - rbg * nub;
    const int rb = rbg + rbi * (is_h ? P.cell_g_h : P.cell_g_o);
    if (rb >= (is_h ? P.nrb_h : P.nrb_o)) return false;
    if ((is_h ? P.cfg_h : P.cfg_o) == 4) seg_cell_tile<4, MODE>(P, is_h, dir, rb, ub, s, smem, sh, res, part, pre);
    else                                 seg_cell_tile<2, MODE>(P, is_h, dir, rb, ub, s, smem, sh, res, part, pre);
    return true;
}

// phases: bit 0 = A (messages), bit 1 = B (cells).  MODE as above.
template <int MODE>
__global__ void __launch_bounds__(REC_THREADS, 1) segment_kernel(const SegParams P, int s_begin, int s_end,
                                                                                          int phases, int persistent) {
    extern __shared__ __align__(16) float smem[];
    __shared__ SegShared sh;
    __shared__ uint32_t tmem_slot;
    ResState res;
    if (threadIdx.x == 0) sh.s_fail = 0;
    // MODE 2: the ring keeps its place at the start of dynamic shared memory; the overflow fragments follow it
    if (MODE == 2) res_init(res, &tmem_slot, reinterpret_cast<uint4*>(smem + P.res_ring_floats));
    // MODE 2 with res_msg: the message weights of this CTA's message tile follow the overflow fragments
    uint4* wmsg = reinterpret_cast<uint4*>(smem + P.res_ring_floats + RES_SMEM_WORDS * REC_THREADS);
    int msg_ready = 0;
    unsigned int epoch = 0;
    bool ok = true;
    CellPre pre;
#ifndef SEG_NO_SHADOW
    if (MODE == 2 && persistent && phases == 3 && P.res_msg && !P.res_multi) {
        // one message tile and one cell tile per CTA for the whole launch (launcher guarantees tilesA, tilesB <= gridDim.x):
        // the set-up of the next phase (pointer tables, epilogue operands) runs in the shadow of the grid barrier
        const int bid = blockIdx.x;
        const bool hasA = bid < P.tilesA, hasB = bid < P.tilesB;
        if (hasA) seg_message_tile<MODE>(P, bid, s_begin, smem, sh, wmsg, msg_ready, 1);
        __syncthreads();
        for (int s = s_begin; s < s_end; ++s) {
            if (hasA) seg_message_tile<MODE>(P, bid, s, smem, sh, wmsg, msg_ready, 2);
            grid_arrive(P.sync, epoch);
            if (hasB) seg_cell_dispatch<MODE>(P, bid, s, smem, sh, res, 1, pre);
            if (!grid_wait(P.sync, epoch, gridDim.x, &sh.s_fail)) break;
            if (hasB) seg_cell_dispatch<MODE>(P, bid, s, smem, sh, res, 2, pre);
            if (s + 1 == s_end) break;
            grid_arrive(P.sync, epoch);
            if (hasA) seg_message_tile<MODE>(P, bid, s + 1, smem, sh, wmsg, msg_ready, 1);
            if (!grid_wait(P.sync, epoch, gridDim.x, &sh.s_fail)) break;
        }
        s_begin = s_end;                        // skip the generic loop
    }
#endif
    if (MODE == 2 && persistent && P.res_multi) {
        // larger batches: a CTA still owns ONE (direction, kind, unit block) message slice and ONE (direction, type, unit block)
        // cell slice with resident weights, and walks over the video blocks / row blocks that share them
        const int bid = blockIdx.x;
        for (int s = s_begin; s < s_end; ++s) {
            if (bid < P.tilesA)
                for (int vbi = 0; seg_message_tile<MODE>(P, bid, s, smem, sh, wmsg, msg_ready, 3, vbi); ++vbi) {}
            if (!grid_barrier(P.sync, epoch, gridDim.x, &sh.s_fail)) break;
            if (bid < P.tilesB)
                for (int rbi = 0; seg_cell_dispatch<MODE>(P, bid, s, smem, sh, res, 3, pre, rbi); ++rbi) {}
            if (s + 1 < s_end && !grid_barrier(P.sync, epoch, gridDim.x, &sh.s_fail)) break;
        }
        s_begin = s_end;
    }
    for (int s = s_begin; s < s_end && ok; ++s) {
        if (phases & 4) {       // timing experiment: two bare grid barriers per step
            if (!grid_barrier(P.sync, epoch, gridDim.x, &sh.s_fail)) { ok = false; break; }
            if (!grid_barrier(P.sync, epoch, gridDim.x, &sh.s_fail)) { ok = false; break; }
        }
        if (phases & 1) {
            for (int tile = blockIdx.x; tile < P.tilesA; tile += gridDim.x)
                seg_message_tile<MODE>(P, tile, s, smem, sh, wmsg, msg_ready, 3);
            if (persistent && !grid_barrier(P.sync, epoch, gridDim.x, &sh.s_fail)) { ok = false; break; }
        }
        if (phases & 2) {
            for (int tile = blockIdx.x; tile < P.tilesB; tile += gridDim.x) seg_cell_dispatch<MODE>(P, tile, s, smem, sh, res, 3, pre);
            if (persistent && s + 1 < s_end && !grid_barrier(P.sync, epoch, gridDim.x, &sh.s_fail)) { ok = false; break; }
        }
    }
    if (MODE == 2) res_finish(res);
}

int launch_segment(SegParams& P, int persistent, cudaStream_t stream) {
    if (P.big_ws != nullptr) return launch_segment_big(P, P.big_ws, P.precision, P.T, stream);
    const int D = P.D, B = P.B, H = P.H, O = P.O;
    TG_REQUIRE(D % 16 == 0, "segment: hidden_size=%d must be a multiple of 16", D);
    TG_REQUIRE(O >= 2, "segment: objects->object messages need at least 2 object slots (got %d)", O);
    TG_REQUIRE(!P.hh || H >= 2, "segment: humans->human messages need at least 2 humans (got %d)", H);
    const int maxE = H > O ? H : O;
    TG_REQUIRE(maxE <= 16, "segment: at most 16 entities per type supported (got %d)", maxE);
    P.nk_h = P.hh ? 2 : 1;
    if (P.mg_T < 1) P.mg_T = 1;
    const int nub_msg = cdiv(D, MSG_UNITS);
    int begin = 0;
    for (int k = 0; k < 4; ++k) {
        const int Es = (k == 0 || k == 2) ? H : O, Er = (k == 0 || k == 1) ? H : O;
        const int me = Es > Er ? Es : Er;
        P.bbv[k] = MSG_ROWS / me;                       // whole videos per message tile (senders and receivers fit)
        P.n_vb[k] = cdiv(B, P.bbv[k]);
        P.msg_tile_begin[k] = begin;
        P.msg_g[k] = P.n_vb[k];
        P.msg_tiles_kind[k] = (k == 0 && !P.hh) ? 0 : P.n_vb[k] * nub_msg;
        begin += P.msg_tiles_kind[k];
    }
    P.res_multi = 0;
    if (!P.hh) P.msg_tile_begin[0] = 0;
    P.msg_tile_begin[4] = begin;
    P.msg_tiles_dir = begin;
    P.tilesA = 2 * begin;

    // tiling of the cell phase (needed to choose the kernel variant)
    P.cfg_h = B * H > 16 ? 4 : 2;            // n8 row tiles per cell tile
    P.cfg_o = B * O > 16 ? 4 : 2;
    P.jeff_h = P.jeff_o = REC_J;
    P.nrb_h = cdiv(B * H, 8 * P.cfg_h); P.nub_h = cdiv(D, REC_J);
    P.nrb_o = cdiv(B * O, 8 * P.cfg_o); P.nub_o = cdiv(D, REC_J);
    P.cell_g_h = P.nrb_h; P.cell_g_o = P.nrb_o;
    P.cell_tiles_h_dir = P.nrb_h * P.nub_h;
    P.cell_tiles_dir = P.cell_tiles_h_dir + P.nrb_o * P.nub_o;
    P.tilesB = 2 * P.cell_tiles_dir;

    int fa = tile_smem_floats(4, 4, 3);
    const int fb = tile_smem_floats(MSG_NGL, 2, 3), fc = tile_smem_floats(4, 2, 3);
    if (fb > fa) fa = fb;
    if (fc > fa) fa = fc;
    const int fr = REC_WARPS * RES_STAGES * 32 * RES_RS;       // activation ring of the resident cell tile
    if (fr > fa) fa = fr;
    // variant: 2 = cell weights resident on chip (every CTA owns at most one cell tile and the
    // per-thread fragment words fit in tensor memory + overflow), 0 = streaming
    static int res_env_cached = -1;
    int& res_env_ref = res_env_cached;
    if (res_env_ref < 0) {
        const char* e = getenv("TGGCN_SEG_RES");
        res_env_ref = (e != nullptr && e[0] == '0') ? 0 : 1;
    }
    const int res_env = P.no_fp16_split ? 0 : res_env_cached;
    int mode = 0;
    const int kmax = (P.nk_h > 2 ? P.nk_h : 2) * D + D;
    const bool res_fits = cdiv(kmax / REC_CK, REC_WARPS) * RES_CHUNK_WORDS <= RES_TMEM_WORDS + RES_SMEM_WORDS;
    if (mode == 0 && res_env && res_fits && P.tilesB <= num_sms() && (persistent ? P.tilesA <= num_sms() : true)) mode = 2;
    // larger batches (persistent only): one resident weight slice per CTA, row blocks / video blocks walked inside the CTA
    if (mode == 0 && res_env && res_fits && persistent && 2 * (P.nub_h + P.nub_o) <= num_sms()) {
        int g[4], total;
        for (int k = 0; k < 4; ++k) g[k] = P.n_vb[k];
        auto count = [&]() { int t = 0; for (int k = 0; k < 4; ++k) t += (k == 0 && !P.hh) ? 0 : g[k] * nub_msg; return 2 * t; };
        while ((total = count()) > num_sms()) {
            int big = -1;
            for (int k = 0; k < 4; ++k)
                if (!(k == 0 && !P.hh) && g[k] > 1 && (big < 0 || g[k] > g[big])) big = k;
            if (big < 0) break;
            g[big] = (g[big] + 1) / 2;
        }
        // row-block groups of the cell tiles: as many CTAs per weight slice as fit (each keeps its own copy of the slice)
        int gh = P.nrb_h, go = P.nrb_o;
        while (2 * (P.nub_h * gh + P.nub_o * go) > num_sms() && (gh > 1 || go > 1)) {
            if (go >= gh) go = (go + 1) / 2; else gh = (gh + 1) / 2;
        }
        if (total <= num_sms()) {
            mode = 2;
            P.res_multi = 1;
            P.cell_g_h = gh; P.cell_g_o = go;
            begin = 0;
            for (int k = 0; k < 4; ++k) {
                P.msg_g[k] = g[k];
                P.msg_tile_begin[k] = begin;
                P.msg_tiles_kind[k] = (k == 0 && !P.hh) ? 0 : g[k] * nub_msg;
                begin += P.msg_tiles_kind[k];
            }
            if (!P.hh) P.msg_tile_begin[0] = 0;
            P.msg_tile_begin[4] = begin;
            P.msg_tiles_dir = begin;
            P.tilesA = 2 * begin;
            P.cell_tiles_h_dir = P.nub_h * gh;
            P.cell_tiles_dir = P.nub_h * gh + P.nub_o * go;
            P.tilesB = 2 * P.cell_tiles_dir;
        }
    }
    // resident message weights as well: one message tile per CTA for the whole launch, D*256 bytes of fragments fit beside the rest
    const size_t static_smem = 12288;                 // upper bound of the kernel's static shared memory (SegShared + barrier slots)
    P.res_msg = 0;
    if (mode == 2 && persistent && MSG_UNITS == RES_MSG_GROUPS * REC_J) {
        int fm = REC_WARPS * RES_STAGES * 32 * RES_RS;                       // activation ring (message and cell tiles)
        const int red_msg = REC_WARPS * MSG_NGL * 2 * 4 * 32, red_cell = 4 * 4 * 4 * 4 * 32;
        if (red_msg > fm) fm = red_msg;
        if (red_cell > fm) fm = red_cell;
        const size_t need = sizeof(float) * ((size_t)fm + RES_SMEM_WORDS * REC_THREADS) + (size_t)cdiv(D / REC_CK, REC_WARPS) * RES_MSG_GROUPS * 2 * REC_THREADS * 16;
        if (need + static_smem <= 227 * 1024) { P.res_msg = 1; fa = fm; }
    }
    P.res_ring_floats = fa;
    auto kern = mode == 2 ? segment_kernel<2> : segment_kernel<0>;
    const int threads = REC_THREADS;
    size_t smem = sizeof(float) * (size_t)fa + (mode == 2 ? sizeof(float) * RES_SMEM_WORDS * REC_THREADS : 0)
                                  + (P.res_msg ? (size_t)cdiv(D / REC_CK, REC_WARPS) * RES_MSG_GROUPS * 2 * REC_THREADS * 16 : 0);
    // The resident variant allocates all 512 tensor-memory columns of its SM: a second CTA of this kernel on the same SM would
    // block in tcgen05.alloc while the first one spins at the grid barrier.  Requesting more than half of the SM's shared memory
    // makes co-residency impossible at small hidden sizes too (2 x (116 KB + static + 1 KB reserved) > 228 KB).
    if (mode == 2 && smem < 116 * 1024) smem = 116 * 1024;
    if (int rc = ensure_smem((const void*)kern, smem)) return rc;
    int per_sm = 0;
    TG_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem));
    TG_REQUIRE(per_sm >= 1, "segment: kernel does not fit on an SM (smem %zu)", smem);
    TG_REQUIRE(mode != 2 || per_sm == 1, "segment: the tensor-memory resident variant must own its SM (occupancy %d)", per_sm);
    const int capacity = per_sm * num_sms();

    if (persistent) {
        int grid = P.tilesA > P.tilesB ? P.tilesA : P.tilesB;
        if (grid > capacity) grid = capacity;
        TG_CUDA_OK(cudaMemsetAsync(P.sync.counter, 0, sizeof(unsigned int), stream));      // the error word belongs to the caller (tggcn_forward zeroes it once)
        int s0 = 0, s1 = P.T, phases = 3, pers = 1;
#ifdef TGGCN_TIMING_EXPERIMENTS      // never in the product build: skipping a phase leaves the outputs undefined
        if (const char* e = getenv("TGGCN_SEG_PHASES")) phases = atoi(e);
#endif
        void* args[] = {(void*)&P, (void*)&s0, (void*)&s1, (void*)&phases, (void*)&pers};
        TG_CUDA_OK(cudaLaunchCooperativeKernel((const void*)kern, dim3(grid), dim3(threads), args, smem, stream));
        ++g_launches;
    } else {
        for (int s = 0; s < P.T; ++s) {
            kern<<<P.tilesA, threads, smem, stream>>>(P, s, s + 1, 1, 0);
            TG_LAUNCH_OK();
            kern<<<P.tilesB, threads, smem, stream>>>(P, s, s + 1, 2, 0);
            TG_LAUNCH_OK();
        }
    }
    return 0;
}

}  // namespace tg
