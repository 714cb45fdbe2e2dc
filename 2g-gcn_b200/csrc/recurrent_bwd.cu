// Backward through time of the two recurrent stages as PERSISTENT cooperative kernels (reverse-time mirrors of
// bigru.cu / segment.cu; autograd of vhoi/models.py:649-651,:983-1002 and :785-880,:1535-1564):
//
//   bigru_bwd_kernel    per reverse step: carry = z (.) dh + dGh W_hh   (gate tile on W_hh^T, K = 3D), then in the
//                       same tile's epilogue the GRU-cell backward of the NEXT reverse step (elementwise in the unit),
//                       which writes that step's dGi / dGh rows.  One grid barrier per step.
//   segment_bwd_kernel  per reverse step three phases separated by grid barriers:
//                         A  d mg   = dGi W_ih[:, segment-message columns]           (gate tiles on W_ih[:,seg]^T)
//                         B  attention/message backward per (direction, video, kind)  -> d pre-activations of the message
//                            MLPs, gradients of the attention logits w.r.t. the previous states
//                         C  carry  = direct + dGh W_hh + d pre W_msg + logit terms   (gate tiles, two K segments), and in
//                            the epilogue the gated-cell backward of the next reverse step (writes dGs / dGhs / d u).
// Weight gradients are NOT accumulated here: dGi/dGh/d pre of every step are kept and contracted afterwards in a few
// large GEMMs over all (video, t, entity) rows (api_bwd.cu).
#include <stdlib.h>
#include "recurrent.cuh"
#include "recurrent_res.cuh"
#include "backward.cuh"

namespace tg {

namespace {

constexpr int BW_MAXE = 16;

struct BwdShared {
    const float* tab1[4 * REC_J + 32];
    const float* tab2[4 * REC_J + 32];
    float al[BW_MAXE * BW_MAXE], da[BW_MAXE * BW_MAXE], dl[BW_MAXE * BW_MAXE];
    int s_fail;
};

// sum over the 16 lanes that share a row (lanes [0,16) and [16,32) of a warp hold two different rows)
__device__ __forceinline__ float half_warp_sum(float v) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---------------------------------------------------------------------------------------------------------------
// frame-level BiGRU
// ---------------------------------------------------------------------------------------------------------------
// RES: the tile's W_hh^T rows stay on chip as fp32 mma fragments in tensor memory (recurrent_res.cuh: res32_*); the CTA owns this
// tile for the whole launch and `ready` says whether they are loaded.
template <int NG, int NT, bool RES>
__device__ __forceinline__ void bigru_bwd_tile(const BiGruBwdParams& P, const BiGruBwdGroup& G, int local, int s, float* smem,
                                               BwdShared& sh, const ResState& rs, int& ready) {
    constexpr int RBT = 8 * NT, NPAIR = NT / 2, WR = NG * REC_J;
    const int D = P.D, T = P.T;
    const int per_dir = G.n_rb * G.n_ub;
    const int dir = local / per_dir;
    const int rem = local - dir * per_dir;
    const int rb = rem / G.n_ub, ub = rem - rb * G.n_ub;
    const int row0 = rb * RBT, unit0 = ub * WR;
    const int t = dir == 0 ? T - 1 - s : s;            // reverse of the forward recurrence order
    const int tr = dir == 0 ? t + 1 : t - 1;           // time handled by the previous reverse step
    const int tprev = dir == 0 ? t - 1 : t + 1;        // the step whose state fed this one in the forward
    const bool has_prev = tprev >= 0 && tprev < T;
    const int tid = threadIdx.x;

    __syncthreads();
    if (tid < WR) {
        const int unit = unit0 + tid;
        sh.tab1[tid] = unit < D ? G.whhT[dir] + (size_t)unit * 3 * D : nullptr;
    } else if (tid < WR + RBT) {
        const int r = row0 + tid - WR;
        const float* ptr = nullptr;
        if (r < G.rows && s > 0) {
            const int b = r / G.E, e = r - b * G.E;
            ptr = G.dgh + (((size_t)(b * T + tr) * G.E + e) * 2 + dir) * 3 * D;
        }
        sh.tab1[tid] = ptr;
    }
    // epilogue operands, fetched before the K loop so that their latency overlaps it
    float e_dh[NPAIR][NG], e_r[NPAIR][NG], e_z[NPAIR][NG], e_n[NPAIR][NG], e_hn[NPAIR][NG], e_hp[NPAIR][NG];
    bool ok[NPAIR];
    size_t e_fe[NPAIR];
#pragma unroll
    for (int p = 0; p < NPAIR; ++p) {
        const int lr = (tid >> 4) + 16 * p, r = row0 + lr;
        ok[p] = r < G.rows && lr < RBT;
        e_fe[p] = 0;
#pragma unroll
        for (int g = 0; g < NG; ++g) e_dh[p][g] = e_r[p][g] = e_z[p][g] = e_n[p][g] = e_hn[p][g] = e_hp[p][g] = 0.0f;
        if (!ok[p]) continue;
        const int b = r / G.E, e = r - b * G.E;
        const size_t fe = (size_t)(b * T + t) * G.E + e;
        e_fe[p] = fe;
#pragma unroll
        for (int g = 0; g < NG; ++g) {
            const int unit = unit0 + g * REC_J + (tid & 15);
            if (unit >= D) continue;
            e_dh[p][g] = G.dhfr[fe * 2 * D + dir * D + unit] + (s > 0 ? G.direct[((size_t)dir * G.rows + r) * D + unit] : 0.0f);
            const float* gt = G.gates + (fe * 2 + dir) * 4 * D + unit;
            e_r[p][g] = gt[0]; e_z[p][g] = gt[D]; e_n[p][g] = gt[2 * D]; e_hn[p][g] = gt[3 * D];
            e_hp[p][g] = has_prev ? G.hfr[((size_t)(b * T + tprev) * G.E + e) * 2 * D + dir * D + unit] : 0.0f;
        }
    }
    __syncthreads();

    float acc[NG][NPAIR];
    if (RES) {
        if (!ready) { res32_fill<NG>(rs, nullptr, 0, sh.tab1, sh.tab1, 3 * D, 0); ready = 1; }
        tile_accumulate_res32<NG, NT>(acc, sh.tab1 + WR, sh.tab1 + WR, s > 0 ? 3 * D : 0, 0, rs, nullptr, 0, G.whhT[dir], smem);
    } else {
        tile_accumulate<NG, NT, 3>(acc, sh.tab1, sh.tab1, s > 0 ? 3 * D : 0, 0, 0u, 0u, G.whhT[dir], smem);
    }

#pragma unroll
    for (int p = 0; p < NPAIR; ++p) {
        if (!ok[p]) continue;
        const int r = row0 + (tid >> 4) + 16 * p;
        const size_t fe = e_fe[p];
#pragma unroll
        for (int g = 0; g < NG; ++g) {
            const int unit = unit0 + g * REC_J + (tid & 15);
            if (unit >= D) continue;
            const float dh = e_dh[p][g] + (s > 0 ? acc[g][p] : 0.0f);
            const float rr = e_r[p][g], z = e_z[p][g], n = e_n[p][g], hn = e_hn[p][g], hprev = e_hp[p][g];
            // h = n + z (hprev - n)
            const float dan = dh * (1.0f - z) * (1.0f - n * n);
            const float daz = dh * (hprev - n) * z * (1.0f - z);
            const float dar = dan * hn * rr * (1.0f - rr);
            float* gi = G.dgi + (fe * 2 + dir) * 3 * D + unit;
            gi[0] = dar; gi[D] = daz; gi[2 * D] = dan;
            float* gh = G.dgh + (fe * 2 + dir) * 3 * D + unit;
            gh[0] = dar; gh[D] = daz; gh[2 * D] = dan * rr;
            G.direct[((size_t)dir * G.rows + r) * D + unit] = dh * z;
        }
    }
}

template <bool RES>
__global__ void __launch_bounds__(REC_THREADS, 1) bigru_bwd_kernel(const BiGruBwdParams P, int s_begin, int s_end, int persistent) {
    extern __shared__ __align__(16) float smem[];
    __shared__ BwdShared sh;
    __shared__ uint32_t tmem_slot;
    if (threadIdx.x == 0) sh.s_fail = 0;
    ResState rs;
    int ready = 0;
    if (RES) res_init(rs, &tmem_slot, nullptr);
    unsigned int epoch = 0;
    for (int s = s_begin; s < s_end; ++s) {
        for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
            int gi = 0;
#pragma unroll 1
            for (int i = 1; i < P.ngroups; ++i)
                if (tile >= P.g[i].tile_begin) gi = i;
            const BiGruBwdGroup& G = P.g[gi];
            const int local = tile - G.tile_begin;
            if (P.NG == 2) {
                if (G.cfg == 4) bigru_bwd_tile<2, 4, RES>(P, G, local, s, smem, sh, rs, ready);
                else            bigru_bwd_tile<2, 2, RES>(P, G, local, s, smem, sh, rs, ready);
            } else {
                if (G.cfg == 4) bigru_bwd_tile<1, 4, RES>(P, G, local, s, smem, sh, rs, ready);
                else            bigru_bwd_tile<1, 2, RES>(P, G, local, s, smem, sh, rs, ready);
            }
        }
        if (persistent && s + 1 < s_end)
            if (!grid_barrier(P.sync, epoch, gridDim.x, &sh.s_fail)) break;
    }
    if (RES) res_finish(rs);
}

// ---------------------------------------------------------------------------------------------------------------
// segment-level recurrent graph
// ---------------------------------------------------------------------------------------------------------------
// Phase A: d mg[dir][row][c] = sum_k dGs[t][row][k] * W_ih[k][col0 + c]
// Resident-weight state of the segment BPTT kernel: the phase-A tile's W_ih[:,seg]^T rows occupy words [0, wordsA) of the
// per-thread word space, the phase-C tile's [W_hh^T | W_msg^T] rows the words after them (tensor memory, then the overflow).
struct SegBwdRes {
    ResState rs;
    uint4* ovf;
    int wordsA;
    int readyA, readyC;
};

template <int NG, int NT, bool RES>
__device__ __forceinline__ void seg_bwd_dmg_tile(const SegBwdParams& P, bool is_h, int dir, int ub, int rb, int s, float* smem,
                                                 BwdShared& sh, SegBwdRes& R) {
    constexpr int RBT = 8 * NT, NPAIR = NT / 2, WR = NG * REC_J;
    const int D = P.D, T = P.T;
    const int E = is_h ? P.H : P.O, rows = P.B * E;
    const int nk = is_h ? P.nk_h : 2;
    const int row0 = rb * RBT, unit0 = ub * WR;
    const int t = dir == 0 ? T - 1 - s : s;
    const int tid = threadIdx.x;
    const float* wT = is_h ? P.wihT_h[dir] : P.wihT_o[dir];       // (nk*D, 3D)
    const float* dgs = is_h ? P.dgs_h : P.dgs_o;
    float* dmg = is_h ? P.dmg_h : P.dmg_o;

    __syncthreads();
    if (tid < WR) {
        const int c = unit0 + tid;
        sh.tab1[tid] = c < nk * D ? wT + (size_t)c * 3 * D : nullptr;
    } else if (tid < WR + RBT) {
        const int r = row0 + tid - WR;
        const float* ptr = nullptr;
        if (r < rows) {
            const int b = r / E, e = r - b * E;
            ptr = dgs + (((size_t)(b * T + t) * E + e) * 2 + dir) * 3 * D;
        }
        sh.tab1[tid] = ptr;
    }
    __syncthreads();
    float acc[NG][NPAIR];
    if (RES) {
        if (!R.readyA) { res32_fill<NG>(R.rs, R.ovf, 0, sh.tab1, sh.tab1, 3 * D, 0); R.readyA = 1; }
        tile_accumulate_res32<NG, NT>(acc, sh.tab1 + WR, sh.tab1 + WR, 3 * D, 0, R.rs, R.ovf, 0, wT, smem);
    } else {
        tile_accumulate<NG, NT, 3>(acc, sh.tab1, sh.tab1, 3 * D, 0, 0u, 0u, wT, smem);
    }
#pragma unroll
    for (int p = 0; p < NPAIR; ++p) {
        const int lr = (tid >> 4) + 16 * p, r = row0 + lr;
        if (r >= rows || lr >= RBT) continue;
#pragma unroll
        for (int g = 0; g < NG; ++g) {
            const int c = unit0 + g * REC_J + (tid & 15);
            if (c < nk * D) dmg[((size_t)dir * rows + r) * nk * D + c] = acc[g][p];
        }
    }
}

// Phase B: one (direction, video, message kind) item.  Reads d mg, the saved messages / attention weights and the previous
// states; writes the gradient of the message MLP pre-activations (dense for phase C, per-step for the weight gradients)
// and the attention-logit terms of d s_prev for receivers (lgr) and senders (lgs).
__device__ __forceinline__ void seg_bwd_msg_item(const SegBwdParams& P, int item, int s, float* sm, BwdShared& sh) {
    const int D = P.D, T = P.T, B = P.B, H = P.H, O = P.O, nk = P.nk_h;
    const int kind = item & 3;
    const int db = item >> 2;
    const int dir = db / B, b = db - dir * B;
    if (kind == 0 && !P.hh) return;
    const int t = dir == 0 ? T - 1 - s : s;
    const int tprev = dir == 0 ? t - 1 : t + 1;
    const bool has_prev = tprev >= 0 && tprev < T;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool send_h = (kind == 0 || kind == 2), recv_h = (kind == 0 || kind == 1);
    const int Es = send_h ? H : O, Er = recv_h ? H : O;
    const int rows_r = B * Er, rows_s = B * Es;
    const int slot = recv_h ? (kind == 0 ? 0 : nk - 1) : kind - 2;
    const int ldr = recv_h ? nk * D : 2 * D;
    const int nks = P.hh ? 2 : 1;
    const int ME = H > O ? H : O;
    float* dmg = sm;                    // [Er][D]  this kind's slice of d mg
    float* msg = dmg + ME * D;          // [Es][D]  saved post-ReLU messages
    float* sr = msg + ME * D;           // [Er][D]  receivers' previous states
    float* ss = sr + ME * D;            // [Es][D]  senders' previous states

    __syncthreads();
    const int d4 = D / 4;
    {   // all four operand blocks in flight at once (cp.async through L2: d mg was written by other CTAs in phase A)
        const float* src = (recv_h ? P.dmg_h : P.dmg_o) + ((size_t)dir * rows_r + b * Er) * ldr + slot * D;
        const float* msrc = P.smsg[kind] + (((size_t)dir * B + b) * T + t) * Es * D;
        const float* hr = (recv_h ? P.hx_h : P.hx_o) + ((size_t)(b * T + (has_prev ? tprev : 0)) * Er) * 2 * D + dir * D;
        const float* hs = (send_h ? P.hx_h : P.hx_o) + ((size_t)(b * T + (has_prev ? tprev : 0)) * Es) * 2 * D + dir * D;
        const int nr = Er * d4, ns = Es * d4;
        for (int i = tid; i < nr; i += REC_THREADS) {
            const int r = i / d4, c = (i - r * d4) * 4;
            cp_async16(dmg + r * D + c, src + (size_t)r * ldr + c);
            if (has_prev) cp_async16(sr + r * D + c, hr + (size_t)r * 2 * D + c);
        }
        for (int i = tid; i < ns; i += REC_THREADS) {
            const int q = i / d4, c = (i - q * d4) * 4;
            cp_async16(msg + q * D + c, msrc + (size_t)q * D + c);
            if (has_prev) cp_async16(ss + q * D + c, hs + (size_t)q * 2 * D + c);
        }
        cp_async_commit();
        if (tid < Er * Es) sh.al[tid] = P.salpha[kind][((((size_t)dir * B + b) * T + t) * Er) * Es + tid];
        cp_async_wait<0>();
    }
    __syncthreads();
    // d alpha[r][q] = <d mg[r], msg[q]>
    for (int p = warp; p < Er * Es; p += REC_WARPS) {
        const int r = p / Es, q = p - r * Es;
        const float4* x = reinterpret_cast<const float4*>(dmg + r * D);
        const float4* y = reinterpret_cast<const float4*>(msg + q * D);
        float acc = 0.0f;
        for (int c = lane; c < d4; c += 32) {
            const float4 u = x[c], v = y[c];
            acc = fmaf(u.x, v.x, acc); acc = fmaf(u.y, v.y, acc); acc = fmaf(u.z, v.z, acc); acc = fmaf(u.w, v.w, acc);
        }
        acc = warp_sum(acc);
        if (lane == 0) sh.da[p] = acc;
    }
    __syncthreads();
    if (tid < Er) {
        const float scale = P.att_noscale ? 1.0f : 1.0f / sqrtf((float)D);
        float dot = 0.0f;
        for (int q = 0; q < Es; ++q) dot = fmaf(sh.al[tid * Es + q], sh.da[tid * Es + q], dot);
        for (int q = 0; q < Es; ++q)                   // mean pooling: the weights do not depend on the states
            sh.dl[tid * Es + q] = (P.mean_pool || P.dist_kind[kind]) ? 0.0f : sh.al[tid * Es + q] * (sh.da[tid * Es + q] - dot) * scale;
    }
    __syncthreads();
    {   // gradient of the pre-activation of every sender's message MLP; attention-logit terms of the senders
        const int col = send_h ? (kind == 0 ? 0 : (nks - 1) * D) : (kind == 1 ? 0 : D);
        const int lds = send_h ? nks * D : 2 * D;
        float* dense = (send_h ? P.dpre_h : P.dpre_o) + ((size_t)dir * rows_s + b * Es) * lds + col;
        float* all = P.dpre_all[kind] + (((size_t)dir * B + b) * T + t) * Es * D;
        float* lgs = P.lgs[kind] + ((size_t)dir * rows_s + b * Es) * D;
        for (int i = tid; i < Es * d4; i += REC_THREADS) {
            const int q = i / d4, c = (i - q * d4) * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f), g = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int r = 0; r < Er; ++r) {
                const float a = sh.al[r * Es + q];
                const float4 dm = *reinterpret_cast<const float4*>(dmg + r * D + c);
                v.x = fmaf(a, dm.x, v.x); v.y = fmaf(a, dm.y, v.y); v.z = fmaf(a, dm.z, v.z); v.w = fmaf(a, dm.w, v.w);
                if (has_prev) {
                    const float l = sh.dl[r * Es + q];
                    const float4 st = *reinterpret_cast<const float4*>(sr + r * D + c);
                    g.x = fmaf(l, st.x, g.x); g.y = fmaf(l, st.y, g.y); g.z = fmaf(l, st.z, g.z); g.w = fmaf(l, st.w, g.w);
                }
            }
            const float4 m = *reinterpret_cast<const float4*>(msg + q * D + c);
            v.x = m.x > 0.0f ? v.x : 0.0f; v.y = m.y > 0.0f ? v.y : 0.0f; v.z = m.z > 0.0f ? v.z : 0.0f; v.w = m.w > 0.0f ? v.w : 0.0f;
            *reinterpret_cast<float4*>(dense + (size_t)q * lds + c) = v;
            *reinterpret_cast<float4*>(all + (size_t)q * D + c) = v;
            if (has_prev) *reinterpret_cast<float4*>(lgs + (size_t)q * D + c) = g;
        }
    }
    if (has_prev) {   // attention-logit terms of the receivers:  d s_r += sum_q dl[r][q] s_q
        float* lgr = P.lgr[kind] + ((size_t)dir * rows_r + b * Er) * D;
        for (int i = tid; i < Er * d4; i += REC_THREADS) {
            const int r = i / d4, c = (i - r * d4) * 4;
            float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int q = 0; q < Es; ++q) {
                const float l = sh.dl[r * Es + q];
                const float4 st = *reinterpret_cast<const float4*>(ss + q * D + c);
                g.x = fmaf(l, st.x, g.x); g.y = fmaf(l, st.y, g.y); g.z = fmaf(l, st.z, g.z); g.w = fmaf(l, st.w, g.w);
            }
            *reinterpret_cast<float4*>(lgr + (size_t)r * D + c) = g;
        }
    }
}

// Phase C (s >= 0): carry into the previous state of reverse step s, then the gated-cell backward of reverse step s + 1.
// s == -1: only the cell backward of reverse step 0 (no carry yet).
template <int NG, int NT, bool RES>
__device__ __forceinline__ void seg_bwd_carry_tile(const SegBwdParams& P, bool is_h, int dir, int ub, int rb, int s, float* smem,
                                                   BwdShared& sh, SegBwdRes& R) {
    constexpr int RBT = 8 * NT, NPAIR = NT / 2, WR = NG * REC_J;
    const int D = P.D, T = P.T;
    const int E = is_h ? P.H : P.O, rows = P.B * E;
    const int nks = is_h ? (P.hh ? 2 : 1) : 2;
    const int row0 = rb * RBT, unit0 = ub * WR;
    const int tid = threadIdx.x;
    const float* whhT = is_h ? P.whhT_h[dir] : P.whhT_o[dir];     // (D, 3D)
    const float* wmT = is_h ? P.wmT_h : P.wmT_o;                  // (D, nks*D)
    const float* dghs = is_h ? P.dghs_h : P.dghs_o;
    const float* dpre = is_h ? P.dpre_h : P.dpre_o;
    const bool gemm = s >= 0;
    const int tcur = dir == 0 ? T - 1 - s : s;                    // time of reverse step s (valid when gemm)
    const int sn = s + 1;                                         // the reverse step whose cell backward runs in the epilogue
    const int t = dir == 0 ? T - 1 - sn : sn;
    const int tprev = dir == 0 ? t - 1 : t + 1;
    const bool has_prev = tprev >= 0 && tprev < T;

    __syncthreads();
    if (tid < WR) {
        const int unit = unit0 + tid;
        sh.tab1[tid] = (gemm && unit < D) ? whhT + (size_t)unit * 3 * D : nullptr;
        sh.tab2[tid] = (gemm && unit < D) ? wmT + (size_t)unit * nks * D : nullptr;
    } else if (tid < WR + RBT) {
        const int r = row0 + tid - WR;
        const float* p1 = nullptr;
        const float* p2 = nullptr;
        if (gemm && r < rows) {
            const int b = r / E, e = r - b * E;
            p1 = dghs + (((size_t)(b * T + tcur) * E + e) * 2 + dir) * 3 * D;
            p2 = dpre + ((size_t)dir * rows + r) * nks * D;
        }
        sh.tab1[tid] = p1;
        sh.tab2[tid] = p2;
    }
    // epilogue operands: fetched before the K loop so that their latency overlaps it (the attention-logit terms were written
    // by phase B, complete since the last grid barrier)
    const float* hx = is_h ? P.hx_h : P.hx_o;
    const float* dhx = is_h ? P.dhx_h : P.dhx_o;
    const float* sgb = is_h ? P.sgates_h : P.sgates_o;
    const float* ub_ = is_h ? P.u_h : P.u_o;
    float* dgs_o = is_h ? P.dgs_h : P.dgs_o;
    float* dghs_o = is_h ? P.dghs_h : P.dghs_o;
    float* du = is_h ? P.du_h : P.du_o;
    float* direct = is_h ? P.direct_h : P.direct_o;
    float e_base[NPAIR][NG], e_dhx[NPAIR][NG], e_r[NPAIR][NG], e_z[NPAIR][NG], e_n[NPAIR][NG], e_hn[NPAIR][NG], e_hp[NPAIR][NG], e_u[NPAIR];
    bool rvalid[NPAIR];
    size_t e_fe[NPAIR];
#pragma unroll
    for (int p = 0; p < NPAIR; ++p) {
        const int lr = (tid >> 4) + 16 * p, r = row0 + lr;
        rvalid[p] = r < rows && lr < RBT;                         // uniform over the 16 lanes that share the row
        e_u[p] = 0.0f;
        e_fe[p] = 0;
#pragma unroll
        for (int g = 0; g < NG; ++g) e_base[p][g] = e_dhx[p][g] = e_r[p][g] = e_z[p][g] = e_n[p][g] = e_hn[p][g] = e_hp[p][g] = 0.0f;
        if (rvalid[p]) {
            const int b = r / E, e = r - b * E;
            const size_t fe = (size_t)(b * T + t) * E + e;
            e_fe[p] = fe;
            e_u[p] = ub_[fe];
#pragma unroll
            for (int g = 0; g < NG; ++g) {
                const int unit = unit0 + g * REC_J + (tid & 15);
                if (unit >= D) continue;
                const size_t ro = ((size_t)dir * rows + r) * D + unit;
                float base = 0.0f;
                if (gemm) {
                    base = direct[ro];
                    // attention-logit terms (kinds: 0 hh, 1 oh, 2 ho, 3 oo)
                    if (is_h) {
                        if (P.hh) base += ld_cg(P.lgr[0] + ro) + ld_cg(P.lgs[0] + ro);
                        base += ld_cg(P.lgr[1] + ro) + ld_cg(P.lgs[2] + ro);
                    } else {
                        base += ld_cg(P.lgs[1] + ro) + ld_cg(P.lgr[2] + ro) + ld_cg(P.lgr[3] + ro) + ld_cg(P.lgs[3] + ro);
                    }
                }
                e_base[p][g] = base;
                e_dhx[p][g] = dhx[fe * 2 * D + dir * D + unit];
                const float* sg = sgb + (fe * 2 + dir) * 4 * D + unit;
                e_r[p][g] = sg[0]; e_z[p][g] = sg[D]; e_n[p][g] = sg[2 * D]; e_hn[p][g] = sg[3 * D];
                e_hp[p][g] = has_prev ? hx[((size_t)(b * T + tprev) * E + e) * 2 * D + dir * D + unit] : 0.0f;
            }
        }
    }
    __syncthreads();
    float acc[NG][NPAIR];
    if (RES) {
        if (gemm && !R.readyC) { res32_fill<NG>(R.rs, R.ovf, R.wordsA, sh.tab1, sh.tab2, 3 * D, nks * D); R.readyC = 1; }
        tile_accumulate_res32<NG, NT>(acc, sh.tab1 + WR, sh.tab2 + WR, gemm ? 3 * D : 0, gemm ? nks * D : 0, R.rs, R.ovf, R.wordsA, whhT, smem);
    } else {
        tile_accumulate<NG, NT, 3>(acc, sh.tab1, sh.tab2, gemm ? 3 * D : 0, gemm ? nks * D : 0, 0u, 0u, whhT, smem);
    }

#pragma unroll
    for (int p = 0; p < NPAIR; ++p) {
        const int lr = (tid >> 4) + 16 * p, r = row0 + lr;
        float du_part = 0.0f;
        if (rvalid[p]) {
            const size_t fe = e_fe[p];
            const float u = e_u[p];
#pragma unroll
            for (int g = 0; g < NG; ++g) {
                const int unit = unit0 + g * REC_J + (tid & 15);
                if (unit >= D) continue;
                const size_t ro = ((size_t)dir * rows + r) * D + unit;
                const float dH = e_dhx[p][g] + e_base[p][g] + (gemm ? acc[g][p] : 0.0f);
                const float rr = e_r[p][g], z = e_z[p][g], n = e_n[p][g], hn = e_hn[p][g], hprev = e_hp[p][g];
                const float hnew = n + z * (hprev - n);
                du_part += dH * (hnew - hprev);
                const float dhnew = u * dH;
                const float dan = dhnew * (1.0f - z) * (1.0f - n * n);
                const float daz = dhnew * (hprev - n) * z * (1.0f - z);
                const float dar = dan * hn * rr * (1.0f - rr);
                float* gs = dgs_o + (fe * 2 + dir) * 3 * D + unit;
                gs[0] = dar; gs[D] = daz; gs[2 * D] = dan;
                float* gh = dghs_o + (fe * 2 + dir) * 3 * D + unit;
                gh[0] = dar; gh[D] = daz; gh[2 * D] = dan * rr;
                direct[ro] = (1.0f - u) * dH + dhnew * z;
            }
        }
        du_part = half_warp_sum(du_part);
        if (rvalid[p] && (tid & 15) == 0) atomicAdd(du + e_fe[p], du_part);
    }
}

// tile index -> (cell type, dir, unit block); one row block per cell type (B*E <= 32 rows per tile, see launcher)
template <int PHASE, bool RES>
__device__ __forceinline__ void seg_bwd_dispatch(const SegBwdParams& P, int tile, int s, float* smem, BwdShared& sh, SegBwdRes& R) {
    const int per_dir = PHASE == 0 ? P.tilesA_dir : P.tilesC_dir;
    const int nh = PHASE == 0 ? P.tilesA_h : P.tilesC_h;           // (row blocks x unit blocks) of the human cell, per direction
    const int dir = tile / per_dir;
    int rem = tile - dir * per_dir;
    const bool is_h = rem < nh;
    if (!is_h) rem -= nh;
    const int nub = PHASE == 0 ? (is_h ? P.nubA_h : P.nubA_o) : P.nubC;
    const int rb = rem / nub, ub = rem - rb * nub;
    const int cfg = is_h ? P.cfg_h : P.cfg_o;
    if (PHASE == 0) {
        if (cfg == 4) seg_bwd_dmg_tile<2, 4, RES>(P, is_h, dir, ub, rb, s, smem, sh, R);
        else          seg_bwd_dmg_tile<2, 2, RES>(P, is_h, dir, ub, rb, s, smem, sh, R);
    } else {
        if (cfg == 4) seg_bwd_carry_tile<1, 4, RES>(P, is_h, dir, ub, rb, s, smem, sh, R);
        else          seg_bwd_carry_tile<1, 2, RES>(P, is_h, dir, ub, rb, s, smem, sh, R);
    }
}

// phases: bit 0 = A, bit 1 = B, bit 2 = C.  Steps s in [s_begin, s_end); s = -1 runs only the initial cell backward (phase C).
// RES (persistent launches whose A and C tiles each fit one per CTA): transposed weights resident on chip; res_ovf_floats = offset of
// the shared-memory overflow fragments behind the ring / phase-B region.
template <bool RES>
__global__ void __launch_bounds__(REC_THREADS, 1) segment_bwd_kernel(const SegBwdParams P, int s_begin, int s_end, int phases,
                                                                    int persistent, int res_ovf_floats, int res_words_a) {
    extern __shared__ __align__(16) float smem[];
    __shared__ BwdShared sh;
    __shared__ uint32_t tmem_slot;
    if (threadIdx.x == 0) sh.s_fail = 0;
    SegBwdRes R;
    R.ovf = reinterpret_cast<uint4*>(smem + res_ovf_floats);
    R.wordsA = res_words_a;
    R.readyA = R.readyC = 0;
    if (RES) res_init(R.rs, &tmem_slot, nullptr);
    unsigned int epoch = 0;
    const int T = P.T;
    for (int s = s_begin; s < s_end; ++s) {
        if (s >= 0 && (phases & 1)) {
            for (int tile = blockIdx.x; tile < 2 * P.tilesA_dir; tile += gridDim.x) seg_bwd_dispatch<0, RES>(P, tile, s, smem, sh, R);
            if (persistent && !grid_barrier(P.sync, epoch, gridDim.x, &sh.s_fail)) break;
        }
        if (s >= 0 && (phases & 2)) {
            for (int item = blockIdx.x; item < 2 * P.B * 4; item += gridDim.x) seg_bwd_msg_item(P, item, s, smem, sh);
            if (persistent && s + 1 < T && !grid_barrier(P.sync, epoch, gridDim.x, &sh.s_fail)) break;
        }
        if (s + 1 < T && (phases & 4)) {
            for (int tile = blockIdx.x; tile < 2 * P.tilesC_dir; tile += gridDim.x) seg_bwd_dispatch<1, RES>(P, tile, s, smem, sh, R);
            if (persistent && !grid_barrier(P.sync, epoch, gridDim.x, &sh.s_fail)) break;
        }
    }
    if (RES) res_finish(R.rs);
}

}  // namespace

// TGGCN_BWD_RES=0 keeps the streaming 3xTF32 tiles (A/B aid)
static bool bwd_res_enabled() {
    static int enabled = -1;
    if (enabled < 0) {
        const char* e = getenv("TGGCN_BWD_RES");
        enabled = (e != nullptr && e[0] == '0') ? 0 : 1;
    }
    return enabled != 0;
}

int launch_bigru_bwd(BiGruBwdParams& P, int persistent, cudaStream_t stream) {
    TG_REQUIRE(P.D % 16 == 0, "bigru_bwd: hidden_size=%d must be a multiple of 16", P.D);
    // tiling first (it decides whether every CTA can own one tile, i.e. whether the weights can stay resident)
    auto plan = [&](int ng) {
        int begin = 0;
        for (int i = 0; i < P.ngroups; ++i) {
            BiGruBwdGroup& G = P.g[i];
            G.cfg = G.rows > 16 ? 4 : 2;
            G.n_rb = cdiv(G.rows, 8 * G.cfg);
            G.n_ub = cdiv(P.D, REC_J * ng);
            G.tile_begin = begin;
            begin += 2 * G.n_rb * G.n_ub;
        }
        P.total_tiles = begin;
        P.NG = ng;
    };
    // resident variant: one tile per CTA for the whole launch, W_hh^T fragments (cdiv(3D/16, 8) * NG * 8 words per thread) in tensor memory
    bool res = false;
    size_t smem_res = 0;
    if (persistent && bwd_res_enabled()) {
        for (int ng = 1; ng <= 2 && !res; ++ng) {
            plan(ng);
            const int words = cdiv(3 * P.D / REC_CK, REC_WARPS) * ng * 8;
            if (P.total_tiles <= num_sms() && words <= RES_TMEM_WORDS) res = true;
        }
        if (res) {
            smem_res = sizeof(float) * (size_t)res32_smem_floats(2, 4);
            if (smem_res < 116 * 1024) smem_res = 116 * 1024;      // one CTA per SM: each allocates all of tensor memory
        }
    }
    if (res) {
        auto kern = bigru_bwd_kernel<true>;
        if (int rc = ensure_smem((const void*)kern, smem_res)) return rc;
        int per_sm = 0;
        TG_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, REC_THREADS, smem_res));
        if (per_sm == 1 && P.total_tiles <= num_sms()) {
            TG_CUDA_OK(cudaMemsetAsync(P.sync.counter, 0, sizeof(unsigned int), stream));      // the error word belongs to the caller
            int s0 = 0, s1 = P.T, pers = 1;
            void* args[] = {(void*)&P, (void*)&s0, (void*)&s1, (void*)&pers};
            TG_CUDA_OK(cudaLaunchCooperativeKernel((const void*)kern, dim3(P.total_tiles), dim3(REC_THREADS), args, smem_res, stream));
            ++g_launches;
            return 0;
        }
    }
    auto kern = bigru_bwd_kernel<false>;
    const size_t smem = sizeof(float) * (size_t)tile_smem_floats(2, 4, 3);
    if (int rc = ensure_smem((const void*)kern, smem)) return rc;
    int per_sm = 0;
    TG_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, REC_THREADS, smem));
    TG_REQUIRE(per_sm >= 1, "bigru_bwd: kernel does not fit on an SM (smem %zu)", smem);
    const int capacity = per_sm * num_sms();
    for (int ng = 1; ng <= 2; ++ng) {
        plan(ng);
        if (P.total_tiles <= capacity) break;          // prefer the finer split while every tile gets its own SM
    }
    if (persistent) {
        const int grid = P.total_tiles < capacity ? P.total_tiles : capacity;
        TG_CUDA_OK(cudaMemsetAsync(P.sync.counter, 0, sizeof(unsigned int), stream));      // the error word belongs to the caller
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

int launch_segment_bwd(SegBwdParams& P, int persistent, cudaStream_t stream) {
    const int D = P.D, B = P.B, H = P.H, O = P.O;
    TG_REQUIRE(D % 16 == 0, "segment_bwd: hidden_size=%d must be a multiple of 16", D);
    TG_REQUIRE(H <= BW_MAXE && O <= BW_MAXE, "segment_bwd: at most %d entities per type", BW_MAXE);
    P.cfg_h = B * H > 16 ? 4 : 2;
    P.cfg_o = B * O > 16 ? 4 : 2;
    const int nrb_h = cdiv(B * H, 8 * P.cfg_h), nrb_o = cdiv(B * O, 8 * P.cfg_o);
    P.nubA_h = cdiv(P.nk_h * D, 2 * REC_J); P.nubA_o = cdiv(2 * D, 2 * REC_J);
    P.tilesA_h = nrb_h * P.nubA_h; P.tilesA_dir = P.tilesA_h + nrb_o * P.nubA_o;
    P.nubC = cdiv(D, REC_J);
    P.tilesC_h = nrb_h * P.nubC; P.tilesC_dir = P.tilesC_h + nrb_o * P.nubC;
    const int items = 2 * B * 4;
    const size_t smem_b = sizeof(float) * (size_t)4 * (H > O ? H : O) * D;
    // resident variant: every CTA owns at most one phase-A tile (NG = 2, K = 3D) and one phase-C tile (NG = 1, K = 3D + nks*D) for the
    // whole launch; their fp32 fragment words go to tensor memory (256 per thread) and a shared-memory overflow behind the ring
    if (persistent && bwd_res_enabled() && 2 * P.tilesA_dir <= num_sms() && 2 * P.tilesC_dir <= num_sms()) {
        const int wordsA = cdiv(3 * D / REC_CK, REC_WARPS) * 2 * 8;
        const int wordsC = cdiv((3 * D + 2 * D) / REC_CK, REC_WARPS) * 1 * 8;
        const int ovf_words = wordsA + wordsC > RES_TMEM_WORDS ? (wordsA + wordsC - RES_TMEM_WORDS + 3) / 4 * 4 : 0;
        size_t region = sizeof(float) * (size_t)res32_smem_floats(2, 4);
        if (smem_b > region) region = smem_b;
        region = (region + 15) / 16 * 16;
        size_t smem = region + (size_t)ovf_words * REC_THREADS * 4;
        if (smem < 116 * 1024) smem = 116 * 1024;                  // one CTA per SM: each allocates all of tensor memory
        auto kern = segment_bwd_kernel<true>;
        int per_sm = 0;
        if (smem <= 200 * 1024 && ensure_smem((const void*)kern, smem) == 0 &&
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, REC_THREADS, smem) == cudaSuccess && per_sm == 1) {
            int grid = 2 * P.tilesA_dir;
            if (2 * P.tilesC_dir > grid) grid = 2 * P.tilesC_dir;
            if (items > grid) grid = items;
            if (grid > num_sms()) grid = num_sms();                 // the phase-B items are strided over the CTAs
            TG_CUDA_OK(cudaMemsetAsync(P.sync.counter, 0, sizeof(unsigned int), stream));      // the error word belongs to the caller
            int s0 = -1, s1 = P.T, phases = 7, pers = 1, ovf_off = (int)(region / sizeof(float)), wa = wordsA;
            void* args[] = {(void*)&P, (void*)&s0, (void*)&s1, (void*)&phases, (void*)&pers, (void*)&ovf_off, (void*)&wa};
            TG_CUDA_OK(cudaLaunchCooperativeKernel((const void*)kern, dim3(grid), dim3(REC_THREADS), args, smem, stream));
            ++g_launches;
            return 0;
        }
    }
    auto kern = segment_bwd_kernel<false>;
    size_t smem = sizeof(float) * (size_t)tile_smem_floats(2, 4, 3);
    if (smem_b > smem) smem = smem_b;
    TG_REQUIRE(smem <= 200 * 1024, "segment_bwd: hidden_size=%d needs %zu bytes of shared memory", D, smem);
    if (int rc = ensure_smem((const void*)kern, smem)) return rc;
    int per_sm = 0;
    TG_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, REC_THREADS, smem));
    TG_REQUIRE(per_sm >= 1, "segment_bwd: kernel does not fit on an SM (smem %zu)", smem);
    const int capacity = per_sm * num_sms();
    int zero = 0;
    if (persistent) {
        int grid = 2 * P.tilesA_dir;
        if (2 * P.tilesC_dir > grid) grid = 2 * P.tilesC_dir;
        if (items > grid) grid = items;
        if (grid > capacity) grid = capacity;
        TG_CUDA_OK(cudaMemsetAsync(P.sync.counter, 0, sizeof(unsigned int), stream));      // the error word belongs to the caller
        int s0 = -1, s1 = P.T, phases = 7, pers = 1;
#ifdef TGGCN_TIMING_EXPERIMENTS      // never in the product build: skipping a phase leaves the gradients undefined
        if (const char* e = getenv("TGGCN_SEGBWD_PHASES")) phases = atoi(e);
#endif
        void* args[] = {(void*)&P, (void*)&s0, (void*)&s1, (void*)&phases, (void*)&pers, (void*)&zero, (void*)&zero};
        TG_CUDA_OK(cudaLaunchCooperativeKernel((const void*)kern, dim3(grid), dim3(REC_THREADS), args, smem, stream));
        ++g_launches;
    } else {
        for (int s = -1; s < P.T; ++s) {
            if (s >= 0) {
                kern<<<2 * P.tilesA_dir, REC_THREADS, smem, stream>>>(P, s, s + 1, 1, 0, 0, 0);
                TG_LAUNCH_OK();
                kern<<<items, REC_THREADS, smem, stream>>>(P, s, s + 1, 2, 0, 0, 0);
                TG_LAUNCH_OK();
            }
            if (s + 1 < P.T) {
                kern<<<2 * P.tilesC_dir, REC_THREADS, smem, stream>>>(P, s, s + 1, 4, 0, 0, 0);
                TG_LAUNCH_OK();
            }
        }
    }
    return 0;
}

}  // namespace tg
