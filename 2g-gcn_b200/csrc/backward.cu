// Hand-written backward kernels of the graph stages (autograd of vhoi/models.py:664-933; Appendix B of SURVEY.md):
//   heads_bwd_kernel     Linear(2D->C)+LogSoftmax backward, scatter through the reorder index
//   frame_bwd_kernel     frame-level attention / aggregation / Gumbel-sigmoid gates (straight-through, filter)
// The recurrent stages run backward through time in recurrent_bwd.cu; dense matrix products of the backward run on the
// projection kernels (gemm_*.cu, gemm_bwd.cu).
#include "backward.cuh"

namespace tg {

// =============================================================================================================
// heads
// =============================================================================================================
constexpr int HB_ROWS = 64;        // rows (video, frame, entity) per CTA

// Two passes per CTA over its HB_ROWS rows: (1) one warp per row recomputes the logits, forms d log-softmax (kept in shared
// memory) and scatters d x = dz W into d hfr / d hx (atomics: several frames share a segment-end row); (2) every thread owns
// 2D/256 input columns and accumulates dW[c][k] = sum_rows dz[row][c] x[row][k] in registers — no shared-memory atomics — and
// flushes once per CTA.
__global__ void __launch_bounds__(256) heads_bwd_kernel(const HeadsBwdParams P) {
    __shared__ float sdz[HB_ROWS][33];
    __shared__ size_t sx[HB_ROWS];
    const int hd = blockIdx.y, src = hd >> 1;
    if (P.dlogp[hd] == nullptr) return;
    const int D2 = 2 * P.D, C = P.C;
    const bool cat = P.cat && src == 1;        // segment heads on [hx | hfr] (models.py:901-903): weight rows of 4D
    const int ldw = cat ? 2 * D2 : D2;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int rows = P.B * P.T * P.E;
    const int row0 = blockIdx.x * HB_ROWS, nrows = min(HB_ROWS, rows - row0);
    const float* W = P.w[hd];
    const float* xin = src == 0 ? P.hfr : P.hx;
    // ---- pass 1 ----
    for (int lr = warp; lr < nrows; lr += 8) {
        const int row = row0 + lr;
        const int e = row % P.E, bt = row / P.E;
        const int t = bt % P.T, b = bt / P.T;
        size_t xrow;
        if (src == 0) xrow = (size_t)row;
        else xrow = (size_t)(b * P.T + P.reidx[(size_t)bt * P.NE + P.e_off + e]) * P.E + e;
        const float* x = xin + xrow * D2;
        const float* x2 = P.hfr + (size_t)row * D2;
        float mine = -INFINITY;
        for (int c = 0; c < C; ++c) {
            const float* wr = W + (size_t)c * ldw;
            float acc = 0.0f;
            for (int k = lane * 4; k < D2; k += 128) {
                const float4 u = __ldg(reinterpret_cast<const float4*>(wr + k));
                const float4 v = *reinterpret_cast<const float4*>(x + k);
                acc = fmaf(u.x, v.x, acc); acc = fmaf(u.y, v.y, acc); acc = fmaf(u.z, v.z, acc); acc = fmaf(u.w, v.w, acc);
            }
            if (cat)
                for (int k = lane * 4; k < D2; k += 128) {
                    const float4 u = __ldg(reinterpret_cast<const float4*>(wr + D2 + k));
                    const float4 v = *reinterpret_cast<const float4*>(x2 + k);
                    acc = fmaf(u.x, v.x, acc); acc = fmaf(u.y, v.y, acc); acc = fmaf(u.z, v.z, acc); acc = fmaf(u.w, v.w, acc);
                }
            acc = warp_sum(acc) + __ldg(P.bias[hd] + c);
            if (lane == c) mine = acc;
        }
        const float m = warp_max(mine);
        const float ex = lane < C ? expf(mine - m) : 0.0f;
        const float prob = ex / warp_sum(ex);
        const float g = lane < C ? __ldg(P.dlogp[hd] + ((size_t)(b * C + lane) * P.T + t) * P.E + e) : 0.0f;
        const float gs = warp_sum(g);
        const float dz = lane < C ? g - prob * gs : 0.0f;     // d log_softmax
        sdz[lr][lane] = dz;
        if (lane == 0) sx[lr] = xrow;
        for (int part = 0; part < (cat ? 2 : 1); ++part) {   // part 1: the frame-level half of a concatenated input
            float* dxrow = part == 0 ? (src == 0 ? P.dhfr : P.dhx) + xrow * D2 : P.dhfr + (size_t)row * D2;
            for (int k0 = 0; k0 < D2; k0 += 128) {           // warp-uniform trip count: the shuffles need every lane
                const int k = k0 + lane * 4;
                const bool ok = k < D2;
                float4 dx = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int c = 0; c < C; ++c) {
                    const float dzc = __shfl_sync(0xffffffffu, dz, c);
                    if (!ok) continue;
                    const float4 wv = __ldg(reinterpret_cast<const float4*>(W + (size_t)c * ldw + part * D2 + k));
                    dx.x = fmaf(dzc, wv.x, dx.x); dx.y = fmaf(dzc, wv.y, dx.y); dx.z = fmaf(dzc, wv.z, dx.z); dx.w = fmaf(dzc, wv.w, dx.w);
                }
                if (ok) { atomicAdd(dxrow + k, dx.x); atomicAdd(dxrow + k + 1, dx.y); atomicAdd(dxrow + k + 2, dx.z); atomicAdd(dxrow + k + 3, dx.w); }
            }
        }
    }
    __syncthreads();
    // ---- pass 2: weight and bias gradients ----
    for (int k0 = 0; k0 < ldw; k0 += 256) {
        const int k = k0 + threadIdx.x;
        if (k >= ldw) continue;
        float acc[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) acc[c] = 0.0f;
        for (int lr = 0; lr < nrows; ++lr) {
            const float xv = k < D2 ? xin[sx[lr] * D2 + k] : P.hfr[(size_t)(row0 + lr) * D2 + (k - D2)];
#pragma unroll
            for (int c = 0; c < 32; ++c)
                if (c < C) acc[c] = fmaf(sdz[lr][c], xv, acc[c]);
        }
#pragma unroll
        for (int c = 0; c < 32; ++c)
            if (c < C) atomicAdd(P.dw[hd] + (size_t)c * ldw + k, acc[c]);
    }
    if (threadIdx.x < C) {
        float sacc = 0.0f;
        for (int lr = 0; lr < nrows; ++lr) sacc += sdz[lr][threadIdx.x];
        atomicAdd(P.db[hd] + threadIdx.x, sacc);
    }
}

int launch_heads_bwd(const HeadsBwdParams& P, cudaStream_t stream) {
    TG_REQUIRE(P.C >= 1 && P.C <= 32 && P.D % 2 == 0, "heads_bwd: unsupported sizes");
    const int rows = P.B * P.T * P.E;
    heads_bwd_kernel<<<dim3(cdiv(rows, HB_ROWS), 4), 256, 0, stream>>>(P);
    TG_LAUNCH_OK();
    return 0;
}

// =============================================================================================================
// frame-level graph; one CTA per (video, frame)
// =============================================================================================================
constexpr int FB_MAXE = 16;

// Gradient of the gate logit of entity `lane` of frame n = (b, t); all 32 lanes of a warp call (lanes >= H + O return 0).
// Straight-through / filter rule of the hard gates, Gumbel-sigmoid and sigmoid backward, and the
// object_segment_update_strategy rules (one human, models.py:1523-1532): 'sah' — the object gates ARE the human's, so everything
// that reaches them is handed to the human's soft gate; 'coh' — hard_o = st(y_o) * st(y_h): each factor receives the other's
// decision as its weight.
__device__ __forceinline__ float gate_dlogit(const FrameBwdParams& P, int n, int b, int t, int lane) {
    const int H = P.H, O = P.O, T = P.T, NE = H + O;
    const int strat = P.update_strategy;
    const bool act = lane < NE, is_h = lane < H;
    const int r = is_h ? lane : lane - H, E = is_h ? H : O;
    const float* given = is_h ? P.human_seg : P.object_seg;
    const bool sampled = act && given == nullptr;
    float dy = 0.0f, hand_over = 0.0f, y = 0.0f;
    if (sampled) {
        const float* soft = is_h ? P.y_hss : P.y_oss;
        const size_t oi = (size_t)(b * T + t) * E + r;
        y = soft[oi];
        float dhard = (is_h ? P.du_h : P.du_o)[oi];
        const float* dyh = is_h ? P.dy_hs : P.dy_os;
        if (dyh != nullptr) dhard += dyh[oi];
        float pass;
        if (P.filter) {
            const float yp = t > 0 ? soft[oi - E] : 0.0f, yn = t + 1 < T ? soft[oi + E] : 0.0f;
            const bool keep = (y > yp) && (y > yn) && (y >= P.thr);
            pass = (keep || y < P.thr) ? 1.0f : 0.0f;       // models.py:1660-1662 (clamp(max=0) passes at u == 0)
        } else {
            pass = t == T - 1 ? 0.0f : 1.0f;                // last step overwritten by 1 (models.py:701-702)
        }
        dy = dhard * pass;
        if (strat == 2 && !is_h) {
            const float yh = P.y_hss[(size_t)(b * T + t) * H];
            hand_over = dy * (y > P.thr ? 1.0f : 0.0f);
            dy *= yh > P.thr ? 1.0f : 0.0f;
        }
        const float* dys = is_h ? P.dy_hss : P.dy_oss;
        if (dys != nullptr) dy += dys[oi];
        if (strat == 1 && !is_h) { hand_over = dy; dy = 0.0f; }
    }
    if (strat != 0) {
        const float total = warp_sum(hand_over);
        if (lane == 0) dy += total;
    }
    float dl_ = 0.0f;
    if (sampled && !(strat == 1 && !is_h)) {
        const float p = P.pgate[(size_t)n * NE + lane];
        // y = sigmoid(log(p+eps) - log(1-p+eps) + g0 - g1), p = sigmoid(logit)
        dl_ = P.straight_through ? dy * p * (1.0f - p)          // y = p
                                 : dy * y * (1.0f - y) * (1.0f / (p + 1e-20f) + 1.0f / ((1.0f - p) + 1e-20f)) * p * (1.0f - p);
    }
    return dl_;
}

// Two-layer gate MLPs (discrete_networks_num_layers == 2), first half of their backward: one CTA per frame.  d hidden =
// dlogit * w2 (.) [hidden > 0] for every sampled entity, dw2 += dlogit * hidden, db2 += dlogit (atomics).
__global__ void __launch_bounds__(128) gate_bwd_kernel(const FrameBwdParams P) {
    __shared__ float dlogit[32];
    const int H = P.H, O = P.O, NE = H + O, D = P.D, T = P.T;
    const int n = blockIdx.x, b = n / T, t = n - b * T, tid = threadIdx.x;
    if (tid < 32) {
        const float dl_ = gate_dlogit(P, n, b, t, tid);
        if (tid < NE) dlogit[tid] = dl_;
    }
    __syncthreads();
    for (int c = tid; c < D; c += 128) {
        float aw_h = 0.0f, aw_o = 0.0f;
        if (P.dhid_h != nullptr) {
            const float w2 = __ldg(P.w2_h + c);
            for (int h = 0; h < H; ++h) {
                const float hv = P.hid_h[((size_t)n * H + h) * D + c];
                P.dhid_h[((size_t)n * H + h) * D + c] = hv > 0.0f ? dlogit[h] * w2 : 0.0f;
                aw_h = fmaf(dlogit[h], hv, aw_h);
            }
            atomicAdd(P.dw2_h + c, aw_h);
        }
        if (P.dhid_o != nullptr) {
            const float w2 = __ldg(P.w2_o + c);
            for (int k = 0; k < O; ++k) {
                const float hv = P.hid_o[((size_t)n * O + k) * D + c];
                P.dhid_o[((size_t)n * O + k) * D + c] = hv > 0.0f ? dlogit[H + k] * w2 : 0.0f;
                aw_o = fmaf(dlogit[H + k], hv, aw_o);
            }
            atomicAdd(P.dw2_o + c, aw_o);
        }
    }
    if (tid == 0) {
        if (P.dhid_h != nullptr) { float v = 0.0f; for (int h = 0; h < H; ++h) v += dlogit[h]; atomicAdd(P.db2_h, v); }
        if (P.dhid_o != nullptr) { float v = 0.0f; for (int k = 0; k < O; ++k) v += dlogit[H + k]; atomicAdd(P.db2_o, v); }
    }
}

int launch_gate_bwd(const FrameBwdParams& P, cudaStream_t stream) {
    TG_REQUIRE(P.H + P.O <= 32, "gate_bwd: at most 32 entities per frame");
    gate_bwd_kernel<<<P.B * P.T, 128, 0, stream>>>(P);
    TG_LAUNCH_OK();
    return 0;
}

__global__ void __launch_bounds__(256) frame_bwd_kernel(const FrameBwdParams P) {
    extern __shared__ __align__(16) float sm[];
    const int D = P.D, H = P.H, O = P.O, T = P.T;
    const int NE = H + O, D2 = 2 * D;
    const int n = blockIdx.x, b = n / T, t = n - b * T;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nkh = P.hh ? 2 : 1;
    const int ts = P.time_position == 1 ? 1 : 0, tu = P.time_position == 2 ? 1 : 0;
    const int gh = P.gh;
    const int wh = (1 + nkh + gh + ts + P.tl) * D, wo = (4 + ts + P.tl) * D;      // rows of xx_h / xx_o (and of their gradients)
    float* sv = sm;                            // [NE][2D]
    float* dmh = sv + NE * D2;                 // [H][nkh*D]   d m_hh | d m_oh
    float* dmo = dmh + H * nkh * D;            // [O][3D]      d m_ho | d m_go | d m_oo
    float* ds = dmo + O * 3 * D;               // [NE][2D]
    __shared__ float a_hh[FB_MAXE * FB_MAXE], a_oh[FB_MAXE * FB_MAXE], a_ho[FB_MAXE * FB_MAXE], a_oo[FB_MAXE * FB_MAXE];
    __shared__ float d_hh[FB_MAXE * FB_MAXE], d_oh[FB_MAXE * FB_MAXE], d_ho[FB_MAXE * FB_MAXE], d_oo[FB_MAXE * FB_MAXE];
    __shared__ float cm[2 * FB_MAXE * 2 * FB_MAXE];
    __shared__ float om[FB_MAXE], dlogit[2 * FB_MAXE];

    for (int i = tid; i < H * D2 / 4; i += 256)
        reinterpret_cast<float4*>(sv)[i] = __ldg(reinterpret_cast<const float4*>(P.s_h + (size_t)n * H * D2) + i);
    for (int i = tid; i < O * D2 / 4; i += 256)
        reinterpret_cast<float4*>(sv + H * D2)[i] = __ldg(reinterpret_cast<const float4*>(P.s_o + (size_t)n * O * D2) + i);
    if (tid < O) om[tid] = P.om[b * O + tid];
    {   // saved attention weights
        const float* al = P.alpha + (size_t)n * (H * H + 2 * H * O + O * O);
        for (int i = tid; i < H * H; i += 256) a_hh[(i / H) * FB_MAXE + i % H] = P.hh ? al[i] : 0.0f;
        for (int i = tid; i < H * O; i += 256) a_oh[(i / O) * FB_MAXE + i % O] = al[H * H + i];
        for (int i = tid; i < O * H; i += 256) a_ho[(i / H) * FB_MAXE + i % H] = al[H * H + H * O + i];
        for (int i = tid; i < O * O; i += 256) a_oo[(i / O) * FB_MAXE + i % O] = al[H * H + 2 * H * O + i];
    }
    // ---- gates: straight-through / filter rule, Gumbel-sigmoid, sigmoid -----------------------------------------
    // (two-layer gate MLPs: gate_bwd_kernel did this already and the gradient of the gate INPUTS arrives in dgin_h / dgin_o)
    if (tid < 32) {                             // NE <= 32: warp 0, one lane per entity
        const float dl_ = P.gate_layers == 2 ? 0.0f : gate_dlogit(P, n, b, t, tid);
        if (tid < NE) dlogit[tid] = dl_;
    }
    __syncthreads();
    // gradient that entity e's gate hands to column `col` of its gate input: one layer = dlogit x weight, two layers = a row of dgin
    auto gin_h = [&](int h, int col) -> float {
        if (P.gate_layers == 2) return P.dgin_h != nullptr ? P.dgin_h[((size_t)n * H + h) * P.gin_h + col] : 0.0f;
        return dlogit[h] * __ldg(P.w_uh + col);
    };
    auto gin_o = [&](int k, int col) -> float {
        if (P.gate_layers == 2) return P.dgin_o != nullptr ? P.dgin_o[((size_t)n * O + k) * P.gin_o + col] : 0.0f;
        return P.w_uo != nullptr ? dlogit[H + k] * __ldg(P.w_uo + col) : 0.0f;
    };
    // ---- gradient of the aggregated messages and the direct parts of d[x|h] ----------------------------------------
    for (int i = tid; i < H * D; i += 256) {
        const int h = i / D, c = i - h * D;
        const float* dx = P.dxx_h + ((size_t)n * H + h) * wh;
        ds[h * D2 + c] = gin_h(h, c);
        ds[h * D2 + D + c] = gin_h(h, D + c) + dx[c];
        if (P.hh) dmh[h * nkh * D + c] = dx[D + c] + gin_h(h, D2 + c);
        dmh[h * nkh * D + (nkh - 1) * D + c] = dx[nkh * D + c] + gin_h(h, D2 + (nkh - 1) * D + c);
    }
    for (int i = tid; i < O * D; i += 256) {
        const int k = i / D, c = i - k * D;
        const float* dx = P.dxx_o + ((size_t)n * O + k) * wo;
        ds[(H + k) * D2 + c] = gin_o(k, c);                                             // ('sah': no object gate MLP -> 0)
        ds[(H + k) * D2 + D + c] = gin_o(k, D + c) + dx[c];
        dmo[k * 3 * D + c] = dx[D + c] + gin_o(k, D2 + c);                              // m_ho   (gate order x,h,m_ho,m_oo,m_go)
        dmo[k * 3 * D + D + c] = dx[2 * D + c] + gin_o(k, D2 + 2 * D + c);              // m_go
        dmo[k * 3 * D + 2 * D + c] = dx[3 * D + c] + gin_o(k, D2 + D + c);              // m_oo
    }
    // time-position features (add_time_position): gradient of this frame's feature vector — strategy 's': the time blocks of all xx
    // rows; strategy 'u': through the last block of the gate inputs, whose weight gradient (one-layer gates) is formed here too
    if (P.time_position != 0 && (tu || P.dtime != nullptr)) {
        const float* te = P.time_emb + (size_t)n * D;
        for (int c = tid; c < D; c += 256) {
            float v = 0.0f;
            if (ts) {
                for (int h = 0; h < H; ++h) v += P.dxx_h[((size_t)n * H + h) * wh + (1 + nkh + gh) * D + c];
                for (int k = 0; k < O; ++k) v += P.dxx_o[((size_t)n * O + k) * wo + 4 * D + c];
            } else {
                if (P.human_seg == nullptr) for (int h = 0; h < H; ++h) v += gin_h(h, D2 + (nkh + gh) * D + c);
                if (P.object_seg == nullptr) for (int k = 0; k < O; ++k) v += gin_o(k, 5 * D + c);
                if (P.gate_layers != 2) {
                    float gsum_h = 0.0f, go = 0.0f;
                    for (int h = 0; h < H; ++h) gsum_h += dlogit[h];
                    for (int k = 0; k < O; ++k) go += dlogit[H + k];
                    if (P.human_seg == nullptr) atomicAdd(P.dw_uh + D2 + (nkh + gh) * D + c, gsum_h * __ldg(te + c));
                    if (P.object_seg == nullptr && P.dw_uo != nullptr) atomicAdd(P.dw_uo + 5 * D + c, go * __ldg(te + c));
                }
            }
            if (P.dtime != nullptr) P.dtime[(size_t)n * D + c] = v;
        }
    }
    // geometry -> human message (one sender, weight 1): gradient from every human's xx row and gate input
    if (gh) {
        for (int c = tid; c < D; c += 256) {
            float v = 0.0f;
            for (int h = 0; h < H; ++h) {
                v += P.dxx_h[((size_t)n * H + h) * wh + (1 + nkh) * D + c];
                if (P.human_seg == nullptr) v += gin_h(h, D2 + nkh * D + c);
            }
            P.dmsg_gh[(size_t)n * D + c] = v;
        }
    }
    // gate weight gradients (one-layer gates): d w[k] += sum_e dlogit[e] * input_e[k]
    if (P.human_seg == nullptr && P.gate_layers != 2) {
        for (int k = tid; k < D2 + (nkh + gh) * D; k += 256) {      // xx_h row = [h, m_hh, m_oh, m_gh ..]: the gate's message blocks in order
            float v = 0.0f;
            for (int h = 0; h < H; ++h) {
                const float in = k < D2 ? sv[h * D2 + k] : __ldg(P.xx_h + ((size_t)n * H + h) * wh + D + (k - D2));
                v = fmaf(dlogit[h], in, v);
            }
            atomicAdd(P.dw_uh + k, v);
        }
        if (tid == 0) { float v = 0.0f; for (int h = 0; h < H; ++h) v += dlogit[h]; atomicAdd(P.db_uh, v); }
    }
    if (P.object_seg == nullptr && P.dw_uo != nullptr && P.gate_layers != 2) {
        for (int k = tid; k < 5 * D; k += 256) {
            float v = 0.0f;
            for (int o = 0; o < O; ++o) {
                float in;
                if (k < D2) in = sv[(H + o) * D2 + k];
                else {
                    const int part = (k - D2) / D, c = (k - D2) - part * D;          // 0: m_ho, 1: m_oo, 2: m_go
                    const int xoff = part == 0 ? D : (part == 1 ? 3 * D : 2 * D);    // xx_o = [h, m_ho, m_go, m_oo]
                    in = __ldg(P.xx_o + ((size_t)n * O + o) * wo + xoff + c);
                }
                v = fmaf(dlogit[H + o], in, v);
            }
            atomicAdd(P.dw_uo + k, v);
        }
        if (tid == 0) { float v = 0.0f; for (int o = 0; o < O; ++o) v += dlogit[H + o]; atomicAdd(P.db_uo, v); }
    }
    __syncthreads();
    // ---- d alpha: one warp per (receiver, sender) pair -----------------------------------------------------------------
    {
        const float* g_hh = P.msg_hh + (size_t)n * H * D;
        const float* g_ho = P.msg_ho + (size_t)n * H * D;
        const float* g_oh = P.msg_oh + (size_t)n * O * D;
        const float* g_oo = P.msg_oo + (size_t)n * O * D;
        const int n_hh = P.hh ? H * H : 0, n_oh = H * O, n_ho = O * H, n_oo = O * O;
        for (int p = warp; p < n_hh + n_oh + n_ho + n_oo; p += 8) {
            const float* a; const float* m; float* dst; float f = 1.0f;
            int q = p;
            if (q < n_hh) { const int h = q / H, j = q % H; a = dmh + h * nkh * D; m = g_hh + j * D; dst = &d_hh[h * FB_MAXE + j]; if (j == h) f = 0.0f; }
            else if ((q -= n_hh) < n_oh) { const int h = q / O, k = q % O; a = dmh + h * nkh * D + (nkh - 1) * D; m = g_oh + k * D; dst = &d_oh[h * FB_MAXE + k]; f = om[k]; }
            else if ((q -= n_oh) < n_ho) { const int k = q / H, h = q % H; a = dmo + k * 3 * D; m = g_ho + h * D; dst = &d_ho[k * FB_MAXE + h]; f = om[k]; }
            else { q -= n_ho; const int k = q / O, j = q % O; a = dmo + k * 3 * D + 2 * D; m = g_oo + j * D; dst = &d_oo[k * FB_MAXE + j]; f = (j == k) ? 0.0f : om[j]; }
            float acc = 0.0f;
            for (int c = lane; c < D; c += 32) acc = fmaf(a[c], __ldg(m + c), acc);
            acc = warp_sum(acc);
            if (lane == 0) *dst = acc * f;
        }
    }
    __syncthreads();
    // ---- softmax backward -> d logits (kept in the d_* arrays), scale 1/sqrt(2D) -------------------------------------------
    if (tid < 2 * NE) {
        const int e = tid % NE;
        const bool second = tid >= NE, recv_h = e < H;
        const int r = recv_h ? e : e - H;
        const int Es = second ? O : H;
        float* al = recv_h ? (second ? a_oh : a_hh) : (second ? a_oo : a_ho);
        float* dd = recv_h ? (second ? d_oh : d_hh) : (second ? d_oo : d_ho);
        if (!(recv_h && !second && !P.hh)) {
            const float scale = P.att_noscale ? 1.0f : 1.0f / sqrtf((float)D2);
            float dot = 0.0f;
            for (int sd = 0; sd < Es; ++sd) dot = fmaf(al[r * FB_MAXE + sd], dd[r * FB_MAXE + sd], dot);
            for (int sd = 0; sd < Es; ++sd)            // mean pooling: the weights do not depend on the states
                dd[r * FB_MAXE + sd] = (P.mean_pool || P.dist_kind[recv_h ? (second ? 1 : 0) : (second ? 3 : 2)])
                                           ? 0.0f : al[r * FB_MAXE + sd] * (dd[r * FB_MAXE + sd] - dot) * scale;
        }
    }
    __syncthreads();
    // symmetric coefficient matrix of the logit gradients: d s[e1] += sum_e2 cm[e1][e2] s[e2]
    for (int i = tid; i < NE * NE; i += 256) {
        const int e1 = i / NE, e2 = i - e1 * NE;
        float v = 0.0f;
        if (e1 < H && e2 < H) { if (P.hh && e1 != e2) v = d_hh[e1 * FB_MAXE + e2] + d_hh[e2 * FB_MAXE + e1]; }
        else if (e1 < H) v = d_oh[e1 * FB_MAXE + (e2 - H)] + d_ho[(e2 - H) * FB_MAXE + e1];
        else if (e2 < H) v = d_oh[e2 * FB_MAXE + (e1 - H)] + d_ho[(e1 - H) * FB_MAXE + e2];
        else if (e1 != e2) v = d_oo[(e1 - H) * FB_MAXE + (e2 - H)] + d_oo[(e2 - H) * FB_MAXE + (e1 - H)];
        cm[e1 * 2 * FB_MAXE + e2] = v;
    }
    __syncthreads();
    // ---- outputs: d messages per sender, d[x|h] per entity ---------------------------------------------------------------
    for (int i = tid; i < H * D; i += 256) {
        const int h = i / D, c = i - h * D;
        if (P.hh) {
            float v = 0.0f;
            for (int r = 0; r < H; ++r)
                if (r != h) v = fmaf(a_hh[r * FB_MAXE + h], dmh[r * nkh * D + c], v);
            P.dmsg_hh[(size_t)n * H * D + i] = v;
        }
        float w = 0.0f;
        for (int k = 0; k < O; ++k) w = fmaf(om[k] * a_ho[k * FB_MAXE + h], dmo[k * 3 * D + c], w);
        P.dmsg_ho[(size_t)n * H * D + i] = w;
    }
    for (int i = tid; i < O * D; i += 256) {
        const int k = i / D, c = i - k * D;
        float v = 0.0f;
        for (int h = 0; h < H; ++h) v = fmaf(a_oh[h * FB_MAXE + k], dmh[h * nkh * D + (nkh - 1) * D + c], v);
        P.dmsg_oh[(size_t)n * O * D + i] = v * om[k];
        float w = 0.0f;
        for (int r = 0; r < O; ++r)
            if (r != k) w = fmaf(a_oo[r * FB_MAXE + k], dmo[r * 3 * D + 2 * D + c], w);
        P.dmsg_oo[(size_t)n * O * D + i] = w * om[k];
    }
    for (int c = tid; c < D; c += 256) {
        float v = 0.0f;
        for (int k = 0; k < O; ++k) v = fmaf(om[k], dmo[k * 3 * D + D + c], v);
        P.dmsg_go[(size_t)n * D + c] = v;
    }
    for (int i = tid; i < NE * D2; i += 256) {
        const int e = i / D2, c = i - e * D2;
        float v = ds[i];
        for (int e2 = 0; e2 < NE; ++e2) v = fmaf(cm[e * 2 * FB_MAXE + e2], sv[e2 * D2 + c], v);
        if (e < H) P.ds_h[(size_t)n * H * D2 + i] = v;
        else P.ds_o[(size_t)n * O * D2 + (i - H * D2)] = v;
    }
}

// One CTA per 32 columns x a slab of frames: column sums of dtime (.) [emb > 0] (x tau), combined with atomics.
__global__ void __launch_bounds__(256) time_embed_bwd_kernel(const float* __restrict__ dtime, const float* __restrict__ emb,
                                                             const float* __restrict__ steps, float* dw, float* db, int B, int T, int D) {
    __shared__ float sw[8][33], sb[8][33];
    const int c = blockIdx.x * 32 + (threadIdx.x & 31), rl = threadIdx.x >> 5;
    const int N = B * T;
    float aw = 0.0f, ab = 0.0f;
    if (c < D)
        for (int n = blockIdx.y * 8 + rl; n < N; n += gridDim.y * 8) {
            const float g = emb[(size_t)n * D + c] > 0.0f ? dtime[(size_t)n * D + c] : 0.0f;
            const int b = n / T, t = n - b * T;
            aw = fmaf(g, (float)(t + 1) / __ldg(steps + b), aw);
            ab += g;
        }
    sw[rl][threadIdx.x & 31] = aw; sb[rl][threadIdx.x & 31] = ab;
    __syncthreads();
    if (rl == 0 && c < D) {
        for (int r = 1; r < 8; ++r) { aw += sw[r][threadIdx.x]; ab += sb[r][threadIdx.x]; }
        atomicAdd(dw + c, aw);
        atomicAdd(db + c, ab);
    }
}

int launch_time_embed_bwd(const float* dtime, const float* emb, const float* steps, float* dw, float* db, int B, int T, int D,
                          cudaStream_t stream) {
    TG_CUDA_OK(cudaMemsetAsync(dw, 0, sizeof(float) * D, stream));
    TG_CUDA_OK(cudaMemsetAsync(db, 0, sizeof(float) * D, stream));
    const int slabs = min(64, cdiv(B * T, 8));
    time_embed_bwd_kernel<<<dim3(cdiv(D, 32), slabs), 256, 0, stream>>>(dtime, emb, steps, dw, db, B, T, D);
    TG_LAUNCH_OK();
    return 0;
}

// One warp per (video, entity): reverse scan over the frames.
__global__ void __launch_bounds__(128) segment_length_bwd_kernel(const SegLenBwdParams P) {
    const int NE = P.H + P.O, D = P.D, T = P.T;
    const int wi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (wi >= P.B * NE) return;
    const int b = wi / NE, e = wi - b * NE;
    const bool is_h = e < P.H;
    const int E = is_h ? P.H : P.O, r = is_h ? e : e - P.H;
    const int ld = is_h ? P.ldh : P.ldo;
    const float* dxx = (is_h ? P.dxx_h : P.dxx_o) + (ld - D);
    const float* xx = (is_h ? P.xx_h : P.xx_o) + (ld - D);
    const float* hard = is_h ? P.y_hs : P.y_os;
    float* du = is_h ? P.du_h : P.du_o;
    const float st = P.periodic ? 1.0f : __ldg(P.steps + b);
    const int half = D / 2;
    float g_acc = 0.0f;
    for (int t = T - 1; t >= 0; --t) {
        const size_t row = (size_t)(b * T + t) * E + r;
        const float x = P.len[(size_t)(b * T + t) * NE + e];
        float g = 0.0f;
        for (int k = lane; k < D; k += 32) {
            const float dv = dxx[row * ld + k];
            float der;
            if (P.periodic) {
                const float f = __ldg(P.freq + (k < half ? k : k - half)), a = x / f;
                der = (k < half ? cosf(a) : -sinf(a)) / f;
            } else {
                der = xx[row * ld + k] > 0.0f ? __ldg(P.w + k) : 0.0f;
            }
            g = fmaf(dv, der, g);
        }
        g = warp_sum(g);
        if (lane == 0) {
            const float pos = P.periodic ? (float)(t + 1) : (float)(t + 1) / st;
            du[row] += pos * (g + g_acc);
        }
        if (hard[row] != 0.0f) g_acc = -g;             // rel = u x - acc, acc' = u x;   else rel = u x, acc' = acc + u x
    }
}

// segment_length_mlp: dw[k] = sum over rows of dxx[row, k] [emb > 0] len[row], db[k] likewise without len; humans then objects.
__global__ void __launch_bounds__(256) segment_length_wgrad_kernel(const SegLenBwdParams P) {
    __shared__ float sw[8][33], sb[8][33];
    const int D = P.D, NE = P.H + P.O, N = P.B * P.T;
    const int c = blockIdx.x * 32 + (threadIdx.x & 31), rl = threadIdx.x >> 5;
    float aw = 0.0f, ab = 0.0f;
    if (c < D)
        for (int i = blockIdx.y * 8 + rl; i < N * NE; i += gridDim.y * 8) {
            const int n = i / NE, e = i - n * NE;
            const bool is_h = e < P.H;
            const int ld = is_h ? P.ldh : P.ldo;
            const size_t row = is_h ? (size_t)n * P.H + e : (size_t)n * P.O + (e - P.H);
            const size_t off = row * ld + (ld - D) + c;
            const float g = (is_h ? P.xx_h : P.xx_o)[off] > 0.0f ? (is_h ? P.dxx_h : P.dxx_o)[off] : 0.0f;
            aw = fmaf(g, P.len[i], aw);
            ab += g;
        }
    sw[rl][threadIdx.x & 31] = aw; sb[rl][threadIdx.x & 31] = ab;
    __syncthreads();
    if (rl == 0 && c < D) {
        for (int r = 1; r < 8; ++r) { aw += sw[r][threadIdx.x]; ab += sb[r][threadIdx.x]; }
        atomicAdd(P.dw + c, aw);
        atomicAdd(P.db + c, ab);
    }
}

int launch_segment_length_bwd(const SegLenBwdParams& P, cudaStream_t stream) {
    segment_length_bwd_kernel<<<cdiv(P.B * (P.H + P.O) * 32, 128), 128, 0, stream>>>(P);
    TG_LAUNCH_OK();
    if (P.dw != nullptr && P.db != nullptr) {
        TG_CUDA_OK(cudaMemsetAsync(P.dw, 0, sizeof(float) * P.D, stream));
        TG_CUDA_OK(cudaMemsetAsync(P.db, 0, sizeof(float) * P.D, stream));
        const int slabs = min(64, cdiv(P.B * P.T * (P.H + P.O), 8));
        segment_length_wgrad_kernel<<<dim3(cdiv(P.D, 32), slabs), 256, 0, stream>>>(P);
        TG_LAUNCH_OK();
    }
    return 0;
}

int launch_frame_bwd(const FrameBwdParams& P, cudaStream_t stream) {
    TG_REQUIRE(P.H <= FB_MAXE && P.O <= FB_MAXE, "frame_bwd: at most %d humans / objects", FB_MAXE);
    const int nkh = P.hh ? 2 : 1;
    const size_t smem = sizeof(float) * ((size_t)(P.H + P.O) * 4 * P.D + (size_t)P.H * nkh * P.D + (size_t)P.O * 3 * P.D);
    TG_REQUIRE(smem <= 200 * 1024, "frame_bwd: shape needs %zu bytes of shared memory", smem);
    if (int rc = ensure_smem((const void*)frame_bwd_kernel, smem)) return rc;
    frame_bwd_kernel<<<P.B * P.T, 256, smem, stream>>>(P);
    TG_LAUNCH_OK();
    return 0;
}

}  // namespace tg
