// Frame-level graph kernels (K-D, K-E, K-G):
//   frame_messages_kernel : per (video, frame) — scaled-dot-product attention logits between the
//       [x | h] features of all humans and objects, masked softmax, aggregation of the per-sender
//       messages, objects_mask multiplies, the Gumbel-sigmoid segmentation gates, and the concatenated
//       segment-level inputs xx_h / xx_o.     (vhoi/models.py:664-749, :1004-1533, :1693-1754;
//       pyrutils/torch/distributions.py:4-36)
//   gate_post_kernel      : optional local-maximum filter of the soft gates (models.py:1637-1664) and
//       the "next segment end" gather index of reorder_hidden_states (models.py:1567-1586)
//   heads_kernel          : Linear(2D->C) + LogSoftmax + (B,T,E,C)->(B,C,T,E) (models.py:909-917)
// All of it is memory-bound elementwise/reduction work: one pass over the frame's rows, everything
// else in shared memory / registers.
#include "common.cuh"
#include "frame.h"

namespace tg {

constexpr int FM_THREADS = 256;
constexpr int FM_MAXE = 16;

// Decision of one gate from its logit (discrete_estimator, models.py:1620-1627): 'gs' = Gumbel-sigmoid sample + threshold
// (distributions.py:4-36), 'st' = the probability itself + threshold.  pos = index of the entity among the sampled ones.
__device__ __forceinline__ void gate_decide(const FrameMsgParams& P, float logit, int b, int t, int n_sampled, int pos, float& y,
                                            float& z, float& p) {
    p = 1.0f / (1.0f + expf(-logit));
    if (P.straight_through) {
        y = p;
        z = p > P.thr ? 1.0f : 0.0f;
        return;
    }
    const float* g = P.noise + ((size_t)(t * n_sampled + pos) * P.B + b) * 2;
    const float la = logf(p + 1e-20f) + __ldg(g);
    const float lb = logf((1.0f - p) + 1e-20f) + __ldg(g + 1);
    const float mx = fmaxf(la, lb);
    const float ea = expf(la - mx), eb = expf(lb - mx);
    y = ea / (ea + eb);
    z = y > P.thr ? 1.0f : 0.0f;
}

// One warp publishes the gates of entity e of frame n = (b, t): imposed segmentation, or the sampled decision with the
// object_segment_update_strategy rules (models.py:1523-1532).  logit_of(entity) is evaluated by all lanes of the warp.
template <typename LogitFn>
__device__ __forceinline__ void gate_publish(const FrameMsgParams& P, int n, int b, int t, int e, int lane, LogitFn logit_of) {
    const int H = P.H, O = P.O, T = P.T, NE = H + O;
    const int strat = P.update_strategy;
    const int n_sampled = (P.human_seg ? 0 : H) + ((P.object_seg || strat == 1) ? 0 : O);
    const bool is_h = e < H;
    const int r = is_h ? e : e - H;
    const float* given = is_h ? P.human_seg : P.object_seg;
    float* y_hard = is_h ? P.y_hs : P.y_os;
    float* y_soft = is_h ? P.y_hss : P.y_oss;
    const int E = is_h ? H : O;
    const size_t oi = (size_t)(b * T + t) * E + r;
    if (given != nullptr) {
        if (lane == 0) { const float v = __ldg(given + oi); y_hard[oi] = v; y_soft[oi] = v; }
        return;
    }
    auto pos_of = [&](int ent) { return ent < H ? ent : (P.human_seg ? 0 : H) + (ent - H); };
    const int src = (!is_h && strat == 1) ? 0 : e;     // 'sah': the object publishes the HUMAN's gate (no object MLP, no noise of its own)
    float y, z, p;
    gate_decide(P, logit_of(src), b, t, n_sampled, pos_of(src), y, z, p);
    float human_hard = 1.0f;
    if (!is_h && strat == 2) {                         // 'coh': hard decision x the human's
        float yh, zh, ph;
        gate_decide(P, logit_of(0), b, t, n_sampled, 0, yh, zh, ph);
        human_hard = P.straight_through ? zh : (zh - yh) + yh;
    }
    if (lane == 0) {
        if (P.pgate_save != nullptr) P.pgate_save[(size_t)n * NE + e] = p;
        float hard = (P.straight_through ? z : (z - y) + y) * human_hard;  // straight-through value, distributions.py:35
        if (t == T - 1) hard = 1.0f;                   // models.py:701-702, :744-745
        y_soft[oi] = y;
        y_hard[oi] = hard;
    }
}

__global__ void __launch_bounds__(FM_THREADS) frame_messages_kernel(const FrameMsgParams P) {
    extern __shared__ __align__(16) float smem[];
    const int D = P.D, H = P.H, O = P.O, T = P.T;
    const int NE = H + O, D2 = 2 * D;
    const int n = blockIdx.x, b = n / T, t = n - b * T;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nkh = P.hh ? 2 : 1;

    float* sv = smem;                         // [NE][2D]   [x | h] of humans then objects
    float* mh = sv + NE * D2;                 // [H][nkh*D] m_hh, m_oh
    float* mo = mh + H * nkh * D;             // [O][3D]    m_ho, m_go, m_oo
    __shared__ float gram[(FM_MAXE * 2) * (FM_MAXE * 2)];
    __shared__ float a_hh[FM_MAXE * FM_MAXE], a_oh[FM_MAXE * FM_MAXE], a_ho[FM_MAXE * FM_MAXE], a_oo[FM_MAXE * FM_MAXE];
    __shared__ float om[FM_MAXE];

    // -- stage rows ---------------------------------------------------------------------------------
    for (int i = tid; i < H * D2 / 4; i += FM_THREADS)
        reinterpret_cast<float4*>(sv)[i] = __ldg(reinterpret_cast<const float4*>(P.s_h + (size_t)n * H * D2) + i);
    for (int i = tid; i < O * D2 / 4; i += FM_THREADS)
        reinterpret_cast<float4*>(sv + H * D2)[i] = __ldg(reinterpret_cast<const float4*>(P.s_o + (size_t)n * O * D2) + i);
    if (tid < O) om[tid] = P.om[b * O + tid];
    __syncthreads();
    // -- Gram matrix of all entity pairs, one warp per pair --------------------------------------------
    const float scale = P.att_noscale ? 1.0f : 1.0f / sqrtf((float)D2);
    for (int pidx = warp; pidx < NE * NE; pidx += FM_THREADS / 32) {
        const int e1 = pidx / NE, e2 = pidx - e1 * NE;
        if (e2 <= e1) continue;
        const float* x = sv + e1 * D2;
        const float* y = sv + e2 * D2;
        float acc = 0.0f;
        for (int k = lane * 4; k < D2; k += 128) {
            const float4 u = *reinterpret_cast<const float4*>(x + k);
            const float4 v = *reinterpret_cast<const float4*>(y + k);
            acc = fmaf(u.x, v.x, acc); acc = fmaf(u.y, v.y, acc); acc = fmaf(u.z, v.z, acc); acc = fmaf(u.w, v.w, acc);
        }
        acc = warp_sum(acc) * scale;
        if (lane == 0) { gram[e1 * NE + e2] = acc; gram[e2 * NE + e1] = acc; }
    }
    __syncthreads();
    // -- masked softmaxes (models.py:1750-1753): one thread per receiver and kind ---------------------
    if (tid < 2 * NE) {
        const int e = tid % NE;
        const bool second = tid >= NE;          // humans: 0 -> hh, 1 -> oh ; objects: 0 -> ho, 1 -> oo
        const bool recv_h = e < H;
        const bool send_h = !second;            // hh / ho have human senders
        const int r = recv_h ? e : e - H;
        const int Es = send_h ? H : O;
        float* out = recv_h ? (second ? a_oh : a_hh) : (second ? a_oo : a_ho);
        const bool same = (recv_h == send_h);
        if (!(recv_h && !second && !P.hh)) {
            const int kind = recv_h ? (second ? 1 : 0) : (second ? 3 : 2);
            float m = -INFINITY;
            for (int sdr = 0; sdr < Es; ++sdr) {
                bool ok = !(same && sdr == r);
                if (!send_h) ok = ok && (om[sdr] != 0.0f);
                float lg; bool dv;
                if (!P.mean_pool && dist_logit(P.dist, kind, (size_t)n, H, O, r, sdr, lg, dv)) {   // models.py:1757-1775
                    gram[e * NE + (send_h ? sdr : H + sdr)] = lg;      // entry (receiver e, this sender) is read by this thread only
                    ok = ok && dv;
                    if (!dv) out[r * FM_MAXE + sdr] = -1.0f;           // remembered for the second pass
                    else out[r * FM_MAXE + sdr] = 0.0f;
                } else {
                    out[r * FM_MAXE + sdr] = 0.0f;
                }
                if (ok) m = fmaxf(m, gram[e * NE + (send_h ? sdr : H + sdr)]);
            }
            float sum = 0.0f;
            for (int sdr = 0; sdr < Es; ++sdr) {
                bool ok = !(same && sdr == r) && out[r * FM_MAXE + sdr] == 0.0f;
                if (!send_h) ok = ok && (om[sdr] != 0.0f);
                // mean pooling (models.py:1033-1036): every valid sender weighs 1 / #valid
                const float ex = ok ? (P.mean_pool ? 1.0f : expf(gram[e * NE + (send_h ? sdr : H + sdr)] - m)) : 0.0f;
                out[r * FM_MAXE + sdr] = ex;
                sum += ex;
            }
            const float inv = sum > 0.0f ? 1.0f / sum : 0.0f;
            for (int sdr = 0; sdr < Es; ++sdr) {
                const float a = out[r * FM_MAXE + sdr] * inv;
                out[r * FM_MAXE + sdr] = a;
                if (recv_h && second && P.att_frame != nullptr) P.att_frame[((size_t)(b * H + r) * T + t) * O + sdr] = a;
                if (P.alpha_save != nullptr) {
                    const int off = recv_h ? (second ? H * H : 0) : (second ? H * H + 2 * H * O : H * H + H * O);
                    P.alpha_save[(size_t)n * (H * H + 2 * H * O + O * O) + off + r * Es + sdr] = a;
                }
            }
        }
    }
    __syncthreads();
    // -- aggregated messages; also the concatenated segment-level inputs -------------------------------
    {
        const float* g_hh = P.msg_hh + (size_t)n * H * D;
        const float* g_ho = P.msg_ho + (size_t)n * H * D;
        const float* g_oh = P.msg_oh + (size_t)n * O * D;
        const float* g_oo = P.msg_oo + (size_t)n * O * D;
        const float* g_go = P.msg_go + (size_t)n * D;
        const int ts = P.time_position == 1 ? 1 : 0;
        const int gh = P.gh;
        const int wh = (1 + nkh + gh + ts + P.tl) * D;  // xx_h row: [h, (m_hh), m_oh (, m_gh) (, time) (, length)]
        const int wo = (4 + ts + P.tl) * D;      // xx_o row: [h, m_ho, m_go, m_oo (, time) (, length)]
        float* xxh = P.xx_h + (size_t)n * H * wh;
        float* xxo = P.xx_o + (size_t)n * O * wo;
        const float* te = ts ? P.time_emb + (size_t)n * D : nullptr;
        // four columns per thread (float4 loads of the sender messages, float4 stores of xx rows); per element the same fmaf
        // order as a scalar loop over the senders
        const int D4 = D / 4;
        auto fma4 = [](float a, const float4& m, float4& v) { v.x = fmaf(a, m.x, v.x); v.y = fmaf(a, m.y, v.y); v.z = fmaf(a, m.z, v.z); v.w = fmaf(a, m.w, v.w); };
        auto ld4 = [](const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); };
        auto scale4 = [](const float4& m, float s) { return make_float4(m.x * s, m.y * s, m.z * s, m.w * s); };
        for (int idx = tid; idx < H * D4; idx += FM_THREADS) {
            const int h = idx / D4, c = (idx - h * D4) * 4;
            float* row = xxh + h * wh;
            *reinterpret_cast<float4*>(row + c) = *reinterpret_cast<const float4*>(sv + h * D2 + D + c);
            if (P.hh) {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int j = 0; j < H; ++j)
                    if (j != h) fma4(a_hh[h * FM_MAXE + j], ld4(g_hh + j * D + c), v);
                *reinterpret_cast<float4*>(mh + h * nkh * D + c) = v;
                *reinterpret_cast<float4*>(row + D + c) = v;
            }
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int k = 0; k < O; ++k) fma4(a_oh[h * FM_MAXE + k], scale4(ld4(g_oh + k * D + c), om[k]), v);
            *reinterpret_cast<float4*>(mh + h * nkh * D + (nkh - 1) * D + c) = v;
            *reinterpret_cast<float4*>(row + nkh * D + c) = v;
            if (gh) *reinterpret_cast<float4*>(row + (1 + nkh) * D + c) = ld4(P.msg_gh + (size_t)n * D + c);   // models.py:694, :705
            if (ts) *reinterpret_cast<float4*>(row + (1 + nkh + gh) * D + c) = ld4(te + c);  // models.py:761
        }
        for (int idx = tid; idx < O * D4; idx += FM_THREADS) {
            const int k = idx / D4, c = (idx - k * D4) * 4;
            float* row = xxo + k * wo;
            *reinterpret_cast<float4*>(row + c) = *reinterpret_cast<const float4*>(sv + (H + k) * D2 + D + c);
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int h = 0; h < H; ++h) fma4(a_ho[k * FM_MAXE + h], ld4(g_ho + h * D + c), v);
            v = scale4(v, om[k]);                                  // models.py:720
            const float4 go = scale4(ld4(g_go + c), om[k]);        // single sender, alpha = 1; models.py:729
            float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int j = 0; j < O; ++j)
                if (j != k) fma4(a_oo[k * FM_MAXE + j], scale4(ld4(g_oo + j * D + c), om[j]), w);
            *reinterpret_cast<float4*>(mo + k * 3 * D + c) = v;
            *reinterpret_cast<float4*>(mo + k * 3 * D + D + c) = go;
            *reinterpret_cast<float4*>(mo + k * 3 * D + 2 * D + c) = w;
            *reinterpret_cast<float4*>(row + D + c) = v;           // models.py:748 order: h, m_ho, m_go, m_oo
            *reinterpret_cast<float4*>(row + 2 * D + c) = go;
            *reinterpret_cast<float4*>(row + 3 * D + c) = w;
            if (ts) *reinterpret_cast<float4*>(row + 4 * D + c) = ld4(te + c);               // models.py:762
        }
    }
    __syncthreads();
    // -- segmentation gates, one warp per entity -----------------------------------------------------------
    if (P.gate_in_h != nullptr) {
        // discrete_networks_num_layers == 2: materialise the gate MLP inputs in the MLPs' own column order (models.py:1494, :1527);
        // the hidden layer is a projection GEMM and gate_sample_kernel finishes.  Imposed segmentations are copied here as before.
        for (int e = warp; e < NE; e += FM_THREADS / 32) {
            const bool is_h = e < H;
            const int r = is_h ? e : e - H, E = is_h ? H : O;
            const float* given = is_h ? P.human_seg : P.object_seg;
            const size_t oi = (size_t)(b * T + t) * E + r;
            if (given != nullptr) {
                if (lane == 0) { const float v = __ldg(given + oi); (is_h ? P.y_hs : P.y_os)[oi] = v; (is_h ? P.y_hss : P.y_oss)[oi] = v; }
                continue;
            }
            float* row = is_h ? P.gate_in_h + ((size_t)n * H + r) * P.gin_h : P.gate_in_o + ((size_t)n * O + r) * P.gin_o;
            const float* s = sv + e * D2;
            for (int k = lane; k < D2; k += 32) row[k] = s[k];
            int col = D2;
            if (is_h) {
                const float* m = mh + r * nkh * D;
                for (int k = lane; k < nkh * D; k += 32) row[col + k] = m[k];
                col += nkh * D;
                if (P.gh) { for (int k = lane; k < D; k += 32) row[col + k] = __ldg(P.msg_gh + (size_t)n * D + k); col += D; }
            } else {
                const float* m = mo + r * 3 * D;          // smem order m_ho, m_go, m_oo -> gate order m_ho, m_oo, m_go
                for (int k = lane; k < D; k += 32) { row[col + k] = m[k]; row[col + D + k] = m[2 * D + k]; row[col + 2 * D + k] = m[D + k]; }
                col += 3 * D;
            }
            if (P.time_position == 2) for (int k = lane; k < D; k += 32) row[col + k] = __ldg(P.time_emb + (size_t)n * D + k);
        }
        return;
    }
    // update_strategy (models.py:1523-1532, one human only): 'sah' — the object warps evaluate the HUMAN's gate and publish it as
    // their own (no object MLP, no noise of their own); 'coh' — they evaluate both and multiply the hard decisions.
    // logit of entity e's one-layer gate MLP (all lanes call; every lane holds the result)
    auto gate_logit = [&](int e) -> float {
        const bool is_h = e < H;
        const int r = is_h ? e : e - H;
        const float* w = is_h ? P.w_uh : P.w_uo;
        float acc = 0.0f;
        const float* s = sv + e * D2;
        for (int k = lane; k < D2; k += 32) acc = fmaf(__ldg(w + k), s[k], acc);
        if (is_h) {
            const float* m = mh + r * nkh * D;       // gate input order [x, h, m_hh, m_oh], models.py:1494
            for (int k = lane; k < nkh * D; k += 32) acc = fmaf(__ldg(w + D2 + k), m[k], acc);
            if (P.gh) {
                const float* g = P.msg_gh + (size_t)n * D;
                for (int k = lane; k < D; k += 32) acc = fmaf(__ldg(w + D2 + nkh * D + k), __ldg(g + k), acc);
            }
        } else {
            const float* m = mo + r * 3 * D;          // gate input order [x, h, m_ho, m_oo, m_go], models.py:1527
            for (int k = lane; k < D; k += 32) {
                acc = fmaf(__ldg(w + D2 + k), m[k], acc);
                acc = fmaf(__ldg(w + D2 + D + k), m[2 * D + k], acc);
                acc = fmaf(__ldg(w + D2 + 2 * D + k), m[D + k], acc);
            }
        }
        if (P.time_position == 2) {                   // strategy 'u': the time feature is the last block of the gate input
            const float* wt = w + D2 + (is_h ? nkh + P.gh : 3) * D;
            const float* te = P.time_emb + (size_t)n * D;
            for (int k = lane; k < D; k += 32) acc = fmaf(__ldg(wt + k), __ldg(te + k), acc);
        }
        return warp_sum(acc) + __ldg(is_h ? P.b_uh : P.b_uo);
    };
    for (int e = warp; e < NE; e += FM_THREADS / 32)
        gate_publish(P, n, b, t, e, lane, gate_logit);
}

size_t frame_messages_smem(int H, int O, int D, int hh) {
    const int nkh = hh ? 2 : 1;
    return sizeof(float) * ((size_t)(H + O) * 2 * D + (size_t)H * nkh * D + (size_t)O * 3 * D);
}

int launch_frame_messages(const FrameMsgParams& P, cudaStream_t stream) {
    TG_REQUIRE(P.H <= FM_MAXE && P.O <= FM_MAXE, "frame_messages: at most %d humans / objects", FM_MAXE);
    TG_REQUIRE(P.D % 4 == 0, "frame_messages: hidden_size must be a multiple of 4");
    TG_REQUIRE(P.O >= 2 && (!P.hh || P.H >= 2), "frame_messages: need >=2 objects (and >=2 humans with humans->human)");
    const size_t smem = frame_messages_smem(P.H, P.O, P.D, P.hh);
    TG_REQUIRE(smem <= 200 * 1024, "frame_messages: shape needs %zu bytes of shared memory", smem);
    if (int rc = ensure_smem((const void*)frame_messages_kernel, smem)) return rc;
    frame_messages_kernel<<<P.B * P.T, FM_THREADS, smem, stream>>>(P);
    TG_LAUNCH_OK();
    return 0;
}

// Two-layer gate MLPs: logit = w2 . hidden + b2 from the hidden rows the projection wrote; one warp per (frame, entity).
__global__ void __launch_bounds__(128) gate_sample_kernel(const FrameMsgParams P) {
    const int H = P.H, O = P.O, NE = H + O, D = P.D, T = P.T;
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= P.B * T * NE) return;
    const int n = w / NE, e = w - n * NE, b = n / T, t = n - b * T;
    auto logit_of = [&](int ent) -> float {
        const bool is_h = ent < H;
        const float* hid = is_h ? P.gate_hid_h + ((size_t)n * H + ent) * D : P.gate_hid_o + ((size_t)n * O + (ent - H)) * D;
        const float* w2 = is_h ? P.w_uh : P.w_uo;
        float acc = 0.0f;
        for (int k = lane; k < D; k += 32) acc = fmaf(__ldg(w2 + k), __ldg(hid + k), acc);
        return warp_sum(acc) + __ldg(is_h ? P.b_uh : P.b_uo);
    };
    const float* given = e < H ? P.human_seg : P.object_seg;
    if (given != nullptr) return;                      // the frame kernel copied the imposed segmentation
    gate_publish(P, n, b, t, e, lane, logit_of);
}

int launch_gate_sample(const FrameMsgParams& P, cudaStream_t stream) {
    const int warps = P.B * P.T * (P.H + P.O);
    gate_sample_kernel<<<cdiv(warps * 32, 128), 128, 0, stream>>>(P);
    TG_LAUNCH_OK();
    return 0;
}

// Time-position features of every frame: one thread per (frame, column).
__global__ void time_embed_kernel(const float* __restrict__ steps, const float* __restrict__ w, const float* __restrict__ bias,
                                  const float* __restrict__ freq, float* __restrict__ out, int B, int T, int D, int periodic) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)B * T * D) return;
    const int k = (int)(i % D);
    const size_t n = i / D;
    const int b = (int)(n / T), t = (int)(n - (size_t)b * T);
    const float pos = (float)(t + 1);
    float v;
    if (periodic) {                                  // make_periodic_embedding, models.py:1777-1794 (no division by the steps, :654)
        const int half = D / 2;
        const float x = pos / __ldg(freq + (k < half ? k : k - half));
        v = k < half ? sinf(x) : cosf(x);
    } else {                                         // _assemble_time_tensor + Linear(1, D) + ReLU, models.py:936-952, :259-260
        const float tau = pos / __ldg(steps + b);
        v = fmaxf(fmaf(__ldg(w + k), tau, __ldg(bias + k)), 0.0f);
    }
    out[i] = v;
}

int launch_time_embed(const float* steps, const float* w, const float* bias, const float* freq, float* out, int B, int T, int D,
                      int periodic, cudaStream_t stream) {
    const size_t total = (size_t)B * T * D;
    time_embed_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(steps, w, bias, freq, out, B, T, D, periodic);
    TG_LAUNCH_OK();
    return 0;
}

// Segment lengths: one thread per (video, entity) scans the frames (models.py:954-979).
__global__ void segment_length_scan_kernel(const float* __restrict__ y_hs, const float* __restrict__ y_os,
                                           const float* __restrict__ steps, float* __restrict__ len, int B, int T, int H, int O,
                                           int periodic) {
    const int NE = H + O;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * NE) return;
    const int b = i / NE, e = i - b * NE;
    const bool is_h = e < H;
    const int E = is_h ? H : O, r = is_h ? e : e - H;
    const float* hard = is_h ? y_hs : y_os;
    const float st = periodic ? 1.0f : __ldg(steps + b);
    float acc = 0.0f;
    for (int t = 0; t < T; ++t) {
        const float x = periodic ? (float)(t + 1) : (float)(t + 1) / st;
        float rel = hard[(size_t)(b * T + t) * E + r] * x;
        if (rel != 0.0f) rel -= acc;
        acc += rel;
        len[(size_t)(b * T + t) * NE + e] = rel;
    }
}

// Embedding of the lengths into the last D columns of the xx rows: one thread per (frame, entity, column).
__global__ void segment_length_embed_kernel(const float* __restrict__ len, const float* __restrict__ w, const float* __restrict__ bias,
                                            const float* __restrict__ freq, float* xx_h, int ldh, float* xx_o, int ldo, int N, int H,
                                            int O, int D, int periodic) {
    const int NE = H + O;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)N * NE * D) return;
    const int k = (int)(i % D);
    const size_t ne = i / D;
    const int e = (int)(ne % NE);
    const size_t n = ne / NE;
    const float x = len[ne];
    float v;
    if (periodic) {
        const int half = D / 2;
        const float a = x / __ldg(freq + (k < half ? k : k - half));
        v = k < half ? sinf(a) : cosf(a);
    } else {
        v = fmaxf(fmaf(__ldg(w + k), x, __ldg(bias + k)), 0.0f);
    }
    if (e < H) xx_h[(n * H + e) * ldh + (ldh - D) + k] = v;
    else       xx_o[(n * O + (e - H)) * ldo + (ldo - D) + k] = v;
}

int launch_segment_length(const float* y_hs, const float* y_os, const float* steps, const float* w, const float* bias,
                          const float* freq, float* len, float* xx_h, int ldh, float* xx_o, int ldo, int B, int T, int H, int O,
                          int D, int periodic, cudaStream_t stream) {
    segment_length_scan_kernel<<<cdiv(B * (H + O), 128), 128, 0, stream>>>(y_hs, y_os, steps, len, B, T, H, O, periodic);
    TG_LAUNCH_OK();
    const size_t total = (size_t)B * T * (H + O) * D;
    segment_length_embed_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(len, w, bias, freq, xx_h, ldh, xx_o, ldo, B * T, H, O,
                                                                                    D, periodic);
    TG_LAUNCH_OK();
    return 0;
}

// One warp per (video, entity): filter + reorder index.
__global__ void gate_post_kernel(float* y_hs, const float* y_hss, float* y_os, const float* y_oss, int* reidx, int B, int T,
                                 int H, int O, int filter, float thr) {
    const int NE = H + O;
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= B * NE) return;
    const int b = w / NE, e = w - b * NE;
    const bool is_h = e < H;
    const int E = is_h ? H : O, r = is_h ? e : e - H;
    float* hard = is_h ? y_hs : y_os;
    const float* soft = is_h ? y_hss : y_oss;
    if (filter) {   // hard_t = 1 iff y_t > y_{t-1} and y_t > y_{t+1} and y_t >= thr (zeros beyond the ends)
        for (int t = lane; t < T; t += 32) {
            const float y = soft[(size_t)(b * T + t) * E + r];
            const float yp = t > 0 ? soft[(size_t)(b * T + t - 1) * E + r] : 0.0f;
            const float yn = t + 1 < T ? soft[(size_t)(b * T + t + 1) * E + r] : 0.0f;
            const bool keep = (y > yp) && (y > yn) && (y >= thr);
            const float z = y >= thr ? 1.0f : 0.0f;
            const float u = (z - y) + y;
            hard[(size_t)(b * T + t) * E + r] = keep ? u : fminf(u, 0.0f);
        }
        __syncwarp();
    }
    // idx[t] = min{ e >= t : hard[e] != 0 } if any, else t   (scan from the end, 32 frames at a time)
    int carry = -1;
    for (int base = ((T - 1) / 32) * 32; base >= 0; base -= 32) {
        const int t = base + lane;
        const bool end = t < T && hard[(size_t)(b * T + t) * E + r] != 0.0f;
        const unsigned m = __ballot_sync(0xffffffffu, end);
        if (t < T) {
            const unsigned mm = m & (0xffffffffu << lane);
            int idx;
            if (mm) idx = base + __ffs(mm) - 1;
            else idx = carry >= 0 ? carry : t;
            reidx[(size_t)(b * T + t) * NE + e] = idx;
        }
        if (m) carry = base + __ffs(m) - 1;
    }
}

int launch_gate_post(float* y_hs, const float* y_hss, float* y_os, const float* y_oss, int* reidx, int B, int T, int H,
                     int O, int filter, float thr, cudaStream_t stream) {
    const int warps = B * (H + O);
    gate_post_kernel<<<cdiv(warps * 32, 128), 128, 0, stream>>>(y_hs, y_hss, y_os, y_oss, reidx, B, T, H, O, filter, thr);
    TG_LAUNCH_OK();
    return 0;
}

// One warp per (row, source): two heads share the same input row.
__global__ void __launch_bounds__(256) heads_kernel(const HeadsParams P) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int rows = P.B * P.T * P.E;
    if (w >= rows * 2) return;
    const int src = w / rows;              // 0: frame level (BiGRU outputs), 1: segment level (reordered states)
    const int row = w - src * rows;
    const int e = row % P.E, bt = row / P.E;
    const int t = bt % P.T, b = bt / P.T;
    const int D2 = 2 * P.D, C = P.C;
    const float* x;
    if (src == 0) x = P.hfr + (size_t)row * D2;
    else {
        const int ti = P.reidx[(size_t)bt * P.NE + P.e_off + e];
        x = P.hx + ((size_t)(b * P.T + ti) * P.E + e) * D2;
    }
    const bool cat = P.cat && src == 1;     // models.py:901-903: segment heads read [reordered segment state | frame-level state]
    const float* x2 = P.hfr + (size_t)row * D2;
    const int ldw = cat ? 2 * D2 : D2;
#pragma unroll 1
    for (int hd = 0; hd < 2; ++hd) {
        const float* W = P.w[src * 2 + hd];
        const float* bias = P.b[src * 2 + hd];
        float* out = P.out[src * 2 + hd];
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
            acc = warp_sum(acc) + __ldg(bias + c);
            if (lane == c) mine = acc;
        }
        const float m = warp_max(mine);
        const float ex = lane < C ? expf(mine - m) : 0.0f;
        const float lse = m + logf(warp_sum(ex));
        if (lane < C) out[((size_t)(b * C + lane) * P.T + t) * P.E + e] = mine - lse;
    }
}

int launch_heads(const HeadsParams& P, cudaStream_t stream) {
    TG_REQUIRE(P.C >= 1 && P.C <= 32, "heads: number of classes %d unsupported (1..32)", P.C);
    TG_REQUIRE(P.D % 2 == 0, "heads: hidden_size must be even");
    const long warps = 2L * P.B * P.T * P.E;
    heads_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, stream>>>(P);
    TG_LAUNCH_OK();
    return 0;
}

}  // namespace tg
