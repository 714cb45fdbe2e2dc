// C ABI, backward half: tggcn_backward = what loss.backward() does through TGGCN.forward (vhoi/models.py:584-933)
// and its building blocks.  The forward must have run with dims.save_for_backward on the same workspace.
// Order = reverse of the forward's launch sequence (api.cu): heads -> segment-level recurrent graph (reverse time) ->
// hoisted segment projections -> frame-level graph + gates -> message MLPs -> Linear(2D->D) -> BiGRU BPTT ->
// input projections -> embeddings / geometry MLP -> geometry GCN.
// No allocation, no device synchronisation, no state between calls.
#include "common.cuh"
#include "gemm.h"
#include "bigru.h"
#include "frame.h"
#include "backward.cuh"
#include "api_internal.h"

using namespace tg;


namespace {

// Regions of the backward workspace (floats).
struct BwdLayout {
    size_t total = 0;
    size_t take(size_t n) {
        const size_t o = total;
        total += (n + 63) / 64 * 64;
        return o;
    }
    size_t dhfr[3], dhx[2], dgs[2], dghs[2], du[2], direct[2], dmg[2], dpre[2], lgr[4], lgs[4], dpre_all[4], gru_direct[3];
    size_t dghid[2], dghid1[2], dgin[2], dtime, dxx[2], ds[3], dmsg[6], dgi[3], dgh[3], bigru_scratch, dgeo_hid, dgcn_out, dxn, wt_seg, wt, tn, tn_floats;
    size_t zero_begin, zero_end;     // region that must be zero before the kernels run (atomically accumulated)
};

void make_bwd_layout(const tggcn_dims& d, BwdLayout& L) {
    const size_t N = (size_t)d.B * d.T, B = d.B, H = d.H, O = d.O, D = d.D, V = d.V;
    const size_t nkh = nkh_of(d);
    const size_t E[3] = {H, O, 1};
    L.zero_begin = L.total;
    for (int g = 0; g < 2; ++g) L.dhfr[g] = L.take(N * E[g] * 2 * D);
    for (int g = 0; g < 2; ++g) L.dhx[g] = L.take(N * E[g] * 2 * D);
    for (int g = 0; g < 2; ++g) L.du[g] = L.take(N * E[g]);
    L.zero_end = L.total;
    L.dhfr[2] = L.take(N * 2 * D);
    for (int g = 0; g < 2; ++g) {
        L.dgs[g] = L.take(N * E[g] * 6 * D);
        L.dghs[g] = L.take(N * E[g] * 6 * D);
        L.direct[g] = L.take(2 * B * E[g] * D);
        L.dmg[g] = L.take(2 * B * E[g] * (g == 0 ? nkh : 2) * D);
        L.dpre[g] = L.take(2 * B * E[g] * 2 * D);
    }
    for (int k = 0; k < 4; ++k) {
        const size_t Er = (k == 0 || k == 1) ? H : O, Es = (k == 0 || k == 2) ? H : O;
        L.lgr[k] = L.take(2 * B * Er * D);
        L.lgs[k] = L.take(2 * B * Es * D);
    }
    for (int g = 0; g < 3; ++g) L.gru_direct[g] = L.take(2 * B * E[g] * D);
    L.dpre_all[0] = L.take(2 * N * H * D);
    L.dpre_all[1] = L.take(2 * N * O * D);
    L.dpre_all[2] = L.take(2 * N * H * D);
    L.dpre_all[3] = L.take(2 * N * O * D);
    L.dxx[0] = L.take(N * H * (size_t)kh_of(d));
    L.dxx[1] = L.take(N * O * (size_t)ko_of(d));
    L.dtime = L.take(d.time_position && !d.time_periodic ? N * D : 0);
    L.dghid[0] = L.take(gate2_of(d) ? N * H * D : 0);
    L.dghid[1] = L.take(gate2_of(d) ? N * O * D : 0);
    L.dghid1[0] = L.take(gate3_of(d) ? N * H * D : 0);
    L.dghid1[1] = L.take(gate3_of(d) ? N * O * D : 0);
    L.dgin[0] = L.take(gate2_of(d) ? N * H * (size_t)ginh_of(d) : 0);
    L.dgin[1] = L.take(gate2_of(d) ? N * O * (size_t)gino_of(d) : 0);
    for (int g = 0; g < 3; ++g) L.ds[g] = L.take(N * E[g] * 2 * D);
    L.dmsg[0] = L.take(N * H * D);
    L.dmsg[1] = L.take(N * H * D);
    L.dmsg[2] = L.take(N * O * D);
    L.dmsg[3] = L.take(N * O * D);
    L.dmsg[4] = L.take(N * D);
    L.dmsg[5] = L.take(d.geo_to_human ? N * D : 0);
    for (int g = 0; g < 3; ++g) {
        L.dgi[g] = L.take(N * E[g] * 6 * D);
        L.dgh[g] = L.take(N * E[g] * 6 * D);
    }
    L.bigru_scratch = L.take(6 * D * 3 * D);               // W_hh^T of the three frame-level BiGRUs, both directions
    L.dgeo_hid = L.take(N * 2048);
    L.dgcn_out = L.take(N * 128 * V);
    L.dxn = L.take(N * V * 4);
    // transposed recurrent weights that live through the whole reverse loop:
    // W_hh^T (4 x D x 3D), W_ih[:, seg]^T (2 x nkh*D x 3D + 2 x 2D x 3D), message MLPs (2 x D x 2D)
    L.wt_seg = L.take(4 * D * 3 * D + 2 * nkh * D * 3 * D + 2 * 2 * D * 3 * D + 2 * D * 2 * D);
    // scratch for one transposed projection weight at a time
    size_t wmax = (size_t)2048 * 128 * V;                 // geometry_embedding_mlp.0
    const size_t cands[] = {(size_t)ko_of(d) * 6 * D, (size_t)kh_of(d) * 6 * D, (size_t)2048 * D, 2 * D * D, D * 6 * D};
    for (size_t c : cands)
        if (c > wmax) wmax = c;
    L.wt = L.take(wmax);
    // transposed operands of one weight-gradient GEMM: (N + K) x M rounded up to 64 (fp32 transposes, or the 16-bit operand planes
    // of gemm16.cu — the same 4 bytes per element — which also holds the planes of one dX GEMM: gradient rows + transposed weight)
    const size_t rp = (N * (H > O ? H : O) + 63) / 64 * 64, np = (N + 63) / 64 * 64;
    const size_t ko_ = ko_of(d);
    size_t tmax = (3 * D + (ko_ > 2048 ? ko_ : 2048)) * rp;
    if ((2048 + 128 * V) * np > tmax) tmax = (2048 + 128 * V) * np;
    const size_t nt_cands[] = {rp * 6 * D + ko_ * 6 * D, np * 2048 + (size_t)2048 * 128 * V, np * 2 * D + (size_t)2048 * D};
    for (size_t c : nt_cands)
        if (c > tmax) tmax = c;
    // grouped weight-gradient launches keep every operand of a stage at once: the segment cells of one direction, the BiGRUs of
    // humans + objects (both directions), the embeddings + geometry MLP
    const size_t kh_ = kh_of(d);
    const size_t grp_cands[] = {rp * ((H ? 1 : 0) * (7 * D + kh_ + nkh * D) + 9 * D + ko_), 2 * rp * 15 * D,
                                2 * rp * (D + 2048) + np * (D + 2048) + np * (2048 + 128 * V), 3 * rp * 4 * D + np * 6 * D,
                                gate2_of(d) ? 2 * rp * (3 * D + (size_t)(ginh_of(d) > gino_of(d) ? ginh_of(d) : gino_of(d))) : 0};
    for (size_t c : grp_cands)
        if (c > tmax) tmax = c;
    tmax += 4096;
    L.tn_floats = tmax;
    L.tn = L.take(tmax);
}

// scratch of the 16-bit operand planes (gemm16.cu) and the status word of the call; g16.ws == nullptr: the tf32 / bf16 kernels of gemm_tc.cu
struct G16Ctx {
    void* ws;
    size_t bytes;
    int precision;
    unsigned int* err;
};

// C[M, Nout] (+)= (A (.) [mask > 0]) [M, K] * Wt[Nout, K]^T
int gemm_nt(const float* A, int lda, const float* mask, int ldm, const float* Wt, int ldw, float* C, int ldc, int M, int Nout, int K,
            int beta, int path, cudaStream_t stream, const G16Ctx& g16) {
    if (g16.ws != nullptr) {
        // A is a gradient: amax-scaled fp16 (hi, lo) planes; Wt a transposed weight: times 2^8 as in the forward
        Gemm16Problem q;
        memset(&q, 0, sizeof(q));
        q.a.src = A; q.a.ld = lda; q.a.rows = M; q.a.cols = K; q.a.mask = mask; q.a.ldm = ldm; q.a.dynamic = 1;
        q.b.src = Wt; q.b.ld = ldw; q.b.rows = Nout; q.b.cols = K; q.b.scale = 256.0f;
        q.C = C; q.ldc = ldc; q.beta = beta;
        if (gemm16_eligible(&q, 1) && gemm16_scratch_bytes(&q, 1) <= g16.bytes)
            return launch_gemm16(&q, 1, g16.precision, g16.ws, g16.bytes, g16.err, stream);
    }
    GemmGroup g;
    g.count = 0;
    gemm_add(g, A, lda, Wt, ldw, nullptr, C, ldc, M, Nout, K, 0);
    g.p[0].amask = mask; g.p[0].ldm = ldm; g.p[0].beta = beta;
    return launch_gemm(g, path, stream);
}

}  // namespace

extern "C" {

size_t tggcn_backward_workspace_bytes(const tggcn_dims* dims) {
    if (dims == nullptr || check_dims(*dims)) return 0;
    BwdLayout L;
    make_bwd_layout(*dims, L);
    return L.total * sizeof(float);
}

int tggcn_backward_bucket(int id) {
    if (id < 0 || id >= TGGCN_W_COUNT) return -1;
    if (id >= TGGCN_W_TIME_W && id <= TGGCN_W_UPD_O_B4) return 1;         // time / length MLPs, geometry -> human message MLP, gate MLP layer 2            // formed with the frame-level graph
    if (id <= TGGCN_W_GCN_S2_B) return 3;                                   // GCN_* (first 13 entries of the table)
    if (id <= TGGCN_W_OBJ_EMB_B) return 2;                                  // geometry MLP, ROI embeddings
    if (id >= TGGCN_W_HSEG_F_WIH || (id >= TGGCN_W_SMSG_HH_W && id <= TGGCN_W_SMSG_OO_B)) return 0;   // cells, heads, segment MLPs
    return 1;
}

int tggcn_backward(const tggcn_dims* dims, const void* const* weights, void* const* grad_weights, int n_weights,
                   const tggcn_io* io, const tggcn_grad_outputs* grads, void* workspace, size_t workspace_bytes,
                   void* bwd_workspace, size_t bwd_workspace_bytes, void* stream_) {
    return tggcn_backward_ex(dims, weights, grad_weights, n_weights, io, grads, workspace, workspace_bytes, bwd_workspace,
                             bwd_workspace_bytes, stream_, nullptr);
}

int tggcn_backward_ex(const tggcn_dims* dims, const void* const* weights, void* const* grad_weights, int n_weights,
                      const tggcn_io* io, const tggcn_grad_outputs* grads, void* workspace, size_t workspace_bytes,
                      void* bwd_workspace, size_t bwd_workspace_bytes, void* stream_, const tggcn_bwd_hooks* hooks) {
    TG_REQUIRE(dims && weights && grad_weights && io && grads && workspace && bwd_workspace, "backward: null argument");
    TG_REQUIRE(n_weights == TGGCN_W_COUNT, "backward: expected %d weight pointers, got %d", (int)TGGCN_W_COUNT, n_weights);
    const tggcn_dims& d = *dims;
    if (int rc = check_dims(d)) return rc;
    TG_REQUIRE(d.save_for_backward, "backward: the forward must run with dims.save_for_backward = 1");
    Layout L;
    make_layout(d, L);
    TG_REQUIRE(workspace_bytes >= L.total, "backward: forward workspace too small (%zu < %zu)", workspace_bytes, L.total);
    BwdLayout BL;
    make_bwd_layout(d, BL);
    TG_REQUIRE(bwd_workspace_bytes >= BL.total * sizeof(float), "backward: workspace too small (%zu < %zu)", bwd_workspace_bytes,
               BL.total * sizeof(float));
    TG_REQUIRE((reinterpret_cast<uintptr_t>(bwd_workspace) & 255) == 0, "backward: workspace must be 256-byte aligned");
    cudaStream_t stream = (cudaStream_t)stream_;
    const int B = d.B, T = d.T, H = d.H, O = d.O, D = d.D, V = d.V, N = B * T;
    const int nkh = nkh_of(d), nks = nkh;
    const int path = (d.precision == 1 && d.gemm_path != 0) ? 3 : d.gemm_path;      // bf16 operands on the tensor-core GEMMs
    auto W = [&](int id) { return (const float*)weights[id]; };
    auto G = [&](int id) { return (float*)grad_weights[id]; };
    auto buf = [&](int id) { return (float*)((char*)workspace + L.off[id]); };
    float* bw = (float*)bwd_workspace;
    auto bb = [&](size_t off) { return bw + off; };

    // every parameter on the gradient path needs a destination
    {
        static const int need[] = {
            TGGCN_W_GCN_W, TGGCN_W_GCN_BN_W, TGGCN_W_GCN_BN_B, TGGCN_W_GCN_C1_W, TGGCN_W_GCN_C1_B, TGGCN_W_GCN_C3_W, TGGCN_W_GCN_C3_B,
            TGGCN_W_GCN_S1_W, TGGCN_W_GCN_S1_B, TGGCN_W_GCN_S2_W, TGGCN_W_GCN_S2_B, TGGCN_W_GEO_MLP0_W, TGGCN_W_GEO_MLP0_B,
            TGGCN_W_GEO_MLP2_W, TGGCN_W_GEO_MLP2_B, TGGCN_W_HUM_EMB_W, TGGCN_W_HUM_EMB_B, TGGCN_W_OBJ_EMB_W, TGGCN_W_OBJ_EMB_B,
            TGGCN_W_GEO_RNN_WIH_F, TGGCN_W_GEO_RNN_WHH_F, TGGCN_W_GEO_RNN_BIH_F, TGGCN_W_GEO_RNN_BHH_F, TGGCN_W_GEO_RNN_WIH_B,
            TGGCN_W_GEO_RNN_WHH_B, TGGCN_W_GEO_RNN_BIH_B, TGGCN_W_GEO_RNN_BHH_B, TGGCN_W_HUM_RNN_WIH_F, TGGCN_W_HUM_RNN_WHH_F,
            TGGCN_W_HUM_RNN_BIH_F, TGGCN_W_HUM_RNN_BHH_F, TGGCN_W_HUM_RNN_WIH_B, TGGCN_W_HUM_RNN_WHH_B, TGGCN_W_HUM_RNN_BIH_B,
            TGGCN_W_HUM_RNN_BHH_B, TGGCN_W_OBJ_RNN_WIH_F, TGGCN_W_OBJ_RNN_WHH_F, TGGCN_W_OBJ_RNN_BIH_F, TGGCN_W_OBJ_RNN_BHH_F,
            TGGCN_W_OBJ_RNN_WIH_B, TGGCN_W_OBJ_RNN_WHH_B, TGGCN_W_OBJ_RNN_BIH_B, TGGCN_W_OBJ_RNN_BHH_B, TGGCN_W_GEO_BD_W,
            TGGCN_W_GEO_BD_B, TGGCN_W_HUM_BD_W, TGGCN_W_HUM_BD_B, TGGCN_W_OBJ_BD_W, TGGCN_W_OBJ_BD_B, TGGCN_W_MSG_HO_W,
            TGGCN_W_MSG_HO_B, TGGCN_W_MSG_OH_W, TGGCN_W_MSG_OH_B, TGGCN_W_MSG_OO_W, TGGCN_W_MSG_OO_B, TGGCN_W_MSG_GO_W,
            TGGCN_W_MSG_GO_B, TGGCN_W_SMSG_HO_W, TGGCN_W_SMSG_HO_B, TGGCN_W_SMSG_OH_W, TGGCN_W_SMSG_OH_B, TGGCN_W_SMSG_OO_W,
            TGGCN_W_SMSG_OO_B, TGGCN_W_HSEG_F_WIH, TGGCN_W_HSEG_F_WHH, TGGCN_W_HSEG_F_BIH, TGGCN_W_HSEG_F_BHH, TGGCN_W_HSEG_B_WIH,
            TGGCN_W_HSEG_B_WHH, TGGCN_W_HSEG_B_BIH, TGGCN_W_HSEG_B_BHH, TGGCN_W_OSEG_F_WIH, TGGCN_W_OSEG_F_WHH, TGGCN_W_OSEG_F_BIH,
            TGGCN_W_OSEG_F_BHH, TGGCN_W_OSEG_B_WIH, TGGCN_W_OSEG_B_WHH, TGGCN_W_OSEG_B_BIH, TGGCN_W_OSEG_B_BHH,
            TGGCN_W_HEAD_H_FREC_W, TGGCN_W_HEAD_H_FREC_B, TGGCN_W_HEAD_H_FPRED_W, TGGCN_W_HEAD_H_FPRED_B, TGGCN_W_HEAD_H_REC_W,
            TGGCN_W_HEAD_H_REC_B, TGGCN_W_HEAD_H_PRED_W, TGGCN_W_HEAD_H_PRED_B};
        for (size_t i = 0; i < sizeof(need) / sizeof(need[0]); ++i)
            TG_REQUIRE(weights[need[i]] && grad_weights[need[i]], "backward: weight / gradient pointer #%d is null", need[i]);
        if (d.hh)
            TG_REQUIRE(G(TGGCN_W_MSG_HH_W) && G(TGGCN_W_MSG_HH_B) && G(TGGCN_W_SMSG_HH_W) && G(TGGCN_W_SMSG_HH_B),
                       "backward: humans->human gradient pointers missing");
        if (d.C_aff > 0)
            for (int id = TGGCN_W_HEAD_O_FREC_W; id <= TGGCN_W_HEAD_O_PRED_B; ++id)
                TG_REQUIRE(weights[id] && grad_weights[id], "backward: object head pointer #%d is null", id);
        if (!d.human_seg_given) TG_REQUIRE(G(TGGCN_W_UPD_H_W) && G(TGGCN_W_UPD_H_B), "backward: human gate gradient pointers missing");
        if (!d.object_seg_given && d.update_strategy != 1)
            TG_REQUIRE(G(TGGCN_W_UPD_O_W) && G(TGGCN_W_UPD_O_B), "backward: object gate gradient pointers missing");
    }

    // weight gradient dW[N,K] (+)= (Z (.) [mask > 0])^T X with an optional row shift of X inside blocks of `period` rows:
    // both operands are transposed into scratch (reduction index contiguous) and contracted by the tcgen05 NT kernel.
    float* tn_scratch = bb(BL.tn);
    G16Ctx g16;
    g16.ws = nullptr; g16.bytes = BL.tn_floats * sizeof(float); g16.precision = d.precision == 1 ? 1 : 0;
    g16.err = (unsigned int*)buf(TGGCN_BUF_SYNC) + 7;
    if (d.gemm_path != 0 && gemm16_enabled() && (d.precision == 1 || !d.no_fp16_split)) g16.ws = tn_scratch;
    // Input gradient of a projection: C[M, kf] (+)= (A (.) [mask > 0])[M, sum nf_i] * [W_0 ; W_1 ; ...] with W_i the FORWARD weight
    // (nf_i rows of kf used columns, row stride ldw) — e.g. the two time directions of a hoisted W_ih.  On the TMA kernel the weights
    // are packed transposed straight from their row-major storage (no fp32 transpose pass), the gradient block of each source is one
    // problem, and the problems add into C atomically; otherwise: fp32 transposes into `wt` and the tf32 / bf16 kernel.
    struct WSrc { const float* W; int ldw; int nf; };
    auto dx_gemm = [&](const float* A, int lda, const float* mask, int ldm, const WSrc* ws, int nws, int kf, float* C, int ldc, int M,
                       int beta, float* wt) -> int {
        int ktot = 0;
        for (int i = 0; i < nws; ++i) ktot += ws[i].nf;
        if (g16.ws != nullptr && nws <= 2) {
            Gemm16Problem q[2];
            memset(q, 0, sizeof(q));
            int koff = 0;
            for (int i = 0; i < nws; ++i) {
                q[i].a.src = A + koff; q[i].a.ld = lda; q[i].a.rows = M; q[i].a.cols = ws[i].nf;
                q[i].a.mask = mask != nullptr ? mask + koff : nullptr; q[i].a.ldm = ldm; q[i].a.dynamic = 1;
                q[i].b.src = ws[i].W; q[i].b.ld = ws[i].ldw; q[i].b.rows = ws[i].nf; q[i].b.cols = kf; q[i].b.transpose = 1; q[i].b.scale = 256.0f;
                q[i].C = C; q[i].ldc = ldc; q[i].beta = i == 0 ? beta : 1;
                koff += ws[i].nf;
            }
            if (gemm16_eligible(q, nws) && gemm16_scratch_bytes(q, nws) <= g16.bytes)
                return launch_gemm16(q, nws, g16.precision, g16.ws, g16.bytes, g16.err, stream);
        }
        int koff = 0;
        for (int i = 0; i < nws; ++i) {                      // wt[kf][ktot] = [W_0^T | W_1^T | ...]
            if (int rc = launch_transpose(ws[i].W, ws[i].ldw, wt + koff, ktot, ws[i].nf, kf, stream)) return rc;
            koff += ws[i].nf;
        }
        return gemm_nt(A, lda, mask, ldm, wt, ktot, C, ldc, M, kf, ktot, beta, path, stream, G16Ctx{nullptr, 0, 0, nullptr});
    };

    // One weight gradient dW[Nn,K] (+)= (Z (.) [mask > 0])^T X (optional row shift of X inside blocks of `period` rows) and, when db is
    // given, the bias gradient db[Nn] (+)= column sums of the masked Z.  Calls are COLLECTED per backward stage and flushed as one
    // grouped launch of the TMA-fed kernel (gemm16.cu): every distinct operand is packed once (transposed: the row index becomes the
    // contiguous reduction index), the column sums ride on the pack of Z, and the small problems of a stage fill the GPU together.
    struct TnCall {
        const float* Z; int ldz; const float* mask; int ldm; const float* X; int ldx; float* dW; int lddw; int M, Nn, K, shift, period, beta;
        float* db; int db_beta;
    };
    TnCall tn_calls[GEMM_MAX_PROBLEMS];
    int tn_count = 0;
    auto tn_single = [&](const TnCall& c, cudaStream_t st) -> int {      // fp32 transposes + the tf32 / bf16 kernel of gemm_tc.cu
        const int Mp = (c.M + 31) / 32 * 32;
        TG_REQUIRE((size_t)(c.Nn + c.K) * Mp <= BL.tn_floats, "backward: weight-gradient scratch too small (%d + %d) x %d", c.Nn, c.K, Mp);
        float* zt = tn_scratch;
        float* xt = zt + (size_t)c.Nn * Mp;
        if (int rc = launch_transpose_prep(c.Z, c.ldz, c.mask, c.ldm, zt, Mp, c.M, c.Nn, 0, 0, nullptr, st)) return rc;
        if (int rc = launch_transpose_prep(c.X, c.ldx, nullptr, 0, xt, Mp, c.M, c.K, c.shift, c.period, nullptr, st)) return rc;
        if (int rc = gemm_nt(zt, Mp, nullptr, 0, xt, Mp, c.dW, c.lddw, c.Nn, c.K, Mp, c.beta, path, st, G16Ctx{nullptr, 0, 0, nullptr})) return rc;
        if (c.db != nullptr) return launch_colsum(c.Z, c.ldz, c.mask, c.ldm, c.db, c.M, c.Nn, c.db_beta, st);
        return 0;
    };
    auto tn_flush = [&](cudaStream_t st) -> int {
        if (tn_count == 0) return 0;
        const int n = tn_count;
        tn_count = 0;
        if (g16.ws != nullptr) {
            Gemm16Problem q[GEMM_MAX_PROBLEMS];
            memset(q, 0, sizeof(q));
            for (int i = 0; i < n; ++i) {
                const TnCall& c = tn_calls[i];
                q[i].a.src = c.Z; q[i].a.ld = c.ldz; q[i].a.rows = c.M; q[i].a.cols = c.Nn; q[i].a.mask = c.mask; q[i].a.ldm = c.ldm;
                q[i].a.transpose = 1; q[i].a.dynamic = 1; q[i].a.colsum = c.db; q[i].a.colsum_beta = c.db_beta;
                q[i].b.src = c.X; q[i].b.ld = c.ldx; q[i].b.rows = c.M; q[i].b.cols = c.K; q[i].b.transpose = 1;
                q[i].b.shift = c.shift; q[i].b.period = c.period; q[i].b.scale = 1.0f;
                q[i].C = c.dW; q[i].ldc = c.lddw; q[i].beta = c.beta;
            }
            if (gemm16_eligible(q, n) && gemm16_scratch_bytes(q, n) <= g16.bytes)
                return launch_gemm16(q, n, g16.precision, g16.ws, g16.bytes, g16.err, st);
            for (int i = 0; i < n; ++i) {                                  // the group does not fit the scratch: one problem at a time
                if (gemm16_eligible(&q[i], 1) && gemm16_scratch_bytes(&q[i], 1) <= g16.bytes) {
                    if (int rc = launch_gemm16(&q[i], 1, g16.precision, g16.ws, g16.bytes, g16.err, st)) return rc;
                } else if (int rc = tn_single(tn_calls[i], st)) return rc;
            }
            return 0;
        }
        for (int i = 0; i < n; ++i)
            if (int rc = tn_single(tn_calls[i], st)) return rc;
        return 0;
    };
    auto tn = [&](const float* Z, int ldz, const float* mask, int ldm, const float* X, int ldx, float* dW, int lddw, int M, int Nn, int K,
                  int shift, int period, int beta, cudaStream_t st, float* db = nullptr, int db_beta = 0) -> int {
        if (tn_count == GEMM_MAX_PROBLEMS)
            if (int rc = tn_flush(st)) return rc;
        tn_calls[tn_count++] = TnCall{Z, ldz, mask, ldm, X, ldx, dW, lddw, M, Nn, K, shift, period, beta, db, db_beta};
        return 0;
    };

    TG_CUDA_OK(cudaMemsetAsync(bb(BL.zero_begin), 0, (BL.zero_end - BL.zero_begin) * sizeof(float), stream));
    TG_CUDA_OK(cudaMemsetAsync((unsigned int*)buf(TGGCN_BUF_SYNC) + 4, 0, 4 * sizeof(unsigned int), stream));      // status words of this call

    // ---- 12. heads --------------------------------------------------------------------------------------------------
    {
        HeadsBwdParams P;
        memset(&P, 0, sizeof(P));
        P.B = B; P.T = T; P.E = H; P.NE = H + O; P.e_off = 0; P.D = D; P.C = d.C_sub; P.cat = d.cat_level_states;
        P.hfr = buf(TGGCN_BUF_HFR_H); P.hx = buf(TGGCN_BUF_HX_H); P.reidx = (const int*)buf(TGGCN_BUF_REIDX);
        const int wid[4] = {TGGCN_W_HEAD_H_FREC_W, TGGCN_W_HEAD_H_FPRED_W, TGGCN_W_HEAD_H_REC_W, TGGCN_W_HEAD_H_PRED_W};
        // (share_level_mlps: the frame-level and segment-level heads are the same tensors; their gradients then accumulate
        //  into the same destination, which is zeroed before the launch like any other)
        for (int i = 0; i < 4; ++i) {
            P.w[i] = W(wid[i]); P.bias[i] = W(wid[i] + 1); P.dlogp[i] = grads->d_out_h[i];
            P.dw[i] = G(wid[i]); P.db[i] = G(wid[i] + 1);
            TG_CUDA_OK(cudaMemsetAsync(P.dw[i], 0, sizeof(float) * (size_t)d.C_sub * ((i >= 2 && d.cat_level_states) ? 4 : 2) * D, stream));
            TG_CUDA_OK(cudaMemsetAsync(P.db[i], 0, sizeof(float) * (size_t)d.C_sub, stream));
        }
        P.dhfr = bb(BL.dhfr[0]); P.dhx = bb(BL.dhx[0]);
        if (int rc = launch_heads_bwd(P, stream)) return rc;
        if (d.C_aff > 0) {
            P.E = O; P.e_off = H; P.C = d.C_aff;
            P.hfr = buf(TGGCN_BUF_HFR_O); P.hx = buf(TGGCN_BUF_HX_O);
            const int oid[4] = {TGGCN_W_HEAD_O_FREC_W, TGGCN_W_HEAD_O_FPRED_W, TGGCN_W_HEAD_O_REC_W, TGGCN_W_HEAD_O_PRED_W};
            for (int i = 0; i < 4; ++i) {
                P.w[i] = W(oid[i]); P.bias[i] = W(oid[i] + 1); P.dlogp[i] = grads->d_out_o[i];
                P.dw[i] = G(oid[i]); P.db[i] = G(oid[i] + 1);
                TG_CUDA_OK(cudaMemsetAsync(P.dw[i], 0, sizeof(float) * (size_t)d.C_aff * ((i >= 2 && d.cat_level_states) ? 4 : 2) * D, stream));
                TG_CUDA_OK(cudaMemsetAsync(P.db[i], 0, sizeof(float) * (size_t)d.C_aff, stream));
            }
            P.dhfr = bb(BL.dhfr[1]); P.dhx = bb(BL.dhx[1]);
            if (int rc = launch_heads_bwd(P, stream)) return rc;
        }
    }

    // ---- 11. segment-level recurrent graph, reverse time ---------------------------------------------------------------
    const int kh = kh_of(d), ldwh = ldwh_of(d);                  // human cell: frame-part columns, row stride of W_ih
    const int ko = ko_of(d), ldwo = ldwo_of(d);                  // object cell
    const int wih_h_id[2] = {TGGCN_W_HSEG_F_WIH, TGGCN_W_HSEG_B_WIH}, wih_o_id[2] = {TGGCN_W_OSEG_F_WIH, TGGCN_W_OSEG_B_WIH};
    const int whh_h_id[2] = {TGGCN_W_HSEG_F_WHH, TGGCN_W_HSEG_B_WHH}, whh_o_id[2] = {TGGCN_W_OSEG_F_WHH, TGGCN_W_OSEG_B_WHH};
    const int smsg_w_id[4] = {TGGCN_W_SMSG_HH_W, TGGCN_W_SMSG_OH_W, TGGCN_W_SMSG_HO_W, TGGCN_W_SMSG_OO_W};
    {
        // transposed weights, kept for the whole loop
        float* p = bb(BL.wt_seg);
        float* whhT_h[2]; float* whhT_o[2]; float* wihT_h[2]; float* wihT_o[2];
        for (int dir = 0; dir < 2; ++dir) { whhT_h[dir] = p; p += (size_t)D * 3 * D; }
        for (int dir = 0; dir < 2; ++dir) { whhT_o[dir] = p; p += (size_t)D * 3 * D; }
        for (int dir = 0; dir < 2; ++dir) { wihT_h[dir] = p; p += (size_t)nkh * D * 3 * D; }
        for (int dir = 0; dir < 2; ++dir) { wihT_o[dir] = p; p += (size_t)2 * D * 3 * D; }
        float* wmT_h = p; p += (size_t)D * nks * D;      // (D, nks*D): [W_hh_msg^T | W_ho_msg^T]
        float* wmT_o = p;                                // (D, 2D):    [W_oh_msg^T | W_oo_msg^T]
        for (int dir = 0; dir < 2; ++dir) {
            if (int rc = launch_transpose(W(whh_h_id[dir]), D, whhT_h[dir], 3 * D, 3 * D, D, stream)) return rc;
            if (int rc = launch_transpose(W(whh_o_id[dir]), D, whhT_o[dir], 3 * D, 3 * D, D, stream)) return rc;
            if (int rc = launch_transpose(W(wih_h_id[dir]) + kh, ldwh, wihT_h[dir], 3 * D, 3 * D, nkh * D, stream)) return rc;
            if (int rc = launch_transpose(W(wih_o_id[dir]) + ko, ldwo, wihT_o[dir], 3 * D, 3 * D, 2 * D, stream)) return rc;
        }
        if (d.hh)
            if (int rc = launch_transpose(W(TGGCN_W_SMSG_HH_W), D, wmT_h, nks * D, D, D, stream)) return rc;
        if (int rc = launch_transpose(W(TGGCN_W_SMSG_HO_W), D, wmT_h + (nks - 1) * D, nks * D, D, D, stream)) return rc;
        if (int rc = launch_transpose(W(TGGCN_W_SMSG_OH_W), D, wmT_o, 2 * D, D, D, stream)) return rc;
        if (int rc = launch_transpose(W(TGGCN_W_SMSG_OO_W), D, wmT_o + D, 2 * D, D, D, stream)) return rc;

        SegBwdParams P;
        memset(&P, 0, sizeof(P));
        P.B = B; P.T = T; P.H = H; P.O = O; P.D = D; P.hh = d.hh; P.nk_h = nkh; P.mean_pool = d.mean_pool; P.att_noscale = d.att_noscale;
        P.dist_kind[0] = (d.hh && io->dist_hh && !d.mean_pool) ? 1 : 0; P.dist_kind[1] = P.dist_kind[2] = (io->dist_ho && !d.mean_pool) ? 1 : 0;
        P.dist_kind[3] = (io->dist_oo && !d.mean_pool) ? 1 : 0;
        P.hx_h = buf(TGGCN_BUF_HX_H); P.hx_o = buf(TGGCN_BUF_HX_O);
        P.sgates_h = buf(TGGCN_BUF_SGATES_H); P.sgates_o = buf(TGGCN_BUF_SGATES_O);
        P.u_h = io->y_hs; P.u_o = io->y_os; P.om = io->objects_mask;
        P.dhx_h = bb(BL.dhx[0]); P.dhx_o = bb(BL.dhx[1]);
        P.dgs_h = bb(BL.dgs[0]); P.dgs_o = bb(BL.dgs[1]);
        P.dghs_h = bb(BL.dghs[0]); P.dghs_o = bb(BL.dghs[1]);
        P.du_h = bb(BL.du[0]); P.du_o = bb(BL.du[1]);
        P.direct_h = bb(BL.direct[0]); P.direct_o = bb(BL.direct[1]);
        P.dmg_h = bb(BL.dmg[0]); P.dmg_o = bb(BL.dmg[1]);
        P.dpre_h = bb(BL.dpre[0]); P.dpre_o = bb(BL.dpre[1]);
        for (int dir = 0; dir < 2; ++dir) {
            P.wihT_h[dir] = wihT_h[dir]; P.wihT_o[dir] = wihT_o[dir]; P.whhT_h[dir] = whhT_h[dir]; P.whhT_o[dir] = whhT_o[dir];
        }
        P.wmT_h = wmT_h; P.wmT_o = wmT_o;
        const int smsg_id[4] = {TGGCN_BUF_SMSG_HH, TGGCN_BUF_SMSG_OH, TGGCN_BUF_SMSG_HO, TGGCN_BUF_SMSG_OO};
        const int salpha_id[4] = {TGGCN_BUF_SALPHA_HH, TGGCN_BUF_SALPHA_OH, TGGCN_BUF_SALPHA_HO, TGGCN_BUF_SALPHA_OO};
        for (int k = 0; k < 4; ++k) {
            P.smsg[k] = buf(smsg_id[k]); P.salpha[k] = buf(salpha_id[k]); P.dpre_all[k] = bb(BL.dpre_all[k]);
            P.lgr[k] = bb(BL.lgr[k]); P.lgs[k] = bb(BL.lgs[k]);
        }
        unsigned int* sync = (unsigned int*)buf(TGGCN_BUF_SYNC);
        P.sync.counter = sync + 6; P.sync.error = sync + 7;
        if (int rc = launch_segment_bwd(P, d.persistent, stream)) return rc;
        // weight gradients of the cells and message MLPs: GEMMs over all (video, t, entity) rows
        for (int dir = 0; dir < 2; ++dir) {
            // humans
            if (int rc = tn(P.dghs_h + (size_t)dir * 3 * D, 6 * D, nullptr, 0, P.hx_h + (size_t)dir * D, 2 * D,
                                        G(whh_h_id[dir]), D, N * H, 3 * D, D, dir == 0 ? -H : H, T * H, 0, stream, G(whh_h_id[dir] + 2))) return rc;
            if (int rc = tn(P.dgs_h + (size_t)dir * 3 * D, 6 * D, nullptr, 0, buf(TGGCN_BUF_XX_H), kh, G(wih_h_id[dir]), ldwh,
                                        N * H, 3 * D, kh, 0, 0, 0, stream, G(wih_h_id[dir] + 2))) return rc;
            if (int rc = tn(P.dgs_h + (size_t)dir * 3 * D, 6 * D, nullptr, 0,
                                        buf(TGGCN_BUF_MG_ALL_H) + (size_t)dir * N * H * nkh * D, nkh * D, G(wih_h_id[dir]) + kh, ldwh,
                                        N * H, 3 * D, nkh * D, 0, 0, 0, stream)) return rc;
            // objects
            if (int rc = tn(P.dghs_o + (size_t)dir * 3 * D, 6 * D, nullptr, 0, P.hx_o + (size_t)dir * D, 2 * D,
                                        G(whh_o_id[dir]), D, N * O, 3 * D, D, dir == 0 ? -O : O, T * O, 0, stream, G(whh_o_id[dir] + 2))) return rc;
            if (int rc = tn(P.dgs_o + (size_t)dir * 3 * D, 6 * D, nullptr, 0, buf(TGGCN_BUF_XX_O), ko, G(wih_o_id[dir]), ldwo,
                                        N * O, 3 * D, ko, 0, 0, 0, stream, G(wih_o_id[dir] + 2))) return rc;
            if (int rc = tn(P.dgs_o + (size_t)dir * 3 * D, 6 * D, nullptr, 0,
                                        buf(TGGCN_BUF_MG_ALL_O) + (size_t)dir * N * O * 2 * D, 2 * D, G(wih_o_id[dir]) + ko, ldwo,
                                        N * O, 3 * D, 2 * D, 0, 0, 0, stream)) return rc;
            if (int rc = tn_flush(stream)) return rc;
        }
        for (int k = d.hh ? 0 : 1; k < 4; ++k) {
            const bool send_h = (k == 0 || k == 2);
            const int Es = send_h ? H : O;
            const float* hx = send_h ? P.hx_h : P.hx_o;
            for (int dir = 0; dir < 2; ++dir)          // both directions accumulate into the same weight and bias gradient
                if (int rc = tn(P.dpre_all[k] + (size_t)dir * N * Es * D, D, nullptr, 0, hx + (size_t)dir * D, 2 * D,
                                            G(smsg_w_id[k]), D, N * Es, D, D, dir == 0 ? -Es : Es, T * Es, dir, stream, G(smsg_w_id[k] + 1), dir)) return rc;
        }
        if (int rc = tn_flush(stream)) return rc;
    }

    if (hooks && hooks->bucket_done[0]) TG_CUDA_OK(cudaEventRecord((cudaEvent_t)hooks->bucket_done[0], stream));

    // ---- 10. hoisted frame-part of the segment cells: d xx = [dGs_f | dGs_b] [W_ih_f[:, :k] ; W_ih_b[:, :k]] -----------------------
    {
        float* wt = bb(BL.wt);
        const WSrc wh[2] = {{W(wih_h_id[0]), ldwh, 3 * D}, {W(wih_h_id[1]), ldwh, 3 * D}};
        if (int rc = dx_gemm(bb(BL.dgs[0]), 6 * D, nullptr, 0, wh, 2, kh, bb(BL.dxx[0]), kh, N * H, 0, wt)) return rc;
        const WSrc wo[2] = {{W(wih_o_id[0]), ldwo, 3 * D}, {W(wih_o_id[1]), ldwo, 3 * D}};
        if (int rc = dx_gemm(bb(BL.dgs[1]), 6 * D, nullptr, 0, wo, 2, ko, bb(BL.dxx[1]), ko, N * O, 0, wt)) return rc;
    }

    // ---- 9b. segment lengths: gradient of the hard gates through them, segment_length_mlp ------------------------------------------
    if (d.segment_length) {
        SegLenBwdParams P;
        memset(&P, 0, sizeof(P));
        P.B = B; P.T = T; P.H = H; P.O = O; P.D = D; P.periodic = d.time_periodic;
        P.dxx_h = bb(BL.dxx[0]); P.xx_h = buf(TGGCN_BUF_XX_H); P.ldh = kh_of(d);
        P.dxx_o = bb(BL.dxx[1]); P.xx_o = buf(TGGCN_BUF_XX_O); P.ldo = ko_of(d);
        P.len = buf(TGGCN_BUF_SEG_LEN); P.y_hs = io->y_hs; P.y_os = io->y_os;
        P.steps = io->steps_per_example; P.w = W(TGGCN_W_LEN_W); P.freq = io->time_freq;
        P.du_h = bb(BL.du[0]); P.du_o = bb(BL.du[1]);
        if (!d.time_periodic) { P.dw = G(TGGCN_W_LEN_W); P.db = G(TGGCN_W_LEN_B); }
        TG_REQUIRE(P.steps != nullptr && (d.time_periodic ? P.freq != nullptr : P.w != nullptr), "backward: segment-length inputs missing");
        if (int rc = launch_segment_length_bwd(P, stream)) return rc;
    }
    // ---- 9/8. gates (straight-through, filter), attention, aggregation ------------------------------------------------------------
    {
        FrameBwdParams P;
        memset(&P, 0, sizeof(P));
        P.B = B; P.T = T; P.H = H; P.O = O; P.D = D; P.hh = d.hh; P.filter = d.filter; P.thr = d.thr; P.mean_pool = d.mean_pool; P.att_noscale = d.att_noscale;
        P.update_strategy = d.update_strategy; P.straight_through = d.straight_through;
        P.gh = d.geo_to_human ? 1 : 0; P.msg_gh = P.gh ? buf(TGGCN_BUF_MSG_GH) : nullptr; P.dmsg_gh = P.gh ? bb(BL.dmsg[5]) : nullptr;
        P.tl = tl_of(d);
        P.dist_kind[0] = (d.hh && io->dist_hh && !d.mean_pool) ? 1 : 0; P.dist_kind[1] = P.dist_kind[2] = (io->dist_ho && !d.mean_pool) ? 1 : 0;
        P.dist_kind[3] = (io->dist_oo && !d.mean_pool) ? 1 : 0;
        P.time_position = d.time_position; P.time_emb = d.time_position ? buf(TGGCN_BUF_TIME_EMB) : nullptr;
        // no gradient pointers for time_position_mlp = it is off the gradient path of this call (strategy 'u' with every gate imposed)
        P.dtime = (d.time_position && !d.time_periodic && G(TGGCN_W_TIME_W) && G(TGGCN_W_TIME_B)) ? bb(BL.dtime) : nullptr;
        P.s_h = buf(TGGCN_BUF_S_H); P.s_o = buf(TGGCN_BUF_S_O);
        P.msg_hh = buf(TGGCN_BUF_MSG_HH); P.msg_ho = buf(TGGCN_BUF_MSG_HO); P.msg_oh = buf(TGGCN_BUF_MSG_OH);
        P.msg_oo = buf(TGGCN_BUF_MSG_OO); P.msg_go = buf(TGGCN_BUF_MSG_GO);
        P.om = io->objects_mask;
        P.w_uh = W(TGGCN_W_UPD_H_W); P.w_uo = W(TGGCN_W_UPD_O_W);
        P.alpha = buf(TGGCN_BUF_ALPHA_F); P.pgate = buf(TGGCN_BUF_PGATE);
        P.y_hss = io->y_hss; P.y_oss = io->y_oss;
        P.human_seg = io->human_seg; P.object_seg = io->object_seg;
        P.xx_h = buf(TGGCN_BUF_XX_H); P.xx_o = buf(TGGCN_BUF_XX_O);
        P.dxx_h = bb(BL.dxx[0]); P.dxx_o = bb(BL.dxx[1]);
        P.du_h = bb(BL.du[0]); P.du_o = bb(BL.du[1]);
        P.dy_hs = grads->d_y_hs; P.dy_os = grads->d_y_os; P.dy_hss = grads->d_y_hss; P.dy_oss = grads->d_y_oss;
        P.ds_h = bb(BL.ds[0]); P.ds_o = bb(BL.ds[1]);
        P.dmsg_hh = bb(BL.dmsg[0]); P.dmsg_ho = bb(BL.dmsg[1]); P.dmsg_oh = bb(BL.dmsg[2]); P.dmsg_oo = bb(BL.dmsg[3]);
        P.dmsg_go = bb(BL.dmsg[4]);
        P.dw_uh = G(TGGCN_W_UPD_H_W); P.db_uh = G(TGGCN_W_UPD_H_B); P.dw_uo = G(TGGCN_W_UPD_O_W); P.db_uo = G(TGGCN_W_UPD_O_B);
        const bool sample_h = !d.human_seg_given, sample_o = !d.object_seg_given && d.update_strategy != 1;
        if (gate2_of(d)) {
            // gate MLPs with hidden layers: dlogit -> d (last hidden) (+ the last layer's weight gradients), then per hidden layer
            // d input = d hidden . W and dW = d hidden^T input, down to the gate inputs
            P.gate_layers = 2;
            P.gin_h = ginh_of(d); P.gin_o = gino_of(d);
            const bool three = gate3_of(d);
            const int last_id[2] = {three ? TGGCN_W_UPD_H_W4 : TGGCN_W_UPD_H_W2, three ? TGGCN_W_UPD_O_W4 : TGGCN_W_UPD_O_W2};
            const int hid1_buf[2] = {TGGCN_BUF_GATE_HID_H, TGGCN_BUF_GATE_HID_O}, hid2_buf[2] = {TGGCN_BUF_GATE_HID2_H, TGGCN_BUF_GATE_HID2_O};
            const int gin_buf[2] = {TGGCN_BUF_GATE_IN_H, TGGCN_BUF_GATE_IN_O};
            const int w1_id[2] = {TGGCN_W_UPD_H_W, TGGCN_W_UPD_O_W}, w2_id[2] = {TGGCN_W_UPD_H_W2, TGGCN_W_UPD_O_W2};
            const bool sample[2] = {sample_h, sample_o};
            const int Ee[2] = {H, O}, gin[2] = {P.gin_h, P.gin_o};
            P.hid_h = buf(three ? hid2_buf[0] : hid1_buf[0]); P.hid_o = buf(three ? hid2_buf[1] : hid1_buf[1]);
            P.w2_h = W(last_id[0]); P.w2_o = W(last_id[1]);
            for (int k = 0; k < 2; ++k) {
                if (!sample[k]) continue;
                TG_REQUIRE(W(last_id[k]) && G(last_id[k]) && G(last_id[k] + 1) && W(w1_id[k]) && G(w1_id[k]) && G(w1_id[k] + 1),
                           "backward: gate MLP pointers of entity type %d missing", k);
                if (three) TG_REQUIRE(W(w2_id[k]) && G(w2_id[k]) && G(w2_id[k] + 1), "backward: middle gate layer pointers of entity type %d missing", k);
                float* dh = bb(BL.dghid[k]);
                float* dwl = G(last_id[k]);
                float* dbl = G(last_id[k] + 1);
                if (k == 0) { P.dhid_h = dh; P.dw2_h = dwl; P.db2_h = dbl; } else { P.dhid_o = dh; P.dw2_o = dwl; P.db2_o = dbl; }
                TG_CUDA_OK(cudaMemsetAsync(dwl, 0, sizeof(float) * (size_t)D, stream));
                TG_CUDA_OK(cudaMemsetAsync(dbl, 0, sizeof(float), stream));
            }
            if (sample_h || sample_o) {
                if (int rc = launch_gate_bwd(P, stream)) return rc;
                float* wt = bb(BL.wt);
                for (int k = 0; k < 2; ++k) {
                    if (!sample[k]) continue;
                    const int M = N * Ee[k];
                    const float* dz = bb(BL.dghid[k]);           // gradient of the last hidden layer, ReLU mask applied by gate_bwd_kernel
                    const float* zmask = nullptr;
                    if (three) {                                 // middle layer: hid2 = ReLU(W2 hid1 + b2)
                        const WSrc w2[1] = {{W(w2_id[k]), D, D}};
                        if (int rc = dx_gemm(dz, D, nullptr, 0, w2, 1, D, bb(BL.dghid1[k]), D, M, 0, wt)) return rc;
                        if (int rc = tn(dz, D, nullptr, 0, buf(hid1_buf[k]), D, G(w2_id[k]), D, M, D, D, 0, 0, 0, stream, G(w2_id[k] + 1))) return rc;
                        dz = bb(BL.dghid1[k]);                   // gradient of hid1 BEFORE its ReLU mask: the mask rides on the operand packs
                        zmask = buf(hid1_buf[k]);
                    }
                    const WSrc w1[1] = {{W(w1_id[k]), gin[k], D}};
                    if (int rc = dx_gemm(dz, D, zmask, D, w1, 1, gin[k], bb(BL.dgin[k]), gin[k], M, 0, wt)) return rc;
                    if (int rc = tn(dz, D, zmask, D, buf(gin_buf[k]), gin[k], G(w1_id[k]), gin[k], M, D, gin[k], 0, 0, 0, stream, G(w1_id[k] + 1))) return rc;
                    if (k == 0) P.dgin_h = bb(BL.dgin[0]); else P.dgin_o = bb(BL.dgin[1]);
                }
                if (int rc = tn_flush(stream)) return rc;
            }
        } else {
            if (sample_h) {
                TG_CUDA_OK(cudaMemsetAsync(P.dw_uh, 0, sizeof(float) * (size_t)(2 + nkh + gh_of(d) + tu_of(d)) * D, stream));
                TG_CUDA_OK(cudaMemsetAsync(P.db_uh, 0, sizeof(float), stream));
            }
            if (sample_o) {
                TG_CUDA_OK(cudaMemsetAsync(P.dw_uo, 0, sizeof(float) * (size_t)(5 + tu_of(d)) * D, stream));
                TG_CUDA_OK(cudaMemsetAsync(P.db_uo, 0, sizeof(float), stream));
            }
        }
        if (int rc = launch_frame_bwd(P, stream)) return rc;
        if (P.dtime != nullptr) {      // time_position_mlp: Linear(1, D) + ReLU of (t+1) / steps
            TG_REQUIRE(io->steps_per_example, "backward: steps_per_example missing");
            if (int rc = launch_time_embed_bwd(P.dtime, P.time_emb, io->steps_per_example, G(TGGCN_W_TIME_W), G(TGGCN_W_TIME_B), B, T, D, stream)) return rc;
        }
    }

    // ---- 7. message MLPs: msg = ReLU(W [x|h] + b) ---------------------------------------------------------------------------------
    {
        struct Kind { int w_id; int msg_buf; int dmsg; int grp; int on; };
        const Kind kinds[6] = {{TGGCN_W_MSG_HH_W, TGGCN_BUF_MSG_HH, 0, 0, d.hh}, {TGGCN_W_MSG_HO_W, TGGCN_BUF_MSG_HO, 1, 0, 1},
                               {TGGCN_W_MSG_OH_W, TGGCN_BUF_MSG_OH, 2, 1, 1},    {TGGCN_W_MSG_OO_W, TGGCN_BUF_MSG_OO, 3, 1, 1},
                               {TGGCN_W_MSG_GO_W, TGGCN_BUF_MSG_GO, 4, 2, 1},    {TGGCN_W_MSG_GH_W, TGGCN_BUF_MSG_GH, 5, 2, d.geo_to_human}};
        const int s_buf[3] = {TGGCN_BUF_S_H, TGGCN_BUF_S_O, TGGCN_BUF_S_G};
        const int Eg[3] = {H, O, 1};
        int touched[3] = {1, 1, 0};      // ds_h / ds_o were written by the frame kernel; ds_g starts here
        float* wt = bb(BL.wt);
        for (int k = 0; k < 6; ++k) {
            if (!kinds[k].on) continue;
            TG_REQUIRE(W(kinds[k].w_id) && G(kinds[k].w_id) && G(kinds[k].w_id + 1), "backward: message MLP #%d pointers missing", k);
            const int gidx = kinds[k].grp, M = N * Eg[gidx];
            const float* dmsg = bb(BL.dmsg[kinds[k].dmsg]);
            const float* msg = buf(kinds[k].msg_buf);
            const WSrc wk[1] = {{W(kinds[k].w_id), 2 * D, D}};
            if (int rc = dx_gemm(dmsg, D, msg, D, wk, 1, 2 * D, bb(BL.ds[gidx]), 2 * D, M, touched[gidx], wt)) return rc;
            touched[gidx] = 1;
            if (int rc = tn(dmsg, D, msg, D, buf(s_buf[gidx]), 2 * D, G(kinds[k].w_id), 2 * D, M, D, 2 * D, 0, 0, 0, stream, G(kinds[k].w_id + 1))) return rc;
        }
        if (int rc = tn_flush(stream)) return rc;
    }

    // ---- 6. Linear(2D->D)+ReLU on the BiGRU outputs: h = S[:, D:2D] ------------------------------------------------------------------
    const int s_buf[3] = {TGGCN_BUF_S_H, TGGCN_BUF_S_O, TGGCN_BUF_S_G};
    const int hfr_buf[3] = {TGGCN_BUF_HFR_H, TGGCN_BUF_HFR_O, TGGCN_BUF_HFR_G};
    const int gates_buf[3] = {TGGCN_BUF_GATES_H, TGGCN_BUF_GATES_O, TGGCN_BUF_GATES_G};
    const int bd_id[3] = {TGGCN_W_HUM_BD_W, TGGCN_W_OBJ_BD_W, TGGCN_W_GEO_BD_W};
    const int Eg[3] = {H, O, 1};
    {
        float* wt = bb(BL.wt);
        for (int g = 0; g < 3; ++g) {
            const int M = N * Eg[g];
            const float* dZ = bb(BL.ds[g]) + D;
            const float* Y = buf(s_buf[g]) + D;
            const int beta = (g == 0 || (g == 1 && d.C_aff > 0)) ? 1 : 0;     // the frame heads already wrote into d hfr
            const WSrc wb[1] = {{W(bd_id[g]), 2 * D, D}};
            if (int rc = dx_gemm(dZ, 2 * D, Y, 2 * D, wb, 1, 2 * D, bb(BL.dhfr[g]), 2 * D, M, beta, wt)) return rc;
            if (int rc = tn(dZ, 2 * D, Y, 2 * D, buf(hfr_buf[g]), 2 * D, G(bd_id[g]), 2 * D, M, D, 2 * D, 0, 0, 0, stream, G(bd_id[g] + 1))) return rc;
        }
        if (int rc = tn_flush(stream)) return rc;
    }

    // ---- 5/4. BiGRU backward through time (all three groups in one persistent kernel), then the hoisted input projections ----------
    {
        const int wih_f[3] = {TGGCN_W_HUM_RNN_WIH_F, TGGCN_W_OBJ_RNN_WIH_F, TGGCN_W_GEO_RNN_WIH_F};
        // table order per group: WIH_F, WHH_F, BIH_F, BHH_F, WIH_B, WHH_B, BIH_B, BHH_B
        BiGruBwdParams P;
        memset(&P, 0, sizeof(P));
        P.ngroups = 3; P.B = B; P.T = T; P.D = D;
        float* whhT = bb(BL.bigru_scratch);
        for (int g = 0; g < 3; ++g) {
            BiGruBwdGroup& Gp = P.g[g];
            for (int dir = 0; dir < 2; ++dir) {
                float* wt = whhT + (size_t)(g * 2 + dir) * D * 3 * D;
                if (int rc = launch_transpose(W(wih_f[g] + 1 + 4 * dir), D, wt, 3 * D, 3 * D, D, stream)) return rc;      // (3D,D) -> (D,3D)
                Gp.whhT[dir] = wt;
            }
            Gp.dhfr = bb(BL.dhfr[g]); Gp.hfr = buf(hfr_buf[g]); Gp.gates = buf(gates_buf[g]);
            Gp.dgi = bb(BL.dgi[g]); Gp.dgh = bb(BL.dgh[g]); Gp.direct = bb(BL.gru_direct[g]);
            Gp.E = Eg[g]; Gp.rows = B * Eg[g];
        }
        unsigned int* sync = (unsigned int*)buf(TGGCN_BUF_SYNC);
        P.sync.counter = sync + 4; P.sync.error = sync + 5;
        if (int rc = launch_bigru_bwd(P, d.persistent, stream)) return rc;
        float* wt = bb(BL.wt);
        for (int g = 0; g < 3; ++g) {
            const int base = wih_f[g];
            const int M = N * Eg[g];
            float* dgi = bb(BL.dgi[g]);
            const float* dgh = bb(BL.dgh[g]);
            for (int dir = 0; dir < 2; ++dir) {
                // dW_hh = dGh^T h_{t-1} (fwd) / h_{t+1} (bwd): row shift inside each video
                if (int rc = tn(dgh + (size_t)dir * 3 * D, 6 * D, nullptr, 0, buf(hfr_buf[g]) + (size_t)dir * D, 2 * D,
                                            G(base + 1 + 4 * dir), D, M, 3 * D, D, dir == 0 ? -Eg[g] : Eg[g], T * Eg[g], 0, stream, G(base + 3 + 4 * dir))) return rc;
            }
            // d x (+)= [dGi_f | dGi_b] [W_ih_f ; W_ih_b]   (x = S[:, :D])
            const WSrc wi[2] = {{W(base), D, 3 * D}, {W(base + 4), D, 3 * D}};
            if (int rc = dx_gemm(dgi, 6 * D, nullptr, 0, wi, 2, D, bb(BL.ds[g]), 2 * D, M, 1, wt)) return rc;
            for (int dir = 0; dir < 2; ++dir) {
                if (int rc = tn(dgi + (size_t)dir * 3 * D, 6 * D, nullptr, 0, buf(s_buf[g]), 2 * D, G(base + 4 * dir), D, M, 3 * D, D,
                                            0, 0, 0, stream, G(base + 4 * dir + 2))) return rc;
            }
        }
        if (int rc = tn_flush(stream)) return rc;
    }

    if (hooks && hooks->bucket_done[1]) TG_CUDA_OK(cudaEventRecord((cudaEvent_t)hooks->bucket_done[1], stream));

    // ---- 3/2. embeddings (inputs are data: weight gradients only) and the geometry MLP ---------------------------------------------------
    {
        // x = ReLU(W roi + b) = S[:, :D]
        if (int rc = tn(bb(BL.ds[0]), 2 * D, buf(TGGCN_BUF_S_H), 2 * D, io->x_human, d.Fh, G(TGGCN_W_HUM_EMB_W), 2048, N * H, D, 2048,
                                    0, 0, 0, stream, G(TGGCN_W_HUM_EMB_B))) return rc;
        if (int rc = tn(bb(BL.ds[1]), 2 * D, buf(TGGCN_BUF_S_O), 2 * D, io->x_objects, 2048, G(TGGCN_W_OBJ_EMB_W), 2048, N * O, D, 2048,
                                    0, 0, 0, stream, G(TGGCN_W_OBJ_EMB_B))) return rc;
        // geometry MLP layer 2: S_G[:, :D] = ReLU(W2 hid + b2)
        float* wt = bb(BL.wt);
        const WSrc w2[1] = {{W(TGGCN_W_GEO_MLP2_W), 2048, D}};
        if (int rc = dx_gemm(bb(BL.ds[2]), 2 * D, buf(TGGCN_BUF_S_G), 2 * D, w2, 1, 2048, bb(BL.dgeo_hid), 2048, N, 0, wt)) return rc;
        if (int rc = tn(bb(BL.ds[2]), 2 * D, buf(TGGCN_BUF_S_G), 2 * D, buf(TGGCN_BUF_GEO_HID), 2048, G(TGGCN_W_GEO_MLP2_W), 2048, N, D,
                                    2048, 0, 0, 0, stream, G(TGGCN_W_GEO_MLP2_B))) return rc;
        // layer 0: hid = ReLU(W0 gcn + b0), gcn = the scrambled view (N, 128V)
        const int KV = 128 * V;
        const WSrc w0[1] = {{W(TGGCN_W_GEO_MLP0_W), KV, 2048}};
        if (int rc = dx_gemm(bb(BL.dgeo_hid), 2048, buf(TGGCN_BUF_GEO_HID), 2048, w0, 1, KV, bb(BL.dgcn_out), KV, N, 0, wt)) return rc;
        if (int rc = tn(bb(BL.dgeo_hid), 2048, buf(TGGCN_BUF_GEO_HID), 2048, buf(TGGCN_BUF_GCN_OUT), KV, G(TGGCN_W_GEO_MLP0_W), KV, N,
                                    2048, KV, 0, 0, 0, stream, G(TGGCN_W_GEO_MLP0_B))) return rc;
        if (int rc = tn_flush(stream)) return rc;
    }

    if (hooks && hooks->bucket_done[2]) TG_CUDA_OK(cudaEventRecord((cudaEvent_t)hooks->bucket_done[2], stream));

    // ---- 1. geometry GCN ---------------------------------------------------------------------------------------------------------------------
    {
        GcnBwdParams P;
        memset(&P, 0, sizeof(P));
        P.xh = io->x_human; P.B = B; P.T = T; P.H = H; P.V = V; P.Fh = d.Fh;
        if (d.bn_train) {
            const float* stats = buf(TGGCN_BUF_SEG_SCRATCH) + 2 * (size_t)B * H * nkh * D + 2 * (size_t)B * O * 2 * D;   // see api.cu
            P.mean = stats; P.var = stats + 4 * V;
        } else {
            P.mean = W(TGGCN_W_GCN_BN_MEAN); P.var = W(TGGCN_W_GCN_BN_VAR);
        }
        P.gamma = W(TGGCN_W_GCN_BN_W); P.beta = W(TGGCN_W_GCN_BN_B);
        P.w1 = W(TGGCN_W_GCN_C1_W); P.b1 = W(TGGCN_W_GCN_C1_B); P.w3 = W(TGGCN_W_GCN_C3_W); P.b3 = W(TGGCN_W_GCN_C3_B);
        P.ws1 = W(TGGCN_W_GCN_S1_W); P.bs1 = W(TGGCN_W_GCN_S1_B); P.ws2 = W(TGGCN_W_GCN_S2_W); P.bs2 = W(TGGCN_W_GCN_S2_B);
        P.wg = W(TGGCN_W_GCN_W);
        P.dout = bb(BL.dgcn_out);
        P.dwg = G(TGGCN_W_GCN_W); P.dws1 = G(TGGCN_W_GCN_S1_W); P.dbs1 = G(TGGCN_W_GCN_S1_B); P.dws2 = G(TGGCN_W_GCN_S2_W);
        P.dbs2 = G(TGGCN_W_GCN_S2_B); P.dw3 = G(TGGCN_W_GCN_C3_W); P.db3 = G(TGGCN_W_GCN_C3_B); P.dw1 = G(TGGCN_W_GCN_C1_W);
        P.db1 = G(TGGCN_W_GCN_C1_B);
        P.dxn = bb(BL.dxn);
        TG_CUDA_OK(cudaMemsetAsync(P.dwg, 0, sizeof(float) * 64 * 128, stream));
        TG_CUDA_OK(cudaMemsetAsync(P.dws1, 0, sizeof(float) * 128 * 64, stream));
        TG_CUDA_OK(cudaMemsetAsync(P.dbs1, 0, sizeof(float) * 128, stream));
        TG_CUDA_OK(cudaMemsetAsync(P.dws2, 0, sizeof(float) * 128 * 64, stream));
        TG_CUDA_OK(cudaMemsetAsync(P.dbs2, 0, sizeof(float) * 128, stream));
        TG_CUDA_OK(cudaMemsetAsync(P.dw3, 0, sizeof(float) * 64 * 64, stream));
        TG_CUDA_OK(cudaMemsetAsync(P.db3, 0, sizeof(float) * 64, stream));
        TG_CUDA_OK(cudaMemsetAsync(P.dw1, 0, sizeof(float) * 64 * 4, stream));
        TG_CUDA_OK(cudaMemsetAsync(P.db1, 0, sizeof(float) * 64, stream));
        if (int rc = launch_geo_gcn_bwd(P, stream)) return rc;
        if (int rc = launch_geo_bn_bwd(io->x_human, P.dxn, P.mean, P.var, P.gamma, G(TGGCN_W_GCN_BN_W), G(TGGCN_W_GCN_BN_B), B, T, H, V, d.Fh,
                                       stream)) return rc;
    }
    if (hooks && hooks->bucket_done[3]) TG_CUDA_OK(cudaEventRecord((cudaEvent_t)hooks->bucket_done[3], stream));
    if (io->status_host != nullptr)
        TG_CUDA_OK(cudaMemcpyAsync(io->status_host, buf(TGGCN_BUF_SYNC), 8 * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
    return 0;
}

}  // extern "C"
