// C ABI (include/tggcn_b200.h): workspace layout and the launch sequence of one TGGCN forward
// (vhoi/models.py:584-933).  No allocation, no device synchronisation, no state between calls.
#include <stdarg.h>
#include <stdlib.h>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "gemm.h"
#include "bigru.h"
#include "frame.h"
#include "api_internal.h"
#include "step_tc.cuh"

namespace tg {

static thread_local char g_err[1024] = "";
unsigned long long g_launches = 0;   // kernels launched by this library since load

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

bool debug_sync() {
    static int cached = -1;
    if (cached < 0) {
        const char* e = getenv("TGGCN_DEBUG_SYNC");
        cached = (e != nullptr && e[0] == '1') ? 1 : 0;
    }
    return cached == 1;
}

constexpr int MAX_DEVICES = 64;

int num_sms() {
    static int cached[MAX_DEVICES] = {0};
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAX_DEVICES) return 148;
    if (cached[dev] == 0) {
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) cached[dev] = n;
        else return 148;
    }
    return cached[dev];
}

int ensure_smem(const void* func, size_t bytes) {
    struct Entry { const void* func; size_t bytes; };
    static std::mutex mu;
    static std::vector<Entry> table[MAX_DEVICES];
    int dev = 0;
    TG_CUDA_OK(cudaGetDevice(&dev));
    TG_REQUIRE(dev >= 0 && dev < MAX_DEVICES, "device ordinal %d out of range", dev);
    std::lock_guard<std::mutex> lock(mu);
    for (Entry& e : table[dev])
        if (e.func == func) {
            if (e.bytes >= bytes) return 0;
            TG_CUDA_OK(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
            e.bytes = bytes;
            return 0;
        }
    TG_CUDA_OK(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    table[dev].push_back({func, bytes});
    return 0;
}

void make_layout(const tggcn_dims& d, Layout& L) {
    const size_t N = (size_t)d.B * d.T, H = d.H, O = d.O, D = d.D, V = d.V;
    const size_t f = sizeof(float);
    const size_t nkh = nkh_of(d);
    size_t sz[TGGCN_BUF_COUNT];
    sz[TGGCN_BUF_GCN_OUT] = N * 128 * V * f;
    sz[TGGCN_BUF_GEO_HID] = N * 2048 * f;
    sz[TGGCN_BUF_S_H] = N * H * 2 * D * f;
    sz[TGGCN_BUF_S_O] = N * O * 2 * D * f;
    sz[TGGCN_BUF_S_G] = N * 2 * D * f;
    sz[TGGCN_BUF_GI_H] = N * H * 6 * D * f;
    sz[TGGCN_BUF_GI_O] = N * O * 6 * D * f;
    sz[TGGCN_BUF_GI_G] = N * 6 * D * f;
    sz[TGGCN_BUF_HFR_H] = N * H * 2 * D * f;
    sz[TGGCN_BUF_HFR_O] = N * O * 2 * D * f;
    sz[TGGCN_BUF_HFR_G] = N * 2 * D * f;
    sz[TGGCN_BUF_MSG_HH] = N * H * D * f;
    sz[TGGCN_BUF_MSG_HO] = N * H * D * f;
    sz[TGGCN_BUF_MSG_OH] = N * O * D * f;
    sz[TGGCN_BUF_MSG_OO] = N * O * D * f;
    sz[TGGCN_BUF_MSG_GO] = N * D * f;
    sz[TGGCN_BUF_XX_H] = N * H * (size_t)kh_of(d) * f;
    sz[TGGCN_BUF_XX_O] = N * O * (size_t)ko_of(d) * f;
    sz[TGGCN_BUF_TIME_EMB] = d.time_position ? N * D * f : 0;
    sz[TGGCN_BUF_MSG_GH] = d.geo_to_human ? N * D * f : 0;
    sz[TGGCN_BUF_SEG_LEN] = d.segment_length ? N * (H + O) * f : 0;
    sz[TGGCN_BUF_GATE_IN_H] = gate2_of(d) ? N * H * (size_t)ginh_of(d) * f : 0;
    sz[TGGCN_BUF_GATE_IN_O] = gate2_of(d) ? N * O * (size_t)gino_of(d) * f : 0;
    sz[TGGCN_BUF_GATE_HID_H] = gate2_of(d) ? N * H * D * f : 0;
    sz[TGGCN_BUF_GATE_HID_O] = gate2_of(d) ? N * O * D * f : 0;
    sz[TGGCN_BUF_GATE_HID2_H] = gate3_of(d) ? N * H * D * f : 0;
    sz[TGGCN_BUF_GATE_HID2_O] = gate3_of(d) ? N * O * D * f : 0;
    sz[TGGCN_BUF_GS_H] = N * H * 6 * D * f;
    sz[TGGCN_BUF_GS_O] = N * O * 6 * D * f;
    sz[TGGCN_BUF_HX_H] = N * H * 2 * D * f;
    sz[TGGCN_BUF_HX_O] = N * O * 2 * D * f;
    sz[TGGCN_BUF_REIDX] = N * (H + O) * sizeof(int);
    sz[TGGCN_BUF_SEG_SCRATCH] = (2 * (size_t)d.B * H * nkh * D + 2 * (size_t)d.B * O * 2 * D + 8 * V) * f;
    sz[TGGCN_BUF_SYNC] = 64;
    sz[TGGCN_BUF_BIG] = 0;
    if (use_big_path(d, 0) || use_big_path(d, 1)) {   // 16-bit operand copies, state rings and message scratch of the large-batch recurrent path
        BigLayout BLy;
        big_layout(d.B, d.H, d.O, d.D, d.hh, BLy);
        sz[TGGCN_BUF_BIG] = BLy.total;
    }
    const size_t sv = d.save_for_backward ? 1 : 0;      // save buffers are empty in inference
    sz[TGGCN_BUF_GATES_H] = sv * N * H * 8 * D * f;
    sz[TGGCN_BUF_GATES_O] = sv * N * O * 8 * D * f;
    sz[TGGCN_BUF_GATES_G] = sv * N * 8 * D * f;
    sz[TGGCN_BUF_ALPHA_F] = sv * N * (H * H + 2 * H * O + O * O) * f;
    sz[TGGCN_BUF_PGATE] = sv * N * (H + O) * f;
    sz[TGGCN_BUF_SGATES_H] = sv * N * H * 8 * D * f;
    sz[TGGCN_BUF_SGATES_O] = sv * N * O * 8 * D * f;
    sz[TGGCN_BUF_MG_ALL_H] = sv * 2 * N * H * nkh * D * f;
    sz[TGGCN_BUF_MG_ALL_O] = sv * 2 * N * O * 2 * D * f;
    sz[TGGCN_BUF_SMSG_HH] = sv * 2 * N * H * D * f;
    sz[TGGCN_BUF_SMSG_OH] = sv * 2 * N * O * D * f;
    sz[TGGCN_BUF_SMSG_HO] = sv * 2 * N * H * D * f;
    sz[TGGCN_BUF_SMSG_OO] = sv * 2 * N * O * D * f;
    sz[TGGCN_BUF_SALPHA_HH] = sv * 2 * N * H * H * f;
    sz[TGGCN_BUF_SALPHA_OH] = sv * 2 * N * H * O * f;
    sz[TGGCN_BUF_SALPHA_HO] = sv * 2 * N * O * H * f;
    sz[TGGCN_BUF_SALPHA_OO] = sv * 2 * N * O * O * f;
    {   // operand planes of the largest projection stage (4 bytes per operand element: fp16 hi + lo, or bf16 + slack)
        const size_t kh = kh_of(d), ko = ko_of(d), KV = 128 * V;
        const size_t st[7] = {N * (H + O) * 2048 + N * KV + 2 * D * 2048 + 2048 * KV,
                              N * 2048 + D * 2048,
                              N * (H + O + 1) * D + 6 * 3 * D * D,
                              N * (H + O + 1) * 2 * D + 3 * D * 2 * D,
                              N * (H + O + 2) * 2 * D + 6 * D * 2 * D,
                              N * H * kh + N * O * ko + 2 * 3 * D * kh + 2 * 3 * D * ko,
                              gate2_of(d) ? N * H * (size_t)ginh_of(d) + N * O * (size_t)gino_of(d) + D * (size_t)(ginh_of(d) + gino_of(d)) : 0};
        size_t m = 0;
        for (size_t v : st) m = v > m ? v : m;
        sz[TGGCN_BUF_PACK] = m * 4 + 32 * 256;
    }
    size_t off = 0;
    for (int i = 0; i < TGGCN_BUF_COUNT; ++i) {
        L.off[i] = off;
        L.bytes[i] = sz[i];
        off += align_up(sz[i], 256);
    }
    L.total = off;
}

int check_dims(const tggcn_dims& d) {
    TG_REQUIRE(d.B >= 1 && d.T >= 1 && d.H >= 1 && d.O >= 1, "dims: B,T,H,O must be positive");
    TG_REQUIRE(d.D >= 16 && d.D % 16 == 0, "dims: hidden_size=%d must be a positive multiple of 16", d.D);
    TG_REQUIRE(d.V >= 1 && d.V <= 32, "dims: gcn_node=%d unsupported", d.V);
    TG_REQUIRE(d.Fh == 2048 + 4 * d.V, "dims: human feature size %d != 2048 + 4*gcn_node", d.Fh);
    TG_REQUIRE(d.C_sub >= 1 && d.C_sub <= 32 && d.C_aff >= 0 && d.C_aff <= 32, "dims: class counts out of range");
    TG_REQUIRE(d.gate_layers >= 0 && d.gate_layers <= 3, "dims: gate_layers=%d (discrete_networks_num_layers) must be 1, 2 or 3", d.gate_layers);
    TG_REQUIRE(d.time_position >= 0 && d.time_position <= 2 && (d.time_periodic == 0 || d.time_periodic == 1),
               "dims: time_position / time_periodic out of range");
    TG_REQUIRE((size_t)d.B * d.T * (size_t)(d.H > d.O ? d.H : d.O) * 6 * d.D < (1ull << 31),
               "dims: problem too large for 32-bit tile indexing");
    return 0;
}

}  // namespace tg

using namespace tg;

extern "C" {

int tggcn_abi_version(void) { return TGGCN_ABI_VERSION; }
const char* tggcn_last_error(void) { return g_err; }

size_t tggcn_workspace_bytes(const tggcn_dims* dims) {
    if (dims == nullptr || check_dims(*dims)) return 0;
    Layout L;
    make_layout(*dims, L);
    return L.total;
}

int tggcn_workspace_view(const tggcn_dims* dims, int buf_id, size_t* offset, size_t* bytes) {
    TG_REQUIRE(dims != nullptr && buf_id >= 0 && buf_id < TGGCN_BUF_COUNT, "workspace_view: bad arguments");
    if (int rc = check_dims(*dims)) return rc;
    Layout L;
    make_layout(*dims, L);
    if (offset) *offset = L.off[buf_id];
    if (bytes) *bytes = L.bytes[buf_id];
    return 0;
}

int tggcn_sync_status(const tggcn_dims* dims, const void* workspace, void* stream) {
    TG_REQUIRE(dims && workspace, "sync_status: null argument");
    if (int rc = check_dims(*dims)) return rc;
    Layout L;
    make_layout(*dims, L);
    unsigned int flags[8] = {0, 0, 0, 0, 0, 0, 0, 0};    // (counter, error) of bigru, segment, bigru_bwd, segment_bwd
    TG_CUDA_OK(cudaMemcpyAsync(flags, (const char*)workspace + L.off[TGGCN_BUF_SYNC], sizeof(flags), cudaMemcpyDeviceToHost,
                               (cudaStream_t)stream));
    TG_CUDA_OK(cudaStreamSynchronize((cudaStream_t)stream));
    return tggcn_status_decode(flags);
}

int tggcn_status_decode(const uint32_t* flags) {
    TG_REQUIRE(flags != nullptr, "status_decode: null argument");
    int rc = 0;
    if ((flags[1] | flags[3] | flags[5] | flags[7]) & 1u) {
        set_error("persistent kernel grid barrier timed out (bigru=%u, segment=%u, bigru_bwd=%u, segment_bwd=%u): the results of "
                  "that call are undefined", flags[1] & 1u, flags[3] & 1u, flags[5] & 1u, flags[7] & 1u);
        rc |= 1;
    }
    if ((flags[1] | flags[3] | flags[5] | flags[7]) & 2u) {
        if (!(rc & 1))
            set_error("a weight (|w| >= 255) or activation (>= 65504) left the range of the fp16-split tensor-core tiles (projections / bigru=%u, "
                      "segment=%u, backward GEMMs=%u): the results of that call are invalid; rerun with dims.no_fp16_split = 1 (3xTF32 kernels)",
                      (flags[1] >> 1) & 1u, (flags[3] >> 1) & 1u, ((flags[5] | flags[7]) >> 1) & 1u);
        rc |= 2;
    }
    return rc;
}

int tggcn_geo_gcn_fwd(const float* x_human, const void* const* weights, float* out, float* bn_running_mean,
                      float* bn_running_var, int64_t* bn_num_batches, void* workspace, int B, int T, int H, int V,
                      int Fh, int bn_train, void* stream) {
    TG_REQUIRE(x_human && weights && out, "geo_gcn_fwd: null pointer");
    return launch_geo_gcn(x_human, weights, out, bn_running_mean, bn_running_var, bn_num_batches, (float*)workspace, B,
                          T, H, V, Fh, bn_train, (cudaStream_t)stream);
}

int tggcn_linear_fwd(const float* A, int lda, const float* W, int ldw, const float* bias, float* C, int ldc, int M,
                     int N, int K, int relu, int gemm_path, void* stream) {
    TG_REQUIRE(A && W && C, "linear_fwd: null pointer");
    GemmGroup g;
    g.count = 0;
    gemm_add(g, A, lda, W, ldw, bias, C, ldc, M, N, K, relu);
    return launch_gemm(g, gemm_path, (cudaStream_t)stream);
}

size_t tggcn_linear16_scratch_bytes(int M, int N, int K) {
    if (M <= 0 || N <= 0 || K <= 0) return 0;
    return 256 + ((size_t)M * K * 4 + 255) / 256 * 256 + ((size_t)N * K * 4 + 255) / 256 * 256;      // scale words + operand planes
}

int tggcn_linear16_fwd(const float* A, int lda, const float* W, int ldw, const float* bias, float* C, int ldc, int M, int N,
                       int K, int relu, int precision, void* scratch, size_t scratch_bytes, uint32_t* status, void* stream) {
    TG_REQUIRE(A && W && C && scratch, "linear16_fwd: null pointer");
    GemmGroup g;
    g.count = 0;
    gemm_add(g, A, lda, W, ldw, bias, C, ldc, M, N, K, relu);
    return launch_gemm16(g, precision, scratch, scratch_bytes, status, (cudaStream_t)stream);
}

}  // extern "C"

static int forward_impl(const tggcn_dims* dims, const void* const* weights, int n_weights, const tggcn_io* io,
                        void* workspace, size_t workspace_bytes, void* stream_, cudaEvent_t* ev) {
    TG_REQUIRE(dims && weights && io && workspace, "forward: null argument");
    TG_REQUIRE(n_weights == TGGCN_W_COUNT, "forward: expected %d weight pointers, got %d", (int)TGGCN_W_COUNT, n_weights);
    const tggcn_dims& d = *dims;
    if (int rc = check_dims(d)) return rc;
    Layout L;
    make_layout(d, L);
    TG_REQUIRE(workspace_bytes >= L.total, "forward: workspace too small (%zu < %zu)", workspace_bytes, L.total);
    TG_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "forward: workspace must be 256-byte aligned");
    cudaStream_t stream = (cudaStream_t)stream_;
    const int B = d.B, T = d.T, H = d.H, O = d.O, D = d.D, V = d.V, N = B * T;
    const int nkh = nkh_of(d);
    auto W = [&](int id) { return (const float*)weights[id]; };
    auto buf = [&](int id) { return (float*)((char*)workspace + L.off[id]); };

    // required pointers
    TG_REQUIRE(io->x_human && io->x_objects && io->objects_mask, "forward: missing inputs");
    TG_REQUIRE(io->y_hs && io->y_hss && io->y_os && io->y_oss, "forward: missing gate outputs");
    for (int i = 0; i < 4; ++i) TG_REQUIRE(io->out_h[i], "forward: missing human head output %d", i);
    if (d.C_aff > 0)
        for (int i = 0; i < 4; ++i) TG_REQUIRE(io->out_o[i], "forward: missing object head output %d", i);
    TG_REQUIRE((d.human_seg_given != 0) == (io->human_seg != nullptr), "forward: human_seg_given flag disagrees with pointer");
    TG_REQUIRE((d.object_seg_given != 0) == (io->object_seg != nullptr), "forward: object_seg_given flag disagrees with pointer");
    TG_REQUIRE((d.human_seg_given && d.object_seg_given) || d.straight_through || io->noise, "forward: Gumbel noise tensor required");
    static const int required[] = {
        TGGCN_W_GCN_W, TGGCN_W_GCN_BN_W, TGGCN_W_GCN_BN_B, TGGCN_W_GCN_BN_MEAN, TGGCN_W_GCN_BN_VAR, TGGCN_W_GCN_C1_W,
        TGGCN_W_GCN_C1_B, TGGCN_W_GCN_C3_W, TGGCN_W_GCN_C3_B, TGGCN_W_GCN_S1_W, TGGCN_W_GCN_S1_B, TGGCN_W_GCN_S2_W,
        TGGCN_W_GCN_S2_B, TGGCN_W_GEO_MLP0_W, TGGCN_W_GEO_MLP0_B, TGGCN_W_GEO_MLP2_W, TGGCN_W_GEO_MLP2_B,
        TGGCN_W_HUM_EMB_W, TGGCN_W_HUM_EMB_B, TGGCN_W_OBJ_EMB_W, TGGCN_W_OBJ_EMB_B, TGGCN_W_GEO_BD_W, TGGCN_W_HUM_BD_W,
        TGGCN_W_OBJ_BD_W, TGGCN_W_MSG_HO_W, TGGCN_W_MSG_OH_W, TGGCN_W_MSG_OO_W, TGGCN_W_MSG_GO_W, TGGCN_W_SMSG_HO_W,
        TGGCN_W_SMSG_OH_W, TGGCN_W_SMSG_OO_W, TGGCN_W_UPD_H_W, TGGCN_W_HSEG_F_WIH, TGGCN_W_OSEG_F_WIH,
        TGGCN_W_HEAD_H_FREC_W, TGGCN_W_HEAD_H_FPRED_W, TGGCN_W_HEAD_H_REC_W, TGGCN_W_HEAD_H_PRED_W};
    for (size_t i = 0; i < sizeof(required) / sizeof(required[0]); ++i)
        TG_REQUIRE(weights[required[i]] != nullptr, "forward: weight #%d is null", required[i]);
    if (d.hh) TG_REQUIRE(W(TGGCN_W_MSG_HH_W) && W(TGGCN_W_SMSG_HH_W), "forward: humans->human weights missing");
    TG_REQUIRE(d.update_strategy >= 0 && d.update_strategy <= 2, "forward: update_strategy %d unknown", d.update_strategy);
    if (d.update_strategy != 0)
        TG_REQUIRE(H == 1 && !d.human_seg_given && !d.object_seg_given && (d.update_strategy == 1 || !d.filter),
                   "forward: update_strategy 'sah'/'coh' acts with one human and sampled gates only ('coh' without the filter); pass 0 otherwise");
    if (d.update_strategy != 1) TG_REQUIRE(W(TGGCN_W_UPD_O_W) && W(TGGCN_W_UPD_O_B), "forward: object gate weights missing");
    if (d.C_aff > 0) TG_REQUIRE(W(TGGCN_W_HEAD_O_REC_W) && W(TGGCN_W_HEAD_O_FREC_W), "forward: object head weights missing");

    float* scratch = buf(TGGCN_BUF_SEG_SCRATCH);
    float* mg_h = scratch;
    float* mg_o = mg_h + 2 * (size_t)B * H * nkh * D;
    float* bn_stats = mg_o + 2 * (size_t)B * O * 2 * D;
    unsigned int* sync = (unsigned int*)buf(TGGCN_BUF_SYNC);

    int stage = 0;
#define STAGE_END()                                                  \
    do {                                                             \
        if (ev) TG_CUDA_OK(cudaEventRecord(ev[stage + 1], stream));  \
        ++stage;                                                     \
    } while (0)
    if (ev) TG_CUDA_OK(cudaEventRecord(ev[0], stream));
    TG_CUDA_OK(cudaMemsetAsync(sync, 0, L.bytes[TGGCN_BUF_SYNC], stream));     // barrier counters + error flags
    // 1. geometry GCN (stored (B,128,V,T); the scrambled view is a reinterpretation as (B*T, 128V))
    if (int rc = launch_geo_gcn(io->x_human, weights, buf(TGGCN_BUF_GCN_OUT), io->bn_running_mean, io->bn_running_var,
                                io->bn_num_batches, bn_stats, B, T, H, V, d.Fh, d.bn_train, stream))
        return rc;

    STAGE_END();
    const int gpath = (d.precision == 1 && d.gemm_path != 0) ? 3 : d.gemm_path;      // bf16 operands on the tensor-core projections
    GemmGroup g;
    // projections: the TMA-fed kernel on 16-bit operand planes (gemm16.cu) where the shapes qualify, else gemm_tc / SIMT
    auto project = [&](GemmGroup& grp) -> int {
        // Narrow outputs (hidden_size 64: N = 64) are bound by reading A once; packing A first would triple that traffic
        // (profiles: Bimanual D=64, pack16x = 11 % of a training step) — those problems stay on the register-producer kernel.
        GemmGroup wide, narrow;
        wide.count = narrow.count = 0;
        wide.precision = narrow.precision = 0;
        for (int i = 0; i < grp.count; ++i) {
            GemmGroup& dst = grp.p[i].N >= 128 ? wide : narrow;
            dst.p[dst.count++] = grp.p[i];
        }
        if (d.gemm_path != 0 && wide.count > 0 && (d.precision == 1 || !d.no_fp16_split) && gemm16_eligible(wide) &&
            gemm16_scratch_bytes(wide) <= L.bytes[TGGCN_BUF_PACK]) {
            if (int rc = launch_gemm16(wide, d.precision == 1 ? 1 : 0, buf(TGGCN_BUF_PACK), L.bytes[TGGCN_BUF_PACK],
                                       d.precision == 1 ? nullptr : sync + 1, stream))
                return rc;
            return narrow.count > 0 ? launch_gemm(narrow, gpath, stream) : 0;
        }
        return launch_gemm(grp, gpath, stream);
    };
    // 2. ROI embeddings and the first geometry MLP layer (models.py:646)
    g.count = 0;
    gemm_add(g, io->x_human, d.Fh, W(TGGCN_W_HUM_EMB_W), 2048, W(TGGCN_W_HUM_EMB_B), buf(TGGCN_BUF_S_H), 2 * D, N * H, D, 2048, 1);
    gemm_add(g, io->x_objects, 2048, W(TGGCN_W_OBJ_EMB_W), 2048, W(TGGCN_W_OBJ_EMB_B), buf(TGGCN_BUF_S_O), 2 * D, N * O, D, 2048, 1);
    gemm_add(g, buf(TGGCN_BUF_GCN_OUT), 128 * V, W(TGGCN_W_GEO_MLP0_W), 128 * V, W(TGGCN_W_GEO_MLP0_B), buf(TGGCN_BUF_GEO_HID), 2048, N, 2048, 128 * V, 1);
    if (int rc = project(g)) return rc;
    STAGE_END();
    // 3. second geometry MLP layer
    g.count = 0;
    gemm_add(g, buf(TGGCN_BUF_GEO_HID), 2048, W(TGGCN_W_GEO_MLP2_W), 2048, W(TGGCN_W_GEO_MLP2_B), buf(TGGCN_BUF_S_G), 2 * D, N, D, 2048, 1);
    if (int rc = project(g)) return rc;
    STAGE_END();
    // 4. BiGRU input pre-activations for both directions (hoisted W_ih x + b_ih)
    g.count = 0;
    gemm_add(g, buf(TGGCN_BUF_S_H), 2 * D, W(TGGCN_W_HUM_RNN_WIH_F), D, W(TGGCN_W_HUM_RNN_BIH_F), buf(TGGCN_BUF_GI_H), 6 * D, N * H, 3 * D, D, 0);
    gemm_add(g, buf(TGGCN_BUF_S_H), 2 * D, W(TGGCN_W_HUM_RNN_WIH_B), D, W(TGGCN_W_HUM_RNN_BIH_B), buf(TGGCN_BUF_GI_H) + 3 * D, 6 * D, N * H, 3 * D, D, 0);
    gemm_add(g, buf(TGGCN_BUF_S_O), 2 * D, W(TGGCN_W_OBJ_RNN_WIH_F), D, W(TGGCN_W_OBJ_RNN_BIH_F), buf(TGGCN_BUF_GI_O), 6 * D, N * O, 3 * D, D, 0);
    gemm_add(g, buf(TGGCN_BUF_S_O), 2 * D, W(TGGCN_W_OBJ_RNN_WIH_B), D, W(TGGCN_W_OBJ_RNN_BIH_B), buf(TGGCN_BUF_GI_O) + 3 * D, 6 * D, N * O, 3 * D, D, 0);
    gemm_add(g, buf(TGGCN_BUF_S_G), 2 * D, W(TGGCN_W_GEO_RNN_WIH_F), D, W(TGGCN_W_GEO_RNN_BIH_F), buf(TGGCN_BUF_GI_G), 6 * D, N, 3 * D, D, 0);
    gemm_add(g, buf(TGGCN_BUF_S_G), 2 * D, W(TGGCN_W_GEO_RNN_WIH_B), D, W(TGGCN_W_GEO_RNN_BIH_B), buf(TGGCN_BUF_GI_G) + 3 * D, 6 * D, N, 3 * D, D, 0);
    if (int rc = project(g)) return rc;
    STAGE_END();
    // 5. frame-level BiGRU recurrences (models.py:649-651)
    {
        BiGruParams P;
        memset(&P, 0, sizeof(P));
        P.ngroups = 3; P.B = B; P.T = T; P.D = D;
        const int gi_id[3] = {TGGCN_BUF_GI_H, TGGCN_BUF_GI_O, TGGCN_BUF_GI_G};
        const int hfr_id[3] = {TGGCN_BUF_HFR_H, TGGCN_BUF_HFR_O, TGGCN_BUF_HFR_G};
        const int whh_f[3] = {TGGCN_W_HUM_RNN_WHH_F, TGGCN_W_OBJ_RNN_WHH_F, TGGCN_W_GEO_RNN_WHH_F};
        const int whh_b[3] = {TGGCN_W_HUM_RNN_WHH_B, TGGCN_W_OBJ_RNN_WHH_B, TGGCN_W_GEO_RNN_WHH_B};
        const int bhh_f[3] = {TGGCN_W_HUM_RNN_BHH_F, TGGCN_W_OBJ_RNN_BHH_F, TGGCN_W_GEO_RNN_BHH_F};
        const int bhh_b[3] = {TGGCN_W_HUM_RNN_BHH_B, TGGCN_W_OBJ_RNN_BHH_B, TGGCN_W_GEO_RNN_BHH_B};
        const int E[3] = {H, O, 1};
        for (int i = 0; i < 3; ++i) {
            P.g[i].gi = buf(gi_id[i]); P.g[i].hfr = buf(hfr_id[i]);
            const int gates_id[3] = {TGGCN_BUF_GATES_H, TGGCN_BUF_GATES_O, TGGCN_BUF_GATES_G};
            P.g[i].gates = d.save_for_backward ? buf(gates_id[i]) : nullptr;
            P.g[i].whh[0] = W(whh_f[i]); P.g[i].whh[1] = W(whh_b[i]);
            P.g[i].bhh[0] = W(bhh_f[i]); P.g[i].bhh[1] = W(bhh_b[i]);
            P.g[i].E = E[i]; P.g[i].rows = B * E[i];
            TG_REQUIRE(P.g[i].whh[0] && P.g[i].whh[1] && P.g[i].bhh[0] && P.g[i].bhh[1], "forward: BiGRU weights missing");
        }
        P.sync.counter = sync; P.sync.error = sync + 1;
        P.no_fp16_split = d.no_fp16_split;
        P.big_ws = use_big_path(d, 0) ? (void*)buf(TGGCN_BUF_BIG) : nullptr; P.precision = d.precision;
        if (int rc = launch_bigru(P, d.persistent, stream)) return rc;
    }
    STAGE_END();
    // 6. Linear(2D->D)+ReLU on the BiGRU outputs, written next to x in the [x | h] rows
    g.count = 0;
    gemm_add(g, buf(TGGCN_BUF_HFR_H), 2 * D, W(TGGCN_W_HUM_BD_W), 2 * D, W(TGGCN_W_HUM_BD_B), buf(TGGCN_BUF_S_H) + D, 2 * D, N * H, D, 2 * D, 1);
    gemm_add(g, buf(TGGCN_BUF_HFR_O), 2 * D, W(TGGCN_W_OBJ_BD_W), 2 * D, W(TGGCN_W_OBJ_BD_B), buf(TGGCN_BUF_S_O) + D, 2 * D, N * O, D, 2 * D, 1);
    gemm_add(g, buf(TGGCN_BUF_HFR_G), 2 * D, W(TGGCN_W_GEO_BD_W), 2 * D, W(TGGCN_W_GEO_BD_B), buf(TGGCN_BUF_S_G) + D, 2 * D, N, D, 2 * D, 1);
    if (int rc = project(g)) return rc;
    STAGE_END();
    // 7. per-sender frame messages, each computed once per sender and message kind (models.py:1693-1718)
    g.count = 0;
    if (d.hh) gemm_add(g, buf(TGGCN_BUF_S_H), 2 * D, W(TGGCN_W_MSG_HH_W), 2 * D, W(TGGCN_W_MSG_HH_B), buf(TGGCN_BUF_MSG_HH), D, N * H, D, 2 * D, 1);
    gemm_add(g, buf(TGGCN_BUF_S_H), 2 * D, W(TGGCN_W_MSG_HO_W), 2 * D, W(TGGCN_W_MSG_HO_B), buf(TGGCN_BUF_MSG_HO), D, N * H, D, 2 * D, 1);
    gemm_add(g, buf(TGGCN_BUF_S_O), 2 * D, W(TGGCN_W_MSG_OH_W), 2 * D, W(TGGCN_W_MSG_OH_B), buf(TGGCN_BUF_MSG_OH), D, N * O, D, 2 * D, 1);
    gemm_add(g, buf(TGGCN_BUF_S_O), 2 * D, W(TGGCN_W_MSG_OO_W), 2 * D, W(TGGCN_W_MSG_OO_B), buf(TGGCN_BUF_MSG_OO), D, N * O, D, 2 * D, 1);
    gemm_add(g, buf(TGGCN_BUF_S_G), 2 * D, W(TGGCN_W_MSG_GO_W), 2 * D, W(TGGCN_W_MSG_GO_B), buf(TGGCN_BUF_MSG_GO), D, N, D, 2 * D, 1);
    if (d.geo_to_human) {
        TG_REQUIRE(W(TGGCN_W_MSG_GH_W) && W(TGGCN_W_MSG_GH_B), "forward: geometry_to_human_message_mlp weights missing");
        gemm_add(g, buf(TGGCN_BUF_S_G), 2 * D, W(TGGCN_W_MSG_GH_W), 2 * D, W(TGGCN_W_MSG_GH_B), buf(TGGCN_BUF_MSG_GH), D, N, D, 2 * D, 1);
    }
    if (int rc = project(g)) return rc;
    STAGE_END();
    // 8. attention, aggregation, gates, segment-level inputs
    if (d.time_position) {      // time-position features of every frame (models.py:656-662 / :755-762)
        TG_REQUIRE(io->steps_per_example != nullptr, "forward: add_time_position needs steps_per_example");
        TG_REQUIRE(d.time_periodic ? io->time_freq != nullptr : (W(TGGCN_W_TIME_W) && W(TGGCN_W_TIME_B)),
                   "forward: time-position parameters missing (time_position_mlp, or the period table of the periodic encoding)");
        if (int rc = launch_time_embed(io->steps_per_example, W(TGGCN_W_TIME_W), W(TGGCN_W_TIME_B), io->time_freq,
                                       buf(TGGCN_BUF_TIME_EMB), B, T, D, d.time_periodic, stream))
            return rc;
    }
    {
        FrameMsgParams P;
        memset(&P, 0, sizeof(P));
        P.B = B; P.T = T; P.H = H; P.O = O; P.D = D; P.hh = d.hh; P.thr = d.thr; P.mean_pool = d.mean_pool; P.att_noscale = d.att_noscale;
        P.update_strategy = d.update_strategy; P.straight_through = d.straight_through;
        P.gh = d.geo_to_human ? 1 : 0; P.msg_gh = d.geo_to_human ? buf(TGGCN_BUF_MSG_GH) : nullptr;
        P.tl = tl_of(d);
        P.dist[0] = d.hh ? io->dist_hh : nullptr; P.dist[1] = io->dist_ho; P.dist[2] = io->dist_oo;
        P.time_position = d.time_position; P.time_emb = d.time_position ? buf(TGGCN_BUF_TIME_EMB) : nullptr;
        P.s_h = buf(TGGCN_BUF_S_H); P.s_o = buf(TGGCN_BUF_S_O);
        P.msg_hh = buf(TGGCN_BUF_MSG_HH); P.msg_ho = buf(TGGCN_BUF_MSG_HO); P.msg_oh = buf(TGGCN_BUF_MSG_OH);
        P.msg_oo = buf(TGGCN_BUF_MSG_OO); P.msg_go = buf(TGGCN_BUF_MSG_GO);
        P.om = io->objects_mask;
        P.w_uh = W(TGGCN_W_UPD_H_W); P.b_uh = W(TGGCN_W_UPD_H_B);
        P.w_uo = W(TGGCN_W_UPD_O_W); P.b_uo = W(TGGCN_W_UPD_O_B);
        P.noise = io->noise; P.human_seg = io->human_seg; P.object_seg = io->object_seg;
        P.xx_h = buf(TGGCN_BUF_XX_H); P.xx_o = buf(TGGCN_BUF_XX_O);
        P.y_hs = io->y_hs; P.y_hss = io->y_hss; P.y_os = io->y_os; P.y_oss = io->y_oss;
        P.att_frame = d.inspect ? io->att_frame : nullptr;
        P.alpha_save = d.save_for_backward ? buf(TGGCN_BUF_ALPHA_F) : nullptr;
        P.pgate_save = d.save_for_backward ? buf(TGGCN_BUF_PGATE) : nullptr;
        const bool sample_h = !d.human_seg_given, sample_o = !d.object_seg_given && d.update_strategy != 1;
        if (gate2_of(d) && (sample_h || sample_o)) {
            // discrete_networks_num_layers == 2 (models.py:532-547): gate inputs -> projection (Linear(in, D) + ReLU) -> layer 2 + sampling
            P.gate_in_h = buf(TGGCN_BUF_GATE_IN_H); P.gin_h = ginh_of(d);
            P.gate_in_o = buf(TGGCN_BUF_GATE_IN_O); P.gin_o = gino_of(d);
            if (int rc = launch_frame_messages(P, stream)) return rc;
            g.count = 0;
            if (sample_h) {
                TG_REQUIRE(W(TGGCN_W_UPD_H_W2) && W(TGGCN_W_UPD_H_B2) && W(TGGCN_W_UPD_H_B), "forward: two-layer human gate MLP weights missing");
                gemm_add(g, P.gate_in_h, P.gin_h, W(TGGCN_W_UPD_H_W), P.gin_h, W(TGGCN_W_UPD_H_B), buf(TGGCN_BUF_GATE_HID_H), D, N * H, D, P.gin_h, 1);
            }
            if (sample_o) {
                TG_REQUIRE(W(TGGCN_W_UPD_O_W2) && W(TGGCN_W_UPD_O_B2) && W(TGGCN_W_UPD_O_B), "forward: two-layer object gate MLP weights missing");
                gemm_add(g, P.gate_in_o, P.gin_o, W(TGGCN_W_UPD_O_W), P.gin_o, W(TGGCN_W_UPD_O_B), buf(TGGCN_BUF_GATE_HID_O), D, N * O, D, P.gin_o, 1);
            }
            if (int rc = project(g)) return rc;
            P.gate_hid_h = buf(TGGCN_BUF_GATE_HID_H); P.gate_hid_o = buf(TGGCN_BUF_GATE_HID_O);
            int last_h = TGGCN_W_UPD_H_W2, last_o = TGGCN_W_UPD_O_W2;         // weight slot of the Linear(D, 1) layer (bias = slot + 1)
            if (gate3_of(d)) {                                                // one more Linear(D, D) + ReLU
                g.count = 0;
                if (sample_h) gemm_add(g, P.gate_hid_h, D, W(TGGCN_W_UPD_H_W2), D, W(TGGCN_W_UPD_H_B2), buf(TGGCN_BUF_GATE_HID2_H), D, N * H, D, D, 1);
                if (sample_o) gemm_add(g, P.gate_hid_o, D, W(TGGCN_W_UPD_O_W2), D, W(TGGCN_W_UPD_O_B2), buf(TGGCN_BUF_GATE_HID2_O), D, N * O, D, D, 1);
                if (int rc = project(g)) return rc;
                P.gate_hid_h = buf(TGGCN_BUF_GATE_HID2_H); P.gate_hid_o = buf(TGGCN_BUF_GATE_HID2_O);
                last_h = TGGCN_W_UPD_H_W4; last_o = TGGCN_W_UPD_O_W4;
                if (sample_h) TG_REQUIRE(W(last_h) && W(last_h + 1), "forward: three-layer human gate MLP weights missing");
                if (sample_o) TG_REQUIRE(W(last_o) && W(last_o + 1), "forward: three-layer object gate MLP weights missing");
            }
            P.w_uh = W(last_h); P.b_uh = W(last_h + 1);
            P.w_uo = W(last_o); P.b_uo = W(last_o + 1);
            if (int rc = launch_gate_sample(P, stream)) return rc;
        } else {
            if (int rc = launch_frame_messages(P, stream)) return rc;
        }
    }
    STAGE_END();
    // 9. optional local-maximum filter + reorder gather index
    if (int rc = launch_gate_post(io->y_hs, io->y_hss, io->y_os, io->y_oss, (int*)buf(TGGCN_BUF_REIDX), B, T, H, O, d.filter,
                                  d.thr, stream))
        return rc;
    if (d.segment_length) {     // segment lengths from the final hard gates, embedded into the last block of every xx row (models.py:763-779)
        TG_REQUIRE(io->steps_per_example != nullptr, "forward: add_segment_length needs steps_per_example");
        TG_REQUIRE(d.time_periodic ? io->time_freq != nullptr : (W(TGGCN_W_LEN_W) && W(TGGCN_W_LEN_B)),
                   "forward: segment-length parameters missing (segment_length_mlp, or the period table of the periodic encoding)");
        if (int rc = launch_segment_length(io->y_hs, io->y_os, io->steps_per_example, W(TGGCN_W_LEN_W), W(TGGCN_W_LEN_B), io->time_freq,
                                           buf(TGGCN_BUF_SEG_LEN), buf(TGGCN_BUF_XX_H), kh_of(d), buf(TGGCN_BUF_XX_O), ko_of(d),
                                           B, T, H, O, D, d.time_periodic, stream))
            return rc;
    }
    STAGE_END();
    // 10. hoisted frame-part of the segment cells' W_ih x + b_ih, both directions
    const int kh = kh_of(d), ldwh = ldwh_of(d), ko = ko_of(d), ldwo = ldwo_of(d);
    g.count = 0;
    gemm_add(g, buf(TGGCN_BUF_XX_H), kh, W(TGGCN_W_HSEG_F_WIH), ldwh, W(TGGCN_W_HSEG_F_BIH), buf(TGGCN_BUF_GS_H), 6 * D, N * H, 3 * D, kh, 0);
    gemm_add(g, buf(TGGCN_BUF_XX_H), kh, W(TGGCN_W_HSEG_B_WIH), ldwh, W(TGGCN_W_HSEG_B_BIH), buf(TGGCN_BUF_GS_H) + 3 * D, 6 * D, N * H, 3 * D, kh, 0);
    gemm_add(g, buf(TGGCN_BUF_XX_O), ko, W(TGGCN_W_OSEG_F_WIH), ldwo, W(TGGCN_W_OSEG_F_BIH), buf(TGGCN_BUF_GS_O), 6 * D, N * O, 3 * D, ko, 0);
    gemm_add(g, buf(TGGCN_BUF_XX_O), ko, W(TGGCN_W_OSEG_B_WIH), ldwo, W(TGGCN_W_OSEG_B_BIH), buf(TGGCN_BUF_GS_O) + 3 * D, 6 * D, N * O, 3 * D, ko, 0);
    if (int rc = project(g)) return rc;
    STAGE_END();
    // 11. segment-level recurrent graph (models.py:785-880)
    {
        SegParams P;
        memset(&P, 0, sizeof(P));
        P.B = B; P.T = T; P.H = H; P.O = O; P.D = D; P.hh = d.hh; P.mean_pool = d.mean_pool; P.att_noscale = d.att_noscale;
        P.gs_h = buf(TGGCN_BUF_GS_H); P.gs_o = buf(TGGCN_BUF_GS_O);
        P.dist[0] = d.hh ? io->dist_hh : nullptr; P.dist[1] = io->dist_ho; P.dist[2] = io->dist_oo;
        P.u_h = io->y_hs; P.u_o = io->y_os; P.om = io->objects_mask;
        P.wih_h[0] = W(TGGCN_W_HSEG_F_WIH); P.wih_h[1] = W(TGGCN_W_HSEG_B_WIH); P.ldw_h = ldwh; P.col_h = kh;
        P.wih_o[0] = W(TGGCN_W_OSEG_F_WIH); P.wih_o[1] = W(TGGCN_W_OSEG_B_WIH); P.ldw_o = ldwo; P.col_o = ko;
        P.whh_h[0] = W(TGGCN_W_HSEG_F_WHH); P.whh_h[1] = W(TGGCN_W_HSEG_B_WHH);
        P.bhh_h[0] = W(TGGCN_W_HSEG_F_BHH); P.bhh_h[1] = W(TGGCN_W_HSEG_B_BHH);
        P.whh_o[0] = W(TGGCN_W_OSEG_F_WHH); P.whh_o[1] = W(TGGCN_W_OSEG_B_WHH);
        P.bhh_o[0] = W(TGGCN_W_OSEG_F_BHH); P.bhh_o[1] = W(TGGCN_W_OSEG_B_BHH);
        P.wm[0] = W(TGGCN_W_SMSG_HH_W); P.bm[0] = W(TGGCN_W_SMSG_HH_B);
        P.wm[1] = W(TGGCN_W_SMSG_OH_W); P.bm[1] = W(TGGCN_W_SMSG_OH_B);
        P.wm[2] = W(TGGCN_W_SMSG_HO_W); P.bm[2] = W(TGGCN_W_SMSG_HO_B);
        P.wm[3] = W(TGGCN_W_SMSG_OO_W); P.bm[3] = W(TGGCN_W_SMSG_OO_B);
        for (int i = 0; i < 2; ++i)
            TG_REQUIRE(P.wih_h[i] && P.wih_o[i] && P.whh_h[i] && P.whh_o[i] && P.bhh_h[i] && P.bhh_o[i],
                       "forward: segment cell weights missing");
        P.hx_h = buf(TGGCN_BUF_HX_H); P.hx_o = buf(TGGCN_BUF_HX_O);
        P.mg_h = mg_h; P.mg_o = mg_o;
        P.mg_T = 1;
        if (d.save_for_backward) {
            P.mg_h = buf(TGGCN_BUF_MG_ALL_H); P.mg_o = buf(TGGCN_BUF_MG_ALL_O); P.mg_T = T;
            P.sgates_h = buf(TGGCN_BUF_SGATES_H); P.sgates_o = buf(TGGCN_BUF_SGATES_O);
            P.smsg[0] = d.hh ? buf(TGGCN_BUF_SMSG_HH) : nullptr; P.smsg[1] = buf(TGGCN_BUF_SMSG_OH);
            P.smsg[2] = buf(TGGCN_BUF_SMSG_HO); P.smsg[3] = buf(TGGCN_BUF_SMSG_OO);
            P.salpha[0] = d.hh ? buf(TGGCN_BUF_SALPHA_HH) : nullptr; P.salpha[1] = buf(TGGCN_BUF_SALPHA_OH);
            P.salpha[2] = buf(TGGCN_BUF_SALPHA_HO); P.salpha[3] = buf(TGGCN_BUF_SALPHA_OO);
        }
        P.att_f = d.inspect ? io->att_seg_f : nullptr;
        P.att_b = d.inspect ? io->att_seg_b : nullptr;
        P.sync.counter = sync + 2; P.sync.error = sync + 3;
        P.no_fp16_split = d.no_fp16_split;
        P.big_ws = use_big_path(d, 1) ? (void*)buf(TGGCN_BUF_BIG) : nullptr; P.precision = d.precision;
        if (int rc = launch_segment(P, d.persistent, stream)) return rc;
    }
    STAGE_END();
    // 12. label heads (models.py:909-917)
    {
        HeadsParams P;
        memset(&P, 0, sizeof(P));
        P.B = B; P.T = T; P.E = H; P.NE = H + O; P.e_off = 0; P.D = D; P.C = d.C_sub; P.cat = d.cat_level_states;
        P.hfr = buf(TGGCN_BUF_HFR_H); P.hx = buf(TGGCN_BUF_HX_H); P.reidx = (const int*)buf(TGGCN_BUF_REIDX);
        const int wid[4] = {TGGCN_W_HEAD_H_FREC_W, TGGCN_W_HEAD_H_FPRED_W, TGGCN_W_HEAD_H_REC_W, TGGCN_W_HEAD_H_PRED_W};
        for (int i = 0; i < 4; ++i) { P.w[i] = W(wid[i]); P.b[i] = W(wid[i] + 1); P.out[i] = io->out_h[i]; }
        if (int rc = launch_heads(P, stream)) return rc;
        if (d.C_aff > 0) {
            P.E = O; P.e_off = H; P.C = d.C_aff;
            P.hfr = buf(TGGCN_BUF_HFR_O); P.hx = buf(TGGCN_BUF_HX_O);
            const int oid[4] = {TGGCN_W_HEAD_O_FREC_W, TGGCN_W_HEAD_O_FPRED_W, TGGCN_W_HEAD_O_REC_W, TGGCN_W_HEAD_O_PRED_W};
            for (int i = 0; i < 4; ++i) { P.w[i] = W(oid[i]); P.b[i] = W(oid[i] + 1); P.out[i] = io->out_o[i]; }
            if (int rc = launch_heads(P, stream)) return rc;
        }
    }
    STAGE_END();
#undef STAGE_END
    if (io->status_host != nullptr)
        TG_CUDA_OK(cudaMemcpyAsync(io->status_host, sync, 8 * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
    return 0;
}

extern "C" {

int tggcn_forward(const tggcn_dims* dims, const void* const* weights, int n_weights, const tggcn_io* io,
                  void* workspace, size_t workspace_bytes, void* stream) {
    return forward_impl(dims, weights, n_weights, io, workspace, workspace_bytes, stream, nullptr);
}

int tggcn_forward_profile(const tggcn_dims* dims, const void* const* weights, int n_weights, const tggcn_io* io,
                          void* workspace, size_t workspace_bytes, void* stream, float* stage_ms_host) {
    TG_REQUIRE(stage_ms_host != nullptr, "forward_profile: null stage_ms_host");
    cudaEvent_t ev[TGGCN_STAGE_COUNT + 1];
    for (int i = 0; i <= TGGCN_STAGE_COUNT; ++i) TG_CUDA_OK(cudaEventCreate(&ev[i]));
    int rc = forward_impl(dims, weights, n_weights, io, workspace, workspace_bytes, stream, ev);
    if (rc == 0) {
        cudaError_t e = cudaStreamSynchronize((cudaStream_t)stream);
        if (e != cudaSuccess) { set_error("forward_profile: %s", cudaGetErrorString(e)); rc = 1; }
    }
    for (int i = 0; i < TGGCN_STAGE_COUNT && rc == 0; ++i) {
        if (cudaEventElapsedTime(&stage_ms_host[i], ev[i], ev[i + 1]) != cudaSuccess) { set_error("forward_profile: event timing failed"); rc = 1; }
    }
    for (int i = 0; i <= TGGCN_STAGE_COUNT; ++i) cudaEventDestroy(ev[i]);
    return rc;
}

unsigned long long tggcn_launch_count(void) { return tg::g_launches; }


}  // extern "C"
