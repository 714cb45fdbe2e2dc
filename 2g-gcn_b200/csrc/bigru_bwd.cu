// C-ABI entries of the single-group frame-level bidirectional GRU: forward with saved gates, and the backward through
// time (autograd of nn.GRU as called at vhoi/models.py:997-1000).
//
// The reverse-time loop is the persistent kernel of recurrent_bwd.cu:
//   dh      = dHFR[t] + carry
//   GRU cell backward (elementwise, saved r, z, n, hn):      dGi = [d a_r, d a_z, d a_n]   (gradient of W_ih x + b_ih)
//                                                           dGh = [d a_r, d a_z, r (.) d a_n] (gradient of W_hh h + b_hh)
//   carry   = z (.) dh + dGh W_hh                            (gate tiles on W_hh^T)
// After the loop the weight gradients are two GEMMs over all (video, t, entity) rows:
//   dW_hh = dGh^T h_{t-1} (row-shifted TN kernel), db_hh = column sums of dGh.
// dGi is the upstream gradient of the hoisted input projection, whose backward is tggcn_linear_bwd.
#include "common.cuh"
#include "gemm.h"
#include "bigru.h"
#include "backward.cuh"

using namespace tg;

extern "C" {

// Single-group frame-level BiGRU recurrence (forward), optionally saving r, z, n, hn for the backward.
// gi (B,T,E,2,3D) = W_ih x + b_ih for both directions; hfr (B,T,E,2D) out; gates (B,T,E,2,4D) out or NULL;
// sync: 16 bytes of device scratch for the grid barrier.
int tggcn_bigru_fwd(const float* gi, const float* whh_f, const float* whh_b, const float* bhh_f, const float* bhh_b, float* hfr,
                    float* gates, void* sync, int B, int T, int E, int D, int persistent, void* stream) {
    TG_REQUIRE(gi && whh_f && whh_b && bhh_f && bhh_b && hfr && sync, "bigru_fwd: null pointer");
    BiGruParams P;
    memset(&P, 0, sizeof(P));
    P.ngroups = 1; P.B = B; P.T = T; P.D = D;
    P.g[0].gi = gi; P.g[0].hfr = hfr; P.g[0].gates = gates;
    P.g[0].whh[0] = whh_f; P.g[0].whh[1] = whh_b; P.g[0].bhh[0] = bhh_f; P.g[0].bhh[1] = bhh_b;
    P.g[0].E = E; P.g[0].rows = B * E;
    P.sync.counter = (unsigned int*)sync; P.sync.error = (unsigned int*)sync + 1;
    TG_CUDA_OK(cudaMemsetAsync(sync, 0, 2 * sizeof(unsigned int), (cudaStream_t)stream));
    return launch_bigru(P, persistent, (cudaStream_t)stream);
}

// scratch floats needed by tggcn_bigru_bwd
size_t tggcn_bigru_bwd_scratch_floats(int B, int T, int E, int D) {
    (void)T;
    const size_t rows = (size_t)B * E;
    return 2 * (size_t)D * 3 * D      // W_hh^T for both directions
           + 2 * rows * D             // z (.) dh of the previous reverse step
           + 64;                      // grid-barrier counters
}

// dHFR (B,T,E,2D): gradient w.r.t. the BiGRU outputs.  Outputs: dGI (B,T,E,2,3D) gradient w.r.t. gi,
// dGH (B,T,E,2,3D) scratch (gradient w.r.t. W_hh h + b_hh), dWhh_* (3D,D), dbhh_* (3D).
int tggcn_bigru_bwd(const float* dhfr, const float* hfr, const float* gates, const float* whh_f, const float* whh_b, float* dgi,
                    float* dgh, float* dwhh_f, float* dwhh_b, float* dbhh_f, float* dbhh_b, float* scratch, int B, int T, int E, int D,
                    int gemm_path, void* stream_) {
    (void)gemm_path;
    TG_REQUIRE(dhfr && hfr && gates && whh_f && whh_b && dgi && dgh && scratch, "bigru_bwd: null pointer");
    TG_REQUIRE(D % 16 == 0, "bigru_bwd: hidden_size must be a multiple of 16");
    cudaStream_t stream = (cudaStream_t)stream_;
    const int rows = B * E;
    float* whhT[2] = {scratch, scratch + (size_t)D * 3 * D};
    float* direct = scratch + 2 * (size_t)D * 3 * D;
    unsigned int* sync = (unsigned int*)(direct + 2 * (size_t)rows * D);
    if (int rc = launch_transpose(whh_f, D, whhT[0], 3 * D, 3 * D, D, stream)) return rc;     // (3D,D) -> (D,3D)
    if (int rc = launch_transpose(whh_b, D, whhT[1], 3 * D, 3 * D, D, stream)) return rc;
    TG_CUDA_OK(cudaMemsetAsync(sync, 0, 16, stream));
    BiGruBwdParams P;
    memset(&P, 0, sizeof(P));
    P.ngroups = 1; P.B = B; P.T = T; P.D = D;
    P.g[0].dhfr = dhfr; P.g[0].hfr = hfr; P.g[0].gates = gates; P.g[0].whhT[0] = whhT[0]; P.g[0].whhT[1] = whhT[1];
    P.g[0].dgi = dgi; P.g[0].dgh = dgh; P.g[0].direct = direct; P.g[0].E = E; P.g[0].rows = rows;
    P.sync.counter = sync; P.sync.error = sync + 1;
    if (int rc = launch_bigru_bwd(P, 1, stream)) return rc;
    // weight gradients over all rows at once; h_{t-1} (fwd) / h_{t+1} (bwd) through the row shift of the TN kernel
    const int M = B * T * E;
    for (int dir = 0; dir < 2; ++dir) {
        float* dw = dir == 0 ? dwhh_f : dwhh_b;
        float* db = dir == 0 ? dbhh_f : dbhh_b;
        if (dw != nullptr)
            if (int rc = launch_gemm_tn(dgh + (size_t)dir * 3 * D, 6 * D, nullptr, 0, hfr + (size_t)dir * D, 2 * D, dw, D, M, 3 * D, D,
                                        dir == 0 ? -E : E, T * E, 0, stream))
                return rc;
        if (db != nullptr)
            if (int rc = launch_colsum(dgh + (size_t)dir * 3 * D, 6 * D, nullptr, 0, db, M, 3 * D, 0, stream)) return rc;
    }
    return 0;
}

}  // extern "C"
