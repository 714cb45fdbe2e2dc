// Backward through time of the frame-level bidirectional GRUs (autograd of nn.GRU as called at
// vhoi/models.py:997-1000), plus a single-group forward entry that also saves the gate values.
//
// Per step (reverse order of the forward recurrence of each direction):
//   dh      = dHFR[t] + carry
//   GRU cell backward (elementwise, saved r, z, n, hn):      dGi = [d a_r, d a_z, d a_n]   (gradient of W_ih x + b_ih)
//                                                           dGh = [d a_r, d a_z, r (.) d a_n] (gradient of W_hh h + b_hh)
//   carry   = z (.) dh + dGh W_hh                            (projection kernel on W_hh^T)
// After the loop the weight gradients are two GEMMs over all (video, t, entity) rows:
//   dW_hh = dGh^T h_{t-1} (row-shifted TN kernel), db_hh = column sums of dGh.
// dGi is the upstream gradient of the hoisted input projection, whose backward is tggcn_linear_bwd.
#include "common.cuh"
#include "gemm.h"
#include "bigru.h"

namespace tg {

// one thread per (row, unit, dir)
__global__ void __launch_bounds__(256) gru_cell_bwd_kernel(const float* __restrict__ dhfr, const float* __restrict__ hfr,
                                                           const float* __restrict__ gates, const float* __restrict__ carry,
                                                           float* __restrict__ dgi, float* __restrict__ dgh, float* __restrict__ carry_out,
                                                           int B, int T, int E, int D, int s) {
    const int rows = B * E;
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= 2L * rows * D) return;
    const int unit = (int)(idx % D);
    const int r = (int)((idx / D) % rows);
    const int dir = (int)(idx / ((long)D * rows));
    const int t = dir == 0 ? T - 1 - s : s;            // reverse of the forward recurrence order
    const int tprev = dir == 0 ? t - 1 : t + 1;        // the step whose state fed this one
    const bool has_prev = tprev >= 0 && tprev < T;
    const int b = r / E, e = r - b * E;
    const size_t fe = (size_t)(b * T + t) * E + e;
    const float dh = dhfr[fe * 2 * D + dir * D + unit] + (s > 0 ? carry[((size_t)dir * rows + r) * D + unit] : 0.0f);
    const float* g = gates + (fe * 2 + dir) * 4 * D + unit;
    const float rr = g[0], z = g[D], n = g[2 * D], hn = g[3 * D];
    const float hprev = has_prev ? hfr[((size_t)(b * T + tprev) * E + e) * 2 * D + dir * D + unit] : 0.0f;
    // h = n + z (hprev - n)
    const float dn = dh * (1.0f - z);
    const float dz = dh * (hprev - n);
    const float dan = dn * (1.0f - n * n);
    const float daz = dz * z * (1.0f - z);
    const float dr = dan * hn;
    const float dar = dr * rr * (1.0f - rr);
    float* gi = dgi + (fe * 2 + dir) * 3 * D + unit;
    gi[0] = dar; gi[D] = daz; gi[2 * D] = dan;
    float* gh = dgh + (fe * 2 + dir) * 3 * D + unit;
    gh[0] = dar; gh[D] = daz; gh[2 * D] = dan * rr;
    carry_out[((size_t)dir * rows + r) * D + unit] = dh * z;       // direct path; the W_hh path is added by the GEMM
}

// carry rows of step s for direction dir live at dgh rows (b, t, e): gather them into a dense (rows, 3D) matrix view
// is not needed: the projection kernel reads A with a leading dimension, but rows of one step are strided by T*E*6D per video.
// We therefore copy the step's dGh rows into a dense staging buffer.
__global__ void __launch_bounds__(256) gather_step_kernel(const float* __restrict__ dgh, float* __restrict__ dense, int B, int T, int E,
                                                          int D, int s) {
    const int rows = B * E;
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= 2L * rows * 3 * D) return;
    const int c = (int)(idx % (3 * D));
    const int r = (int)((idx / (3 * D)) % rows);
    const int dir = (int)(idx / ((long)3 * D * rows));
    const int t = dir == 0 ? T - 1 - s : s;
    const int b = r / E, e = r - b * E;
    dense[((size_t)dir * rows + r) * 3 * D + c] = dgh[(((size_t)(b * T + t) * E + e) * 2 + dir) * 3 * D + c];
}

}  // namespace tg

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
    return launch_bigru(P, persistent, (cudaStream_t)stream);
}

// scratch floats needed by tggcn_bigru_bwd
size_t tggcn_bigru_bwd_scratch_floats(int B, int T, int E, int D) {
    (void)T;
    const size_t rows = (size_t)B * E;
    return 2 * (size_t)D * 3 * D      // W_hh^T for both directions
           + 2 * 2 * rows * D         // carry ping-pong
           + 2 * rows * 3 * D;        // dense dGh of the current step
}

// dHFR (B,T,E,2D): gradient w.r.t. the BiGRU outputs.  Outputs: dGI (B,T,E,2,3D) gradient w.r.t. gi,
// dGH (B,T,E,2,3D) scratch (gradient w.r.t. W_hh h + b_hh), dWhh_* (3D,D), dbhh_* (3D).
int tggcn_bigru_bwd(const float* dhfr, const float* hfr, const float* gates, const float* whh_f, const float* whh_b, float* dgi,
                    float* dgh, float* dwhh_f, float* dwhh_b, float* dbhh_f, float* dbhh_b, float* scratch, int B, int T, int E, int D,
                    int gemm_path, void* stream_) {
    TG_REQUIRE(dhfr && hfr && gates && whh_f && whh_b && dgi && dgh && scratch, "bigru_bwd: null pointer");
    TG_REQUIRE(D % 16 == 0, "bigru_bwd: hidden_size must be a multiple of 16");
    cudaStream_t stream = (cudaStream_t)stream_;
    const int rows = B * E;
    float* whhT[2] = {scratch, scratch + (size_t)D * 3 * D};
    float* carry[2] = {scratch + 2 * (size_t)D * 3 * D, scratch + 2 * (size_t)D * 3 * D + 2 * (size_t)rows * D};
    float* dense = carry[1] + 2 * (size_t)rows * D;
    if (int rc = launch_transpose(whh_f, D, whhT[0], 3 * D, 3 * D, D, stream)) return rc;     // (3D,D) -> (D,3D)
    if (int rc = launch_transpose(whh_b, D, whhT[1], 3 * D, 3 * D, D, stream)) return rc;
    const long cell_threads = 2L * rows * D, gather_threads = 2L * rows * 3 * D;
    for (int s = 0; s < T; ++s) {
        float* cin = carry[s & 1];
        float* cout = carry[(s + 1) & 1];
        gru_cell_bwd_kernel<<<(unsigned)((cell_threads + 255) / 256), 256, 0, stream>>>(dhfr, hfr, gates, cin, dgi, dgh, cout, B, T, E, D, s);
        TG_LAUNCH_OK();
        if (s + 1 < T) {
            gather_step_kernel<<<(unsigned)((gather_threads + 255) / 256), 256, 0, stream>>>(dgh, dense, B, T, E, D, s);
            TG_LAUNCH_OK();
            GemmGroup g;
            g.count = 0;
            for (int dir = 0; dir < 2; ++dir) {
                // carry_out[dir] (rows, D) += dGh_step[dir] (rows, 3D) * W_hh[dir] (3D, D)
                gemm_add(g, dense + (size_t)dir * rows * 3 * D, 3 * D, whhT[dir], 3 * D, nullptr, cout + (size_t)dir * rows * D, D, rows, D,
                         3 * D, 0);
                g.p[dir].beta = 1;
            }
            if (int rc = launch_gemm(g, gemm_path, stream)) return rc;
        }
    }
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
