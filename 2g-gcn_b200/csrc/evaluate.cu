// Evaluation post-processing on the device: what predict.py does on the host with numpy after the
// forward — up-sampling of the down-sampled log-probabilities back to the target's frame rate (predict.py:64-70:
// repeat_interleave along time + match_shape :95-122), the per-frame argmax (process_output :186-202) and the segmental
// F1@k of pyrutils/metrics.py:7-81 with the (B,T,E) -> (B*E, T) row convention of predict.py:236-240.
#include "common.cuh"

namespace tg {

// labels[b][tt][e] = argmax_c logp[b][c][min(tt / ds, T-1)][e]   (first maximum, as np.argmax)
__global__ void __launch_bounds__(256) upsample_argmax_kernel(const float* __restrict__ logp, long long* __restrict__ labels, int B,
                                                             int C, int T, int E, int Tt, int ds) {
    const long long n = (long long)B * Tt * E;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int e = (int)(i % E);
        const int tt = (int)((i / E) % Tt);
        const int b = (int)(i / ((long long)E * Tt));
        int t = tt / ds;
        if (t > T - 1) t = T - 1;                         // match_shape: the last step is repeated when the target is longer
        const float* p = logp + ((size_t)b * C * T + t) * E + e;
        float best = p[0];
        int arg = 0;
        for (int c = 1; c < C; ++c) {
            const float v = p[(size_t)c * T * E];
            if (v > best) { best = v; arg = c; }
        }
        labels[i] = arg;
    }
}

// One thread per (video, entity) row.  seg: per-row scratch of 3 * Tt ints for the target segments (start, end, id) and
// 3 * Tt ints for the predicted ones; used: Tt bytes per row.  f1[k * rows + row] (double), valid[row] = row has frames.
__global__ void __launch_bounds__(64) f1_at_k_kernel(const long long* __restrict__ target, const long long* __restrict__ pred,
                                                    int B, int Tt, int E, int num_classes, const double* __restrict__ overlaps,
                                                    int n_overlaps, long long ignore, int* __restrict__ seg,
                                                    unsigned char* __restrict__ used, double* __restrict__ f1,
                                                    int* __restrict__ valid) {
    const int rows = B * E;
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= rows) return;
    const int b = row / E, e = row - b * E;
    int* ts = seg + (size_t)row * 6 * Tt;                 // target: start, end, id ; prediction: start, end, id
    int* te = ts + Tt; int* ti = te + Tt;
    int* os = ti + Tt; int* oe = os + Tt; int* oi = oe + Tt;
    unsigned char* u = used + (size_t)row * Tt;
    // run-length encoding of both label rows over the frames whose target is not the padding value
    int nt = 0, no = 0, p = 0;
    long long prev_t = 0, prev_o = 0;
    for (int tt = 0; tt < Tt; ++tt) {
        const size_t idx = ((size_t)b * Tt + tt) * E + e;
        const long long yt = target[idx];
        if (yt == ignore) continue;
        const long long yo = pred[idx];
        if (p == 0 || yt != prev_t) { if (nt) te[nt - 1] = p; ts[nt] = p; ti[nt] = (int)yt; ++nt; }
        if (p == 0 || yo != prev_o) { if (no) oe[no - 1] = p; os[no] = p; oi[no] = (int)yo; ++no; }
        prev_t = yt; prev_o = yo;
        ++p;
    }
    valid[row] = p > 0;
    if (p == 0) {
        for (int k = 0; k < n_overlaps; ++k) f1[(size_t)k * rows + row] = 0.0;
        return;
    }
    te[nt - 1] = p; oe[no - 1] = p;
    for (int k = 0; k < n_overlaps; ++k) {
        const double overlap = overlaps[k];
        for (int j = 0; j < nt; ++j) u[j] = 0;
        double tp = 0.0, fp = 0.0;
        for (int i = 0; i < no; ++i) {
            // IoU of the predicted segment against every target segment of the same class; first maximum (np.argmax)
            const int a = os[i], z = oe[i], id = oi[i];
            double best = 0.0;
            int arg = 0;
            for (int j = 0; j < nt; ++j) {
                const int inter = min(z, te[j]) - max(a, ts[j]);
                const int uni = max(z, te[j]) - min(a, ts[j]);
                const double iou = (ti[j] == id) ? (double)inter / (double)uni : 0.0;
                if (j == 0 || iou > best) { best = iou; arg = j; }
            }
            if (id >= num_classes) continue;
            if (best >= overlap && !u[arg]) { tp += 1.0; u[arg] = 1; }
            else fp += 1.0;
        }
        double nused = 0.0;
        for (int j = 0; j < nt; ++j) nused += u[j];
        const double fn = (double)nt - nused;
        const double prec = (tp + fp) != 0.0 ? tp / (tp + fp) : 0.0;
        const double rec = (tp + fn) != 0.0 ? tp / (tp + fn) : 0.0;
        f1[(size_t)k * rows + row] = (prec + rec) != 0.0 ? 2.0 * (prec * rec) / (prec + rec) : 0.0;
    }
}

}  // namespace tg

using namespace tg;

extern "C" {

int tggcn_upsample_argmax(const float* logp, int64_t* labels, int B, int C, int T, int E, int Tt, int downsampling, void* stream) {
    TG_REQUIRE(logp && labels, "upsample_argmax: null pointer");
    TG_REQUIRE(B > 0 && C > 0 && T > 0 && E > 0 && Tt > 0 && downsampling >= 1, "upsample_argmax: bad dimensions");
    const long long n = (long long)B * Tt * E;
    int grid = (int)((n + 255) / 256);
    if (grid > 8 * num_sms()) grid = 8 * num_sms();
    upsample_argmax_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(logp, (long long*)labels, B, C, T, E, Tt, downsampling);
    TG_LAUNCH_OK();
    return 0;
}

size_t tggcn_f1_at_k_scratch_bytes(int B, int Tt, int E) { return (size_t)B * E * Tt * (6 * sizeof(int) + 1) + 16; }

int tggcn_f1_at_k(const int64_t* target, const int64_t* pred, int B, int Tt, int E, int num_classes, const double* overlaps,
                  int n_overlaps, int64_t ignore_value, void* scratch, double* f1_rows, int32_t* valid_rows, void* stream) {
    TG_REQUIRE(target && pred && overlaps && scratch && f1_rows && valid_rows, "f1_at_k: null pointer");
    TG_REQUIRE(B > 0 && Tt > 0 && E > 0 && n_overlaps > 0, "f1_at_k: bad dimensions");
    const int rows = B * E;
    int* seg = (int*)scratch;
    unsigned char* used = (unsigned char*)(seg + (size_t)rows * 6 * Tt);
    f1_at_k_kernel<<<cdiv(rows, 64), 64, 0, (cudaStream_t)stream>>>((const long long*)target, (const long long*)pred, B, Tt, E,
                                                                   num_classes, overlaps, n_overlaps, (long long)ignore_value, seg,
                                                                   used, f1_rows, valid_rows);
    TG_LAUNCH_OK();
    return 0;
}

}  // extern "C"
