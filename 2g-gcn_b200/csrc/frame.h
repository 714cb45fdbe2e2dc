// Parameter blocks of the frame-level graph kernels (frame.cu).
#pragma once
#include "common.cuh"

namespace tg {

struct FrameMsgParams {
    int B, T, H, O, D, hh;
    int mean_pool;          // message_aggregation 'mp': uniform weights over the valid senders instead of attention
    int att_noscale;        // attention_style 'v2': plain dot-product logits
    int update_strategy;    // 0 'ind', 1 'sah' (object gates = the single human's), 2 'coh' (hard object gate x the human's)
    const float* dist[3];   // distance-based attention: hh (B,T,H,H), ho (B,T,H,O), oo (B,T,O,O); each may be null
    // discrete_networks_num_layers == 2: the frame kernel writes the gate MLP inputs instead of sampling; the hidden layer is a
    // projection, launch_gate_sample finishes (Linear(D, 1) + sigmoid + sampling)
    float* gate_in_h; int gin_h;    // (B,T,H,gin_h) [x, h, m_hh?, m_oh, m_gh?, time?] or null
    float* gate_in_o; int gin_o;    // (B,T,O,gin_o) [x, h, m_ho, m_oo, m_go, time?]
    const float* gate_hid_h;        // (B,T,H,D) ReLU(W1 gate_in + b1): read by launch_gate_sample only
    const float* gate_hid_o;
    int tl;                 // add_segment_length: one more block at the end of every xx row (written by launch_segment_length)
    int gh;                 // message_geometry_to_human: block m_gh after m_oh in the humans' xx rows and gate inputs
    const float* msg_gh;    // (B,T,1,D) ReLU(W_gh s_g + b), or null
    int straight_through;   // discrete_optimization_strategy 'st': soft = sigmoid probability, hard = (p > thr), no noise
    int time_position;      // 0 off, 1 's': time block appended to the xx rows, 2 'u': appended to the gate inputs
    const float* time_emb;  // (B*T, D) time-position features, or null
    float thr;
    const float* s_h;       // (B,T,H,2D) [x | h]
    const float* s_o;       // (B,T,O,2D)
    const float* msg_hh;    // (B,T,H,D) ReLU(W_hh_msg s_h + b)      (unused when !hh)
    const float* msg_ho;    // (B,T,H,D)
    const float* msg_oh;    // (B,T,O,D)
    const float* msg_oo;    // (B,T,O,D)
    const float* msg_go;    // (B,T,1,D)
    const float* om;        // (B,O)
    const float* w_uh; const float* b_uh;   // update_human_segment_mlp.0  (1, 2D + nkh*D)
    const float* w_uo; const float* b_uo;   // update_object_segment_mlp.0 (1, 5D)
    const float* noise;     // (T*n_sampled, B, 2)
    const float* human_seg; // (B,T,H) or null
    const float* object_seg;// (B,T,O) or null
    float* xx_h;            // (B,T,H,(1+nkh)D)
    float* xx_o;            // (B,T,O,4D)
    float* y_hs; float* y_hss; float* y_os; float* y_oss;
    float* att_frame;       // (B,H,T,O) or null
    // saved for the backward (null in inference)
    float* alpha_save;      // (B*T, H*H + H*O + O*H + O*O): a_hh | a_oh | a_ho | a_oo, row-major [receiver][sender]
    float* pgate_save;      // (B*T, H+O): sigmoid probability of every sampled gate
};

struct HeadsParams {
    int B, T, E, NE, e_off, D, C;
    int cat;                // cat_level_states: the two segment-level heads read [hx | hfr] with (C,4D) weights
    const float* hfr;       // (B,T,E,2D) frame-level BiGRU outputs
    const float* hx;        // (B,T,E,2D) segment-level states (gathered through reidx)
    const int* reidx;       // (B,T,NE)
    const float* w[4]; const float* b[4];   // frame_rec, frame_pred, seg_rec, seg_pred
    float* out[4];          // (B,C,T,E)
};

int launch_frame_messages(const FrameMsgParams& P, cudaStream_t stream);
// time-position features: 'e' ReLU(w * (t+1)/steps[b] + bias), 'p' [sin((t+1)/freq_i), cos((t+1)/freq_i)]  (models.py:936-952, :1777-1794)
int launch_time_embed(const float* steps, const float* w, const float* bias, const float* freq, float* out, int B, int T, int D,
                      int periodic, cudaStream_t stream);
// add_segment_length: len[(b,t), e] from the hard gates (models.py:954-979), then its embedding ('e': ReLU(w * len + bias),
// 'p': periodic) into the last D columns of every xx_h / xx_o row (row strides ldh / ldo)
int launch_segment_length(const float* y_hs, const float* y_os, const float* steps, const float* w, const float* bias,
                          const float* freq, float* len, float* xx_h, int ldh, float* xx_o, int ldo, int B, int T, int H, int O,
                          int D, int periodic, cudaStream_t stream);
// Second layer of the two-layer gate MLPs + sampling: same FrameMsgParams (w_uh / w_uo = the (1, D) weights of layer 2)
int launch_gate_sample(const FrameMsgParams& P, cudaStream_t stream);
int launch_gate_post(float* y_hs, const float* y_hss, float* y_os, const float* y_oss, int* reidx, int B, int T, int H,
                     int O, int filter, float thr, cudaStream_t stream);
int launch_heads(const HeadsParams& P, cudaStream_t stream);
int launch_geo_gcn(const float* x_human, const void* const* w, float* out, float* bn_running_mean,
                   float* bn_running_var, int64_t* bn_num_batches, float* stats_ws, int B, int T, int H, int V, int Fh,
                   int bn_train, cudaStream_t stream);

}  // namespace tg
