// Parameter blocks and launchers of the hand-written backward kernels (backward.cu, geo_gcn_bwd.cu).
#pragma once
#include "common.cuh"

namespace tg {

// ---- label heads: Linear(2D->C) + LogSoftmax, gathered through the reorder index (models.py:909-917) -------------
struct HeadsBwdParams {
    int B, T, E, NE, e_off, D, C;
    int cat;                 // cat_level_states: the segment heads read [hx | hfr], weights (C,4D)
    const float* hfr;        // (B,T,E,2D) frame-level BiGRU outputs (input of the two frame heads)
    const float* hx;         // (B,T,E,2D) segment states (input of the two segment heads, through reidx)
    const int* reidx;        // (B,T,NE)
    const float* w[4];       // frame_rec, frame_pred, seg_rec, seg_pred weights (C,2D)
    const float* bias[4];
    const float* dlogp[4];   // upstream gradients (B,C,T,E); null = no gradient for that head
    float* dw[4];            // (C,2D)  accumulated with atomics: zero before the launch
    float* db[4];            // (C)
    float* dhfr;             // (B,T,E,2D) accumulated with atomics
    float* dhx;              // (B,T,E,2D) accumulated with atomics
};
int launch_heads_bwd(const HeadsBwdParams& P, cudaStream_t stream);

// ---- frame-level BiGRU, backward through time (recurrent_bwd.cu) ---------------------------------------------------
struct BiGruBwdGroup {
    const float* dhfr;       // (B,T,E,2D) gradient w.r.t. the BiGRU outputs
    const float* hfr;        // (B,T,E,2D) forward outputs
    const float* gates;      // (B,T,E,2,4D) r, z, n, hn saved by the forward
    const float* whhT[2];    // (D,3D) W_hh^T per direction
    float* dgi;              // (B,T,E,2,3D) out: gradient w.r.t. W_ih x + b_ih
    float* dgh;              // (B,T,E,2,3D) out: gradient w.r.t. W_hh h + b_hh
    float* direct;           // (2,rows,D) scratch: z (.) dh of the previous reverse step
    int E, rows;
    int cfg, n_rb, n_ub, tile_begin;     // tiling (filled by the launcher)
};
struct BiGruBwdParams {
    BiGruBwdGroup g[3];
    int ngroups, B, T, D;
    int NG, total_tiles;                 // filled by the launcher
    GridSync sync;
};
int launch_bigru_bwd(BiGruBwdParams& P, int persistent, cudaStream_t stream);

// ---- segment-level recurrent graph, backward through time (models.py:785-880; recurrent_bwd.cu) ----------------------
struct SegBwdParams {
    int B, T, H, O, D, hh, nk_h;
    int mean_pool;                                 // uniform sender weights: no gradient through attention logits
    int att_noscale;                               // attention_style 'v2': plain dot-product logits
    int dist_kind[4];                              // message kind (hh, oh, ho, oo) used distance-based weights: no logit gradient either
    const float* hx_h; const float* hx_o;          // forward states (B,T,E,2D)
    const float* sgates_h; const float* sgates_o;  // (B,T,E,2,4D) r, z, n, hn
    const float* u_h; const float* u_o;            // hard gates (B,T,E)
    const float* om;                               // (B,O)
    const float* dhx_h; const float* dhx_o;        // upstream gradient of the states (B,T,E,2D)
    const float* smsg[4]; const float* salpha[4];  // saved messages / attention weights (kinds hh, oh, ho, oo)
    // transposed weights (rows = output of the backward product)
    const float* wihT_h[2]; const float* wihT_o[2];   // (nk*D, 3D): segment-message columns of W_ih
    const float* whhT_h[2]; const float* whhT_o[2];   // (D, 3D)
    const float* wmT_h; const float* wmT_o;           // (D, nks*D): [W_hh_msg^T | W_ho_msg^T], [W_oh_msg^T | W_oo_msg^T]
    // outputs kept for the weight-gradient GEMMs
    float* dgs_h; float* dgs_o;                    // (B,T,E,2,3D) gradient of the hoisted pre-activations (= of W_ih x + b_ih)
    float* dghs_h; float* dghs_o;                  // (B,T,E,2,3D) gradient of W_hh h + b_hh
    float* du_h; float* du_o;                      // (B,T,E) gradient of the hard gates (atomics; zero before the launch)
    float* dpre_all[4];                            // per kind [dir][b][t][sender][D]: message MLP pre-activation gradients
    // per-step scratch (one step, both directions)
    float* direct_h; float* direct_o;              // [2][rows][D]: (1-u) dH + u z dH of the previous reverse step
    float* dmg_h; float* dmg_o;                    // [2][rows][nk*D] gradient of the aggregated messages
    float* dpre_h; float* dpre_o;                  // [2][rows_sender][nks*D]: [hh | ho] for human senders, [oh | oo] for objects
    float* lgr[4]; float* lgs[4];                  // per kind [2][rows][D]: attention-logit terms for receivers / senders
    // tiling (filled by the launcher)
    int cfg_h, cfg_o, nubA_h, nubA_o, tilesA_h, tilesA_dir, nubC, tilesC_h, tilesC_dir;
    GridSync sync;
};
int launch_segment_bwd(SegBwdParams& P, int persistent, cudaStream_t stream);

// ---- frame-level graph: attention, aggregation, gates (models.py:664-749, :1004-1533; distributions.py:4-36) ---------
struct FrameBwdParams {
    int B, T, H, O, D, hh, filter;
    int mean_pool;                                 // uniform sender weights: no gradient through attention logits
    int att_noscale;                               // attention_style 'v2': plain dot-product logits
    int update_strategy;                           // 0 'ind', 1 'sah', 2 'coh' (tggcn_dims.update_strategy)
    int dist_kind[4];                              // message kind (hh, oh, ho, oo) used distance-based weights: no logit gradient
    int tl;                                        // add_segment_length: one more block at the end of every xx row
    // discrete_networks_num_layers == 2: launch_gate_bwd turns the gate gradients into d hidden (and the layer-2 weight gradients),
    // two GEMMs turn d hidden into the gradient of the gate INPUTS, which launch_frame_bwd then adds instead of dlogit x weight
    int gate_layers;
    const float* hid_h; const float* hid_o;        // (B,T,E,D) forward hidden layers (post-ReLU)
    const float* w2_h; const float* w2_o;          // (D) layer-2 weights
    float* dhid_h; float* dhid_o;                  // (B,T,E,D) out of launch_gate_bwd (null: that entity type is not sampled)
    float* dw2_h; float* db2_h; float* dw2_o; float* db2_o;     // accumulated with atomics: zero before the launch
    const float* dgin_h; int gin_h;                // (B,T,H,gin_h) gradient of the gate inputs, read by launch_frame_bwd, or null
    const float* dgin_o; int gin_o;
    int gh;                                        // message_geometry_to_human: block m_gh after m_oh in the humans' rows
    const float* msg_gh;                           // (B,T,1,D) forward message, or null
    float* dmsg_gh;                                // (B,T,1,D) out: its gradient
    int straight_through;                          // discrete_optimization_strategy 'st': y = p, identity gradient of the hard gate
    int time_position;                             // 0 off, 1 's' (time block in the xx rows), 2 'u' (in the gate inputs)
    const float* time_emb;                         // (B*T, D) forward time-position features, or null
    float* dtime;                                  // (B*T, D) out: their gradient (null for the periodic encoding: no parameters)
    float thr;
    const float* s_h; const float* s_o;            // (B,T,E,2D)
    const float* msg_hh; const float* msg_ho; const float* msg_oh; const float* msg_oo; const float* msg_go;
    const float* om;
    const float* w_uh; const float* w_uo;          // gate weights
    const float* alpha; const float* pgate;        // saved by the forward
    const float* y_hss; const float* y_oss;        // soft gates (forward outputs)
    const float* human_seg; const float* object_seg;   // given segmentations (then no gate gradient)
    const float* xx_h; const float* xx_o;          // forward segment-level inputs [h, m..] (gate inputs are read from here)
    const float* dxx_h; const float* dxx_o;        // gradient of the segment-level inputs
    const float* du_h; const float* du_o;          // gradient of the hard gates from the segment loop
    const float* dy_hs; const float* dy_os;        // direct upstream gradient of the hard-gate outputs (may be null)
    const float* dy_hss; const float* dy_oss;      // direct upstream gradient of the soft-gate outputs (may be null)
    float* ds_h; float* ds_o;                      // (B,T,E,2D) written
    float* dmsg_hh; float* dmsg_ho; float* dmsg_oh; float* dmsg_oo; float* dmsg_go;   // (B,T,E,D) written
    float* dw_uh; float* db_uh; float* dw_uo; float* db_uo;   // accumulated with atomics: zero before the launch
};
int launch_frame_bwd(const FrameBwdParams& P, cudaStream_t stream);
int launch_gate_bwd(const FrameBwdParams& P, cudaStream_t stream);
// add_segment_length backward (models.py:763-779, :954-979): from the gradient of the length blocks of the xx rows — the gradient
// of the hard gates (added into du_h / du_o, reverse scan over the frames) and, for the embedding encoding, of segment_length_mlp
struct SegLenBwdParams {
    int B, T, H, O, D, periodic;
    const float* dxx_h; const float* xx_h; int ldh;   // gradient / forward rows of the humans (length block = last D columns)
    const float* dxx_o; const float* xx_o; int ldo;
    const float* len;                                  // (B,T,H+O) forward segment lengths
    const float* y_hs; const float* y_os;              // hard gates
    const float* steps; const float* w; const float* freq;
    float* du_h; float* du_o;                          // (B,T,E) accumulated
    float* dw; float* db;                              // segment_length_mlp gradients (null: periodic / not on the gradient path)
};
int launch_segment_length_bwd(const SegLenBwdParams& P, cudaStream_t stream);
// time_position_mlp (Linear(1, D) + ReLU of (t+1)/steps[b]): dw[k] = sum_n dtime[n,k] [emb[n,k] > 0] tau_n, db[k] likewise without tau
int launch_time_embed_bwd(const float* dtime, const float* emb, const float* steps, float* dw, float* db, int B, int T, int D,
                          cudaStream_t stream);

// ---- geometry GCN (models_gcn.py:30-100) -----------------------------------------------------------------------------
struct GcnBwdParams {
    const float* xh; int B, T, H, V, Fh;
    const float* mean; const float* var;           // the statistics the forward normalised with
    const float* gamma; const float* beta;
    const float* w1; const float* b1; const float* w3; const float* b3;
    const float* ws1; const float* bs1; const float* ws2; const float* bs2; const float* wg;
    const float* dout;                             // (B,128,V,T)
    float* dwg; float* dws1; float* dbs1; float* dws2; float* dbs2; float* dw3; float* db3; float* dw1; float* db1;   // atomics
    float* dxn;                                    // (B*T, V, 4) gradient of the normalised input (BN backward is a second pass)
};
int launch_geo_gcn_bwd(const GcnBwdParams& P, cudaStream_t stream);
int launch_geo_bn_bwd(const float* xh, const float* dxn, const float* mean, const float* var, const float* gamma, float* dgamma,
                      float* dbeta, int B, int T, int H, int V, int Fh, cudaStream_t stream);

}  // namespace tg
