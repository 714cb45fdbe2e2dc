/*
 * tggcn_b200.h — C ABI of the B200-native 2G-GCN (TGGCN) hot path.
 *
 * The reference (tanqiu98/2G-GCN) is pure Python/PyTorch and has no FFI of its own; the boundary this
 * library replaces is the body of `TGGCN.forward` (vhoi/models.py:584-933) together with the building
 * blocks it calls (pyrutils/torch/models_gcn.py:6-100, pyrutils/torch/distributions.py:4-36).  A
 * maintainer binds it with a ctypes stub; our own host-side mirror of the model class
 * (2g-gcn_b200/model.py) is that stub plus the nn.Module parameter holders.
 *
 * Conventions
 *  - plain pointers and sizes only; every pointer is a DEVICE pointer unless named *_host;
 *  - all tensors are fp32, contiguous, row-major with the shapes given below (labels of the reference);
 *  - every entry returns 0 on success, non-zero on error (message via tggcn_last_error()); no entry
 *    allocates, synchronises the device, or keeps state between calls;
 *  - `stream` is a cudaStream_t passed as void* (0 = legacy default stream).
 */
#ifndef TGGCN_B200_H_
#define TGGCN_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TGGCN_ABI_VERSION 16

#if defined(__GNUC__)
#define TGGCN_API __attribute__((visibility("default")))
#else
#define TGGCN_API
#endif

/* Problem description: the constructor arguments of TGGCN.__init__ (vhoi/models.py:179-190) that change
 * the arithmetic, plus the batch geometry of one forward call (vhoi/models.py:624, :640). */
typedef struct tggcn_dims {
    int32_t B, T, H, O;          /* videos, padded frames, humans, object slots                         */
    int32_t V, D;                /* gcn_node, hidden_size                                                */
    int32_t Fh;                  /* x_human feature size = 2048 + 4V (models.py:631-639)                 */
    int32_t C_sub, C_aff;        /* num_classes; C_aff = 0 when num_affordances is None (models.py:560)  */
    int32_t hh;                  /* message_humans_to_human                                              */
    int32_t filter;              /* filter_discrete_updates (models.py:751-753)                          */
    int32_t bn_train;            /* BatchNorm1d uses batch statistics and updates running stats          */
    int32_t human_seg_given;     /* human_segmentation passed (models.py:697-698)                        */
    int32_t object_seg_given;    /* objects_segmentation passed (models.py:738-739)                      */
    int32_t inspect;             /* inspect_model: also emit attention weights (models.py:927-932)       */
    int32_t persistent;          /* 1 = persistent cooperative recurrent kernels, 0 = one launch per step */
    int32_t gemm_path;           /* projections: 0 = fp32 SIMT, 1 = tcgen05 3xTF32, 2 = tcgen05 where K%32==0 */
    float   thr;                 /* update_segment_threshold                                             */
    int32_t save_for_backward;   /* forward also stores what tggcn_backward needs (bigger workspace)     */
    int32_t cat_level_states;    /* segment-level heads read [segment state | frame-level state] (models.py:901-903): their
                                    weights are (C, 4D) instead of (C, 2D)                              */
    int32_t mean_pool;           /* message_aggregation 'mp': senders averaged with weight mask / max(#valid, 1) instead of
                                    the scaled-dot-product attention (models.py:1033-1036 and the other message functions) */
    int32_t recurrent_mode;      /* recurrent kernels (BiGRU + segment level): 0 = choose by rows per step, 1 = latency path
                                    (persistent mma.sync kernels, weights resident on chip where they fit), 2 = large-batch path
                                    (one tcgen05 + TMA step kernel per recurrent step, activation rows as the UMMA M operand)  */
    int32_t no_fp16_split;       /* 1 = never use the fp16 (hi, lo) operand split in the recurrent kernels (3xTF32 streaming
                                    tiles only); set by the host after a range violation was reported in the status words     */
    int32_t precision;           /* 0 = fp32-class products everywhere (3-term split MMAs; the parity configuration),
                                    1 = bf16 operands with fp32 accumulation in the projections and the large-batch recurrent
                                    kernels (BASELINE.json configs[2]); gate / softmax / loss arithmetic stays fp32           */
    int32_t att_noscale;         /* attention_style 'v2' / 'dot-product': logits <q, k> without the 1/sqrt(size) factor of 'v3'
                                    (compute_attention_weights, models.py:1740-1745)                                           */
    int32_t update_strategy;     /* object_segment_update_strategy (models.py:1523-1532): 0 = 'ind'; 1 = 'sah': with exactly one human
                                    and no objects_segmentation the object gates (hard and soft) ARE the human's, no object gate MLP,
                                    no noise drawn for objects; 2 = 'coh': hard object gate = own decision x the human's hard gate
                                    (one human, no local-maximum filter; otherwise identical to 'ind')                          */
    int32_t time_position;       /* add_time_position (models.py:656-662, :755-762): 0 = off; 1 = strategy 's': the time feature of frame
                                    (b, t) is appended to every segment-level input row [h, m.., time] (so W_ih of the segment cells
                                    has D more columns before the segment-message columns); 2 = strategy 'u': appended to the gate
                                    MLP inputs instead (update_*_segment_mlp weights have D more columns at the end)             */
    int32_t time_periodic;       /* positional_encoding_style: 0 = 'e': ReLU(time_position_mlp((t+1) / steps_per_example[b]))
                                    (models.py:259-260, :936-952); 1 = 'p': [sin((t+1)/w_i), cos((t+1)/w_i)] with the D/2
                                    frequencies w_i = 1e4^(i/(D/2-1)) passed in tggcn_io.time_freq (models.py:1777-1794)         */
    int32_t straight_through;    /* discrete_optimization_strategy 'st' (discrete_estimator, models.py:1620-1622; StraightThroughEstimator,
                                    distributions.py:39-53): soft gate = the sigmoid probability, hard gate = (p > thr) exactly, identity
                                    gradient; no Gumbel noise is read.  0 = 'gs' (Gumbel-sigmoid, distributions.py:4-36)                */
    int32_t geo_to_human;        /* message_geometry_to_human (models.py:690-695, :1432-1475): the humans' segment-level input rows and
                                    gate inputs carry one more block m_gh = ReLU(geometry_to_human_message_mlp([x_g | h_g])) after
                                    m_oh (single sender: weight 1, no mask)                                                     */
    int32_t segment_length;      /* add_segment_length (models.py:763-779, :954-979): every segment-level input row ends with the embedding
                                    (time_periodic selects segment_length_mlp or the periodic encoding) of its entity's segment length: at
                                    a frame with a non-zero hard gate the (normalised) time since the previous such frame, else 0; the
                                    hard gates receive a gradient through it.  Needs tggcn_io.steps_per_example                  */
    int32_t gate_layers;         /* discrete_networks_num_layers (models.py:532-547): 0 / 1 = Linear(in, 1) + sigmoid; 2 = Linear(in, D) +
                                    ReLU + Linear(D, 1) + sigmoid: update_*_segment_mlp.0 is then (D, in) and .2 is (1, D); 3 = one more
                                    Linear(D, D) + ReLU in between (.2 is (D, D), .4 is (1, D)).  From 2 on the gate inputs are
                                    materialised (TGGCN_BUF_GATE_IN_*) and the hidden layers are projection GEMMs                  */
} tggcn_dims;

/* Parameter table.  One device pointer per reference state_dict() entry, in this order
 * (keys: SURVEY.md §8b; vhoi/models.py:264-580).  Entries a configuration does not have are NULL. */
#define TGGCN_WEIGHT_LIST(X)                                                                            \
    X(GCN_W,        "geometry_embedding_gcn.weight")                                                   \
    X(GCN_BN_W,     "geometry_embedding_gcn.joint_embed.cnn.0.bn.weight")                              \
    X(GCN_BN_B,     "geometry_embedding_gcn.joint_embed.cnn.0.bn.bias")                                \
    X(GCN_BN_MEAN,  "geometry_embedding_gcn.joint_embed.cnn.0.bn.running_mean")                        \
    X(GCN_BN_VAR,   "geometry_embedding_gcn.joint_embed.cnn.0.bn.running_var")                         \
    X(GCN_C1_W,     "geometry_embedding_gcn.joint_embed.cnn.1.cnn.weight")                             \
    X(GCN_C1_B,     "geometry_embedding_gcn.joint_embed.cnn.1.cnn.bias")                               \
    X(GCN_C3_W,     "geometry_embedding_gcn.joint_embed.cnn.3.cnn.weight")                             \
    X(GCN_C3_B,     "geometry_embedding_gcn.joint_embed.cnn.3.cnn.bias")                               \
    X(GCN_S1_W,     "geometry_embedding_gcn.get_s.s1.cnn.weight")                                      \
    X(GCN_S1_B,     "geometry_embedding_gcn.get_s.s1.cnn.bias")                                        \
    X(GCN_S2_W,     "geometry_embedding_gcn.get_s.s2.cnn.weight")                                      \
    X(GCN_S2_B,     "geometry_embedding_gcn.get_s.s2.cnn.bias")                                        \
    X(GEO_MLP0_W,   "geometry_embedding_mlp.0.weight")                                                 \
    X(GEO_MLP0_B,   "geometry_embedding_mlp.0.bias")                                                   \
    X(GEO_MLP2_W,   "geometry_embedding_mlp.2.weight")                                                 \
    X(GEO_MLP2_B,   "geometry_embedding_mlp.2.bias")                                                   \
    X(HUM_EMB_W,    "human_embedding_mlp.0.weight")                                                    \
    X(HUM_EMB_B,    "human_embedding_mlp.0.bias")                                                      \
    X(OBJ_EMB_W,    "object_embedding_mlp.0.weight")                                                   \
    X(OBJ_EMB_B,    "object_embedding_mlp.0.bias")                                                     \
    X(GEO_RNN_WIH_F, "geometry_bd_rnn.weight_ih_l0")                                                   \
    X(GEO_RNN_WHH_F, "geometry_bd_rnn.weight_hh_l0")                                                   \
    X(GEO_RNN_BIH_F, "geometry_bd_rnn.bias_ih_l0")                                                     \
    X(GEO_RNN_BHH_F, "geometry_bd_rnn.bias_hh_l0")                                                     \
    X(GEO_RNN_WIH_B, "geometry_bd_rnn.weight_ih_l0_reverse")                                           \
    X(GEO_RNN_WHH_B, "geometry_bd_rnn.weight_hh_l0_reverse")                                           \
    X(GEO_RNN_BIH_B, "geometry_bd_rnn.bias_ih_l0_reverse")                                             \
    X(GEO_RNN_BHH_B, "geometry_bd_rnn.bias_hh_l0_reverse")                                             \
    X(HUM_RNN_WIH_F, "human_bd_rnn.weight_ih_l0")                                                      \
    X(HUM_RNN_WHH_F, "human_bd_rnn.weight_hh_l0")                                                      \
    X(HUM_RNN_BIH_F, "human_bd_rnn.bias_ih_l0")                                                        \
    X(HUM_RNN_BHH_F, "human_bd_rnn.bias_hh_l0")                                                        \
    X(HUM_RNN_WIH_B, "human_bd_rnn.weight_ih_l0_reverse")                                              \
    X(HUM_RNN_WHH_B, "human_bd_rnn.weight_hh_l0_reverse")                                              \
    X(HUM_RNN_BIH_B, "human_bd_rnn.bias_ih_l0_reverse")                                                \
    X(HUM_RNN_BHH_B, "human_bd_rnn.bias_hh_l0_reverse")                                                \
    X(OBJ_RNN_WIH_F, "object_bd_rnn.weight_ih_l0")                                                     \
    X(OBJ_RNN_WHH_F, "object_bd_rnn.weight_hh_l0")                                                     \
    X(OBJ_RNN_BIH_F, "object_bd_rnn.bias_ih_l0")                                                       \
    X(OBJ_RNN_BHH_F, "object_bd_rnn.bias_hh_l0")                                                       \
    X(OBJ_RNN_WIH_B, "object_bd_rnn.weight_ih_l0_reverse")                                             \
    X(OBJ_RNN_WHH_B, "object_bd_rnn.weight_hh_l0_reverse")                                             \
    X(OBJ_RNN_BIH_B, "object_bd_rnn.bias_ih_l0_reverse")                                               \
    X(OBJ_RNN_BHH_B, "object_bd_rnn.bias_hh_l0_reverse")                                               \
    X(GEO_BD_W,     "geometry_bd_embedding_mlp.0.weight")                                              \
    X(GEO_BD_B,     "geometry_bd_embedding_mlp.0.bias")                                                \
    X(HUM_BD_W,     "human_bd_embedding_mlp.0.weight")                                                 \
    X(HUM_BD_B,     "human_bd_embedding_mlp.0.bias")                                                   \
    X(OBJ_BD_W,     "object_bd_embedding_mlp.0.weight")                                                \
    X(OBJ_BD_B,     "object_bd_embedding_mlp.0.bias")                                                  \
    X(MSG_HH_W,     "humans_to_human_message_mlp.0.weight")                                            \
    X(MSG_HH_B,     "humans_to_human_message_mlp.0.bias")                                              \
    X(MSG_HO_W,     "human_to_object_message_mlp.0.weight")                                            \
    X(MSG_HO_B,     "human_to_object_message_mlp.0.bias")                                              \
    X(MSG_OH_W,     "objects_to_human_message_mlp.0.weight")                                           \
    X(MSG_OH_B,     "objects_to_human_message_mlp.0.bias")                                             \
    X(MSG_OO_W,     "objects_to_object_message_mlp.0.weight")                                          \
    X(MSG_OO_B,     "objects_to_object_message_mlp.0.bias")                                            \
    X(MSG_GO_W,     "geometry_to_object_message_mlp.0.weight")                                         \
    X(MSG_GO_B,     "geometry_to_object_message_mlp.0.bias")                                           \
    X(SMSG_HH_W,    "humans_to_human_segment_message_mlp.0.weight")                                    \
    X(SMSG_HH_B,    "humans_to_human_segment_message_mlp.0.bias")                                      \
    X(SMSG_HO_W,    "human_to_object_segment_message_mlp.0.weight")                                    \
    X(SMSG_HO_B,    "human_to_object_segment_message_mlp.0.bias")                                      \
    X(SMSG_OH_W,    "objects_to_human_segment_message_mlp.0.weight")                                   \
    X(SMSG_OH_B,    "objects_to_human_segment_message_mlp.0.bias")                                     \
    X(SMSG_OO_W,    "objects_to_object_segment_message_mlp.0.weight")                                  \
    X(SMSG_OO_B,    "objects_to_object_segment_message_mlp.0.bias")                                    \
    X(UPD_H_W,      "update_human_segment_mlp.0.weight")                                               \
    X(UPD_H_B,      "update_human_segment_mlp.0.bias")                                                 \
    X(UPD_O_W,      "update_object_segment_mlp.0.weight")                                              \
    X(UPD_O_B,      "update_object_segment_mlp.0.bias")                                                \
    X(HSEG_F_WIH,   "human_segment_rnn_fcell.weight_ih")                                               \
    X(HSEG_F_WHH,   "human_segment_rnn_fcell.weight_hh")                                               \
    X(HSEG_F_BIH,   "human_segment_rnn_fcell.bias_ih")                                                 \
    X(HSEG_F_BHH,   "human_segment_rnn_fcell.bias_hh")                                                 \
    X(HSEG_B_WIH,   "human_segment_rnn_bcell.weight_ih")                                               \
    X(HSEG_B_WHH,   "human_segment_rnn_bcell.weight_hh")                                               \
    X(HSEG_B_BIH,   "human_segment_rnn_bcell.bias_ih")                                                 \
    X(HSEG_B_BHH,   "human_segment_rnn_bcell.bias_hh")                                                 \
    X(OSEG_F_WIH,   "object_segment_rnn_fcell.weight_ih")                                              \
    X(OSEG_F_WHH,   "object_segment_rnn_fcell.weight_hh")                                              \
    X(OSEG_F_BIH,   "object_segment_rnn_fcell.bias_ih")                                                \
    X(OSEG_F_BHH,   "object_segment_rnn_fcell.bias_hh")                                                \
    X(OSEG_B_WIH,   "object_segment_rnn_bcell.weight_ih")                                              \
    X(OSEG_B_WHH,   "object_segment_rnn_bcell.weight_hh")                                              \
    X(OSEG_B_BIH,   "object_segment_rnn_bcell.bias_ih")                                                \
    X(OSEG_B_BHH,   "object_segment_rnn_bcell.bias_hh")                                                \
    X(HEAD_H_FREC_W, "human_frame_recognition_mlp.0.weight")                                           \
    X(HEAD_H_FREC_B, "human_frame_recognition_mlp.0.bias")                                             \
    X(HEAD_H_FPRED_W, "human_frame_prediction_mlp.0.weight")                                           \
    X(HEAD_H_FPRED_B, "human_frame_prediction_mlp.0.bias")                                             \
    X(HEAD_H_REC_W, "human_recognition_mlp.0.weight")                                                  \
    X(HEAD_H_REC_B, "human_recognition_mlp.0.bias")                                                    \
    X(HEAD_H_PRED_W, "human_prediction_mlp.0.weight")                                                  \
    X(HEAD_H_PRED_B, "human_prediction_mlp.0.bias")                                                    \
    X(HEAD_O_FREC_W, "object_frame_recognition_mlp.0.weight")                                          \
    X(HEAD_O_FREC_B, "object_frame_recognition_mlp.0.bias")                                            \
    X(HEAD_O_FPRED_W, "object_frame_prediction_mlp.0.weight")                                          \
    X(HEAD_O_FPRED_B, "object_frame_prediction_mlp.0.bias")                                            \
    X(HEAD_O_REC_W, "object_recognition_mlp.0.weight")                                                 \
    X(HEAD_O_REC_B, "object_recognition_mlp.0.bias")                                                   \
    X(HEAD_O_PRED_W, "object_prediction_mlp.0.weight")                                                 \
    X(HEAD_O_PRED_B, "object_prediction_mlp.0.bias")                                                   \
    X(TIME_W,       "time_position_mlp.0.weight")                                                      \
    X(TIME_B,       "time_position_mlp.0.bias")                                                        \
    X(MSG_GH_W,     "geometry_to_human_message_mlp.0.weight")                                          \
    X(MSG_GH_B,     "geometry_to_human_message_mlp.0.bias")                                            \
    X(LEN_W,        "segment_length_mlp.0.weight")                                                     \
    X(LEN_B,        "segment_length_mlp.0.bias")                                                       \
    X(UPD_H_W2,     "update_human_segment_mlp.2.weight")                                               \
    X(UPD_H_B2,     "update_human_segment_mlp.2.bias")                                                 \
    X(UPD_O_W2,     "update_object_segment_mlp.2.weight")                                              \
    X(UPD_O_B2,     "update_object_segment_mlp.2.bias")                                                \
    X(UPD_H_W4,     "update_human_segment_mlp.4.weight")                                               \
    X(UPD_H_B4,     "update_human_segment_mlp.4.bias")                                                 \
    X(UPD_O_W4,     "update_object_segment_mlp.4.weight")                                              \
    X(UPD_O_B4,     "update_object_segment_mlp.4.bias")

enum tggcn_weight_id {
#define TGGCN_X_ENUM(id, key) TGGCN_W_##id,
    TGGCN_WEIGHT_LIST(TGGCN_X_ENUM)
#undef TGGCN_X_ENUM
    TGGCN_W_COUNT
};

/* Inputs / outputs of one forward call: the keyword arguments of TGGCN.forward (models.py:584-586,
 * as passed by gcn_forward, vhoi/data_loading.py:1245-1279) and the tensors of its output list
 * (models.py:919-932). */
typedef struct tggcn_io {
    const float* x_human;        /* (B,T,H,Fh)                                                          */
    const float* x_objects;      /* (B,T,O,2048)                                                        */
    const float* objects_mask;   /* (B,O) in {0,1}                                                      */
    const float* human_seg;      /* (B,T,H) or NULL                                                     */
    const float* object_seg;     /* (B,T,O) or NULL                                                     */
    const float* noise;          /* (T*n_sampled, B, 2) Gumbel(0,1) draws in the reference's call order
                                    (t-major; sampled humans then sampled objects), or NULL if none     */
    float* y_hs;  float* y_hss;  /* (B,T,H) hard / soft human gates                                     */
    float* y_os;  float* y_oss;  /* (B,T,O) hard / soft object gates                                    */
    float* out_h[4];             /* frame_rec, frame_pred, seg_rec, seg_pred: (B,C_sub,T,H) log-probs   */
    float* out_o[4];             /* same for objects (B,C_aff,T,O); NULL when C_aff == 0                */
    float* att_frame;            /* (B,H,T,O) objects->human frame attention, NULL unless inspect       */
    float* att_seg_f;            /* (B,H,T,O) segment-level, forward direction, NULL unless inspect     */
    float* att_seg_b;            /* (B,H,T,O) segment-level, backward direction, NULL unless inspect    */
    float* bn_running_mean;      /* (4V) updated in place when bn_train                                  */
    float* bn_running_var;       /* (4V) updated in place when bn_train                                  */
    int64_t* bn_num_batches;     /* scalar, incremented when bn_train                                    */
    /* misc.make_attention_distance_based (vhoi/data_loading.py:1264-1276): entity distances; where a pointer is given the attention
     * weights of the message kinds it covers — frame level AND segment level — are softmax(1 / (d + 1e-7)) over the real senders
     * at a non-zero distance (compute_distance_based_attention_weights, models.py:1757-1775) instead of the dot-product attention;
     * no gradient flows through them.  Ignored under mean pooling.  Each may be NULL. */
    const float* dist_hh;        /* (B,T,H,H) humans -> human                                            */
    const float* dist_ho;        /* (B,T,H,O) objects -> human and human -> objects                      */
    const float* dist_oo;        /* (B,T,O,O) objects -> object                                          */
    const float* steps_per_example; /* (B) number of real frames per video, or NULL; required when dims.time_position != 0   */
    const float* time_freq;      /* (D/2) periods w_i of the periodic encoding (dims.time_periodic), else NULL               */
    uint32_t* status_host;       /* PINNED HOST memory, 8 words, or NULL.  When set, tggcn_forward / tggcn_backward end with an
                                    asynchronous copy of the status words (see tggcn_status_decode) into it, so the caller can
                                    test them after any later synchronisation point without an extra round trip.           */
} tggcn_io;

/* Named regions of the workspace, exposed so that tests can compare intermediates with the oracle. */
enum tggcn_buf_id {
    TGGCN_BUF_GCN_OUT = 0,   /* (B,128,V,T)          Geo_gcn output, models_gcn.py:30-37                 */
    TGGCN_BUF_GEO_HID,       /* (B*T,2048)           geometry_embedding_mlp hidden                       */
    TGGCN_BUF_S_H,           /* (B,T,H,2D)           [x_h | h_h]                                         */
    TGGCN_BUF_S_O,           /* (B,T,O,2D)                                                               */
    TGGCN_BUF_S_G,           /* (B,T,1,2D)                                                               */
    TGGCN_BUF_GI_H,          /* (B,T,H,2,3D)         BiGRU input pre-activations (fwd, bwd)              */
    TGGCN_BUF_GI_O,
    TGGCN_BUF_GI_G,
    TGGCN_BUF_HFR_H,         /* (B,T,H,2D)           BiGRU outputs                                       */
    TGGCN_BUF_HFR_O,
    TGGCN_BUF_HFR_G,
    TGGCN_BUF_MSG_HH,        /* (B,T,H,D)            per-sender frame messages                           */
    TGGCN_BUF_MSG_HO,        /* (B,T,H,D)                                                                */
    TGGCN_BUF_MSG_OH,        /* (B,T,O,D)                                                                */
    TGGCN_BUF_MSG_OO,        /* (B,T,O,D)                                                                */
    TGGCN_BUF_MSG_GO,        /* (B,T,1,D)                                                                */
    TGGCN_BUF_XX_H,          /* (B,T,H,3D or 2D [+D]) segment-level frame inputs, models.py:705 (+ time block, :761) */
    TGGCN_BUF_XX_O,          /* (B,T,O,4D [+D])      models.py:748 (+ time block, :762)                  */
    TGGCN_BUF_GS_H,          /* (B,T,H,2,3D)         hoisted segment-cell input pre-activations          */
    TGGCN_BUF_GS_O,          /* (B,T,O,2,3D)                                                             */
    TGGCN_BUF_HX_H,          /* (B,T,H,2D)           segment states [fwd | bwd], before reorder          */
    TGGCN_BUF_HX_O,          /* (B,T,O,2D)                                                               */
    TGGCN_BUF_REIDX,         /* (B,T,H+O) int32      reorder gather index, models.py:1567-1586           */
    TGGCN_BUF_SEG_SCRATCH,   /* per-step message scratch of the segment kernel                            */
    TGGCN_BUF_SYNC,          /* grid-barrier counters + error flag                                        */
    TGGCN_BUF_BIG,           /* large-batch recurrent path: 16-bit operand copies of the recurrent weights, state rings,
                                aggregated-message operand rows, per-step message scratch (empty on the latency path)  */
    /* saved for the backward (empty unless dims.save_for_backward) */
    TGGCN_BUF_GATES_H,       /* (B,T,H,2,4D)         BiGRU gates r, z, n, W_hn h + b_hn                           */
    TGGCN_BUF_GATES_O,
    TGGCN_BUF_GATES_G,
    TGGCN_BUF_ALPHA_F,       /* (B*T, H*H+2*H*O+O*O) frame-level attention weights                                */
    TGGCN_BUF_PGATE,         /* (B*T, H+O)           sigmoid probability of the sampled gates                     */
    TGGCN_BUF_SGATES_H,      /* (B,T,H,2,4D)         segment-cell gates                                           */
    TGGCN_BUF_SGATES_O,
    TGGCN_BUF_MG_ALL_H,      /* (2,B,T,H,nk*D)       aggregated segment messages of every step                    */
    TGGCN_BUF_MG_ALL_O,      /* (2,B,T,O,2D)                                                                      */
    TGGCN_BUF_SMSG_HH,       /* (2,B,T,senders,D)    per-sender segment messages (post-ReLU), kinds hh, oh, ho, oo */
    TGGCN_BUF_SMSG_OH,
    TGGCN_BUF_SMSG_HO,
    TGGCN_BUF_SMSG_OO,
    TGGCN_BUF_SALPHA_HH,     /* (2,B,T,receivers,senders) segment-level attention weights                         */
    TGGCN_BUF_SALPHA_OH,
    TGGCN_BUF_SALPHA_HO,
    TGGCN_BUF_SALPHA_OO,
    TGGCN_BUF_PACK,          /* 16-bit operand planes of the projection stage in flight (gemm16.cu)                       */
    TGGCN_BUF_TIME_EMB,      /* (B*T, D)             time-position features (empty unless dims.time_position)             */
    TGGCN_BUF_MSG_GH,        /* (B,T,1,D)            geometry -> human frame message (empty unless dims.geo_to_human)     */
    TGGCN_BUF_SEG_LEN,       /* (B,T,H+O)            segment lengths (empty unless dims.segment_length)                   */
    TGGCN_BUF_GATE_IN_H,     /* (B,T,H,in_h)         gate MLP inputs [x, h, m_hh, m_oh, (m_gh), (time)] (dims.gate_layers == 2 only) */
    TGGCN_BUF_GATE_IN_O,     /* (B,T,O,in_o)         [x, h, m_ho, m_oo, m_go, (time)]                                    */
    TGGCN_BUF_GATE_HID_H,    /* (B,T,H,D)            hidden layer of the gate MLPs (post-ReLU)                            */
    TGGCN_BUF_GATE_HID_O,    /* (B,T,O,D)                                                                                 */
    TGGCN_BUF_GATE_HID2_H,   /* (B,T,H,D)            second hidden layer (dims.gate_layers == 3 only)                     */
    TGGCN_BUF_GATE_HID2_O,   /* (B,T,O,D)                                                                                 */
    TGGCN_BUF_COUNT
};

TGGCN_API int         tggcn_abi_version(void);
TGGCN_API const char* tggcn_last_error(void);

/* Bytes of device scratch tggcn_forward needs for these dims. */
TGGCN_API size_t tggcn_workspace_bytes(const tggcn_dims* dims);
/* Offset/size of one named region inside the workspace (for tests / debugging). */
TGGCN_API int    tggcn_workspace_view(const tggcn_dims* dims, int buf_id, size_t* offset, size_t* bytes);

/* 0 if no grid barrier of the persistent kernels timed out and no operand left the range of the fp16-split tiles during the
 * work queued so far on `stream` for this workspace (synchronises the stream), non-zero otherwise. */
TGGCN_API int tggcn_sync_status(const tggcn_dims* dims, const void* workspace, void* stream);
/* Interpret 8 status words (host memory, as delivered through tggcn_io.status_host): 0 = healthy; bit 0 of the result = a grid
 * barrier timed out (results undefined), bit 1 = fp16-split range violation (|w| >= 255 or an activation >= 65504: results of
 * that call are invalid; rerun with dims.no_fp16_split).  The message is available through tggcn_last_error(). */
TGGCN_API int tggcn_status_decode(const uint32_t* status_words_host);

/* The whole forward pass: replaces TGGCN.forward (vhoi/models.py:584-933).
 * `weights` holds TGGCN_W_COUNT device pointers ordered by enum tggcn_weight_id. */
TGGCN_API int tggcn_forward(const tggcn_dims* dims, const void* const* weights, int n_weights, const tggcn_io* io,
                  void* workspace, size_t workspace_bytes, void* stream);

/* Stages of the forward, in launch order (tggcn_forward_profile reports one duration per stage). */
enum tggcn_stage_id {
    TGGCN_STAGE_GEO_GCN = 0,     /* geo_gcn_kernel                                   (K-A)               */
    TGGCN_STAGE_GEMM_EMBED,      /* ROI embeddings + geometry MLP layer 0            (K-B)               */
    TGGCN_STAGE_GEMM_GEO2,       /* geometry MLP layer 2                                                  */
    TGGCN_STAGE_GEMM_GI,         /* BiGRU input pre-activations, 3 groups x 2 directions                  */
    TGGCN_STAGE_BIGRU,           /* persistent BiGRU recurrences                     (K-C)               */
    TGGCN_STAGE_GEMM_BD,         /* Linear(2D->D) on BiGRU outputs                                        */
    TGGCN_STAGE_GEMM_MSG,        /* per-sender frame message MLPs                                         */
    TGGCN_STAGE_FRAME_MSG,       /* attention + aggregation + gates                  (K-D)               */
    TGGCN_STAGE_GATE_POST,       /* filter + reorder index                           (K-E)               */
    TGGCN_STAGE_GEMM_GS,         /* hoisted frame-part of the segment cells' W_ih x                       */
    TGGCN_STAGE_SEGMENT,         /* persistent segment-level recurrent graph         (K-F)               */
    TGGCN_STAGE_HEADS,           /* label heads                                      (K-G)               */
    TGGCN_STAGE_COUNT
};

/* Same as tggcn_forward, but brackets every stage with CUDA events on `stream`, synchronises the stream and
 * writes TGGCN_STAGE_COUNT durations (milliseconds) to stage_ms_host.  Measurement aid for bench.py. */
TGGCN_API int tggcn_forward_profile(const tggcn_dims* dims, const void* const* weights, int n_weights,
                                    const tggcn_io* io, void* workspace, size_t workspace_bytes, void* stream,
                                    float* stage_ms_host);

/* Number of kernels this library has launched since it was loaded (host-side counter). */
TGGCN_API unsigned long long tggcn_launch_count(void);

/* Kernel-family entry points (also used by tggcn_forward). */

/* Geo_gcn.forward + the input split of models.py:631-642 (pyrutils/torch/models_gcn.py:30-100).
 * x_human (B,T,H,Fh) -> out (B,128,V,T).  gcn_weights: the 13 GCN_* pointers of the weight table. */
TGGCN_API int tggcn_geo_gcn_fwd(const float* x_human, const void* const* weights, float* out, float* bn_running_mean,
                      float* bn_running_var, int64_t* bn_num_batches, void* workspace, int B, int T, int H, int V,
                      int Fh, int bn_train, void* stream);

/* nn.Linear (+ReLU): C[M,N] = act(A[M,K] W[N,K]^T + bias[N]) with leading dimensions
 * (build_mlp, pyrutils/torch/models.py:31-33).  gemm_path as in tggcn_dims. */
TGGCN_API int tggcn_linear_fwd(const float* A, int lda, const float* W, int ldw, const float* bias, float* C, int ldc,
                     int M, int N, int K, int relu, int gemm_path, void* stream);

/* The same nn.Linear on the TMA-fed tcgen05 kernel over 16-bit operand planes (csrc/gemm16.cu), which the forward uses whenever
 * K % 64 == 0 and N % 16 == 0: precision 0 = fp16 (hi, lo) operand split, fp32-class accuracy; 1 = bf16 operands.
 * scratch: tggcn_linear16_scratch_bytes(M, N, K) bytes of device memory, 256-byte aligned (the operand planes).
 * status: one device word, bit 1 is set when an operand leaves the fp16 range (precision 0); may be NULL. */
TGGCN_API size_t tggcn_linear16_scratch_bytes(int M, int N, int K);
TGGCN_API int tggcn_linear16_fwd(const float* A, int lda, const float* W, int ldw, const float* bias, float* C, int ldc,
                       int M, int N, int K, int relu, int precision, void* scratch, size_t scratch_bytes,
                       uint32_t* status, void* stream);

/* Upstream gradients of the forward's output list (models.py:919-926), same shapes as the outputs; NULL = no gradient. */
typedef struct tggcn_grad_outputs {
    const float* d_y_hs;  const float* d_y_hss;   /* (B,T,H) hard / soft human gates                              */
    const float* d_y_os;  const float* d_y_oss;   /* (B,T,O)                                                      */
    const float* d_out_h[4];                      /* (B,C_sub,T,H) frame_rec, frame_pred, seg_rec, seg_pred       */
    const float* d_out_o[4];                      /* (B,C_aff,T,O)                                                */
} tggcn_grad_outputs;

/* Backward of tggcn_forward (what loss.backward() does through vhoi/models.py:584-933).  The forward must have run with
 * dims.save_for_backward on the same `workspace`, which must still hold its contents.  `grad_weights` holds TGGCN_W_COUNT
 * device pointers (same order as `weights`): each non-NULL entry receives the gradient of that parameter (overwritten);
 * NULL entries are skipped where possible.  `bwd_workspace`: tggcn_backward_workspace_bytes(dims) bytes of scratch. */
TGGCN_API size_t tggcn_backward_workspace_bytes(const tggcn_dims* dims);
TGGCN_API int tggcn_backward(const tggcn_dims* dims, const void* const* weights, void* const* grad_weights, int n_weights,
                             const tggcn_io* io, const tggcn_grad_outputs* grads, void* workspace, size_t workspace_bytes,
                             void* bwd_workspace, size_t bwd_workspace_bytes, void* stream);

/* Gradient buckets for a data-parallel caller (the reference is single-device).  The backward finishes the
 * parameter gradients in reverse model order; bucket k is complete when stage group k has been queued:
 *   0  label heads + segment-level cells and message MLPs (after the segment BPTT and its weight-gradient GEMMs)
 *   1  gate MLPs, frame-level message MLPs, Linear(2D->D), BiGRU weights
 *   2  ROI embeddings + geometry MLP
 *   3  geometry GCN
 * tggcn_backward_bucket(weight id) names the bucket of every weight-table entry, so the caller can lay its gradient buffer out
 * bucket by bucket; tggcn_backward_ex records hooks->bucket_done[k] (a cudaEvent_t, or NULL) on `stream` when bucket k is
 * complete, so an all-reduce of bucket k on another stream can wait for exactly that event and overlap the rest of the backward. */
#define TGGCN_BWD_BUCKETS 4
typedef struct tggcn_bwd_hooks {
    void* bucket_done[TGGCN_BWD_BUCKETS];
} tggcn_bwd_hooks;
TGGCN_API int tggcn_backward_bucket(int weight_id);
TGGCN_API int tggcn_backward_ex(const tggcn_dims* dims, const void* const* weights, void* const* grad_weights, int n_weights,
                                const tggcn_io* io, const tggcn_grad_outputs* grads, void* workspace, size_t workspace_bytes,
                                void* bwd_workspace, size_t bwd_workspace_bytes, void* stream, const tggcn_bwd_hooks* hooks);

/* Backward of nn.Linear (+ReLU), i.e. what autograd does for y = act(x W^T + b) (pyrutils/torch/models.py:31-33):
 *   Z = dY (.) [Y > 0] (Y = forward output, NULL when there was no ReLU);  dX = Z W (added to dX when beta_dx);
 *   dW = Z^T X;  db = column sums of Z.  Any of dX / dW / db may be NULL.  wt_scratch: K*N floats (W^T) when dX != NULL. */
TGGCN_API int tggcn_linear_bwd(const float* dY, int ldy, const float* Y, int ldyf, const float* X, int ldx,
                               const float* W, int ldw, float* dX, int lddx, int beta_dx, float* dW, int lddw,
                               float* db, float* wt_scratch, int M, int N, int K, int gemm_path, void* stream);

/* One group of the frame-level bidirectional GRU recurrence (nn.GRU as called at vhoi/models.py:997-1000), given the
 * hoisted input pre-activations gi = W_ih x + b_ih (B,T,E,2,3D).  hfr (B,T,E,2D) out; gates (B,T,E,2,4D) out or NULL
 * (r, z, n, W_hn h + b_hn: what the backward needs); sync: 16 bytes of device scratch. */
TGGCN_API int tggcn_bigru_fwd(const float* gi, const float* whh_f, const float* whh_b, const float* bhh_f,
                              const float* bhh_b, float* hfr, float* gates, void* sync, int B, int T, int E, int D,
                              int persistent, void* stream);
/* Backward through time of the same recurrence (autograd of nn.GRU).  dhfr: gradient w.r.t. hfr.  Outputs: dgi
 * (gradient w.r.t. gi), dgh (scratch of the same size), dW_hh / db_hh per direction (may be NULL). */
TGGCN_API size_t tggcn_bigru_bwd_scratch_floats(int B, int T, int E, int D);
TGGCN_API int tggcn_bigru_bwd(const float* dhfr, const float* hfr, const float* gates, const float* whh_f,
                              const float* whh_b, float* dgi, float* dgh, float* dwhh_f, float* dwhh_b, float* dbhh_f,
                              float* dbhh_b, float* scratch, int B, int T, int E, int D, int gemm_path, void* stream);

/* ---- fused criterion: multi_task_loss (pyrutils/torch/losses.py:39-51) with the function tuple of vhoi/losses.py:41-60 ----
 * One term per model output.  budget_loss (pyrutils/torch/losses.py:24-36) and binary_cross_entropy_loss (:7-21) take the
 * (B,T,E) gate outputs with float targets (-1 = ignore); F.nll_loss(ignore_index=-1, 'mean') takes (B,C,T,E) log-probs with
 * int64 targets (B,T,E). */
enum tggcn_loss_kind { TGGCN_LOSS_BUDGET = 0, TGGCN_LOSS_BCE = 1, TGGCN_LOSS_NLL = 2 };
typedef struct tggcn_loss_term {
    int32_t kind;                /* enum tggcn_loss_kind                                                          */
    float   weight;              /* vhoi/losses.py:8-61                                                            */
    const float* out;            /* model output                                                                   */
    const void*  target;         /* float (budget / bce) or int64 (nll), -1 = ignore                               */
    float*  d_out;               /* backward: gradient w.r.t. `out`, same shape, overwritten; NULL = skip           */
    int64_t numel;               /* budget / bce: elements; nll: B*T*E positions                                    */
    int32_t B, C, T, E;          /* nll only                                                                        */
} tggcn_loss_term;
/* losses: n_terms floats (device) = weight_i * loss_i; scratch: 2*n_terms floats (device), kept for the backward. */
TGGCN_API int tggcn_loss_fwd(const tggcn_loss_term* terms, int n_terms, float* losses, float* scratch, void* stream);
/* grad_losses: n_terms floats (device) upstream gradients of the loss values, or NULL for ones. */
TGGCN_API int tggcn_loss_bwd(const tggcn_loss_term* terms, int n_terms, const float* scratch, const float* grad_losses, void* stream);

/* ---- evaluation post-processing: what predict.py does on the host after the forward ------------------------
 * tggcn_upsample_argmax: predict.py:64-70 (torch.repeat_interleave(out, downsampling, dim=-2) + match_shape :95-122) followed
 *   by np.argmax(axis=1) of process_output (:186-202).  logp (B,C,T,E) float -> labels (B,Tt,E) int64.
 * tggcn_f1_at_k: pyrutils/metrics.py:7-81 (f1_at_k / f1_at_k_single_example) on the rows predict.py:236-240 builds
 *   ((B,Tt,E) -> swapaxes(1,2) -> (B*E, Tt)), frames whose target equals ignore_value removed.  Per row and overlap the F1 in
 *   f1_rows[k * B*E + row] (double, device) and valid_rows[row] = the row has at least one frame; the caller averages the valid
 *   rows in row order (as the reference's accumulation does).  overlaps: n_overlaps doubles (device).
 *   scratch: tggcn_f1_at_k_scratch_bytes(B, Tt, E) bytes (device). */
TGGCN_API int tggcn_upsample_argmax(const float* logp, int64_t* labels, int B, int C, int T, int E, int Tt, int downsampling,
                                    void* stream);
TGGCN_API size_t tggcn_f1_at_k_scratch_bytes(int B, int Tt, int E);
TGGCN_API int tggcn_f1_at_k(const int64_t* target, const int64_t* pred, int B, int Tt, int E, int num_classes,
                            const double* overlaps, int n_overlaps, int64_t ignore_value, void* scratch, double* f1_rows,
                            int32_t* valid_rows, void* stream);

/* One step of torch.optim.Adam (the optimiser train.py:40 builds; amsgrad off, weight_decay added to the gradient) over flat,
 * 16-byte aligned device buffers of n floats: parameters p, gradients g, first / second moments m, v (zero before step 1).
 * `step` is the 1-based update count (bias corrections 1 - beta^step).  One launch; used by 2g-gcn_b200/optim.py (FlatAdam). */
TGGCN_API int tggcn_adam_step(float* p, const float* g, float* m, float* v, size_t n, float lr, float beta1, float beta2, float eps,
                              float weight_decay, int step, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TGGCN_B200_H_ */
