// Parameter blocks of the recurrent kernels.
#pragma once
#include "common.cuh"

namespace tg {

struct BiGruGroup {
    const float* gi;        // (B,T,E,2,3D) input pre-activations incl. b_ih, [dir][gate][unit]
    float* hfr;             // (B,T,E,2D)   outputs [fwd | bwd]; also the recurrent state
    float* gates;           // (B,T,E,2,4D) r, z, n, hn saved for the backward, or null
    const float* whh[2];    // (3D,D) per direction
    const float* bhh[2];    // (3D)
    int E;                  // entities of this group
    int rows;               // B*E
    int cfg, jeff, n_rb, n_ub, tile_begin;   // tiling (filled by the launcher); cfg: see rec_cfg_for_rows
    int skip_dirs;          // resident kernel only: bit d set = direction d of this group is computed elsewhere (hybrid launch)
};

struct BiGruParams {
    BiGruGroup g[3];
    int ngroups;
    int B, T, D;
    int total_tiles;
    int no_fp16_split;      // 1: never the SMEM-resident fp16-split variant (dims.no_fp16_split)
    void* big_ws;           // non-null: large-batch path (step_tc.cu) on this TGGCN_BUF_BIG region
    int precision;          // dims.precision (large-batch path)
    GridSync sync;
};

int launch_bigru(BiGruParams& P, int persistent, cudaStream_t stream);
// SMEM-resident W_hh variant (bigru_res.cu): 0 = launched, -1 = shape does not qualify (use launch_bigru's streaming kernel), > 0 = error
int launch_bigru_resident(BiGruParams& P, cudaStream_t stream, int nblk = 3);     // nblk: blocks of 8 hidden units per CTA (3 or 2)
// cluster / distributed-shared-memory variant (bigru_cl.cu, hidden_size 512, small batches): same return convention
int launch_bigru_cluster(BiGruParams& P, cudaStream_t stream);
int launch_transpose(const float* in, int ldi, float* out, int ldo, int R, int C, cudaStream_t stream);
int launch_transpose_prep(const float* in, int ldi, const float* mask, int ldm, float* out, int Mp, int M, int C, int shift, int period,
                          float* colsum, cudaStream_t stream);
int launch_gemm_tn(const float* Z, int ldz, const float* mask, int ldm, const float* A, int lda, float* C, int ldc, int M, int N,
                   int K, int a_shift, int period, int beta, cudaStream_t stream);
int launch_colsum(const float* Z, int ldz, const float* mask, int ldm, float* out, int M, int N, int beta, cudaStream_t stream);

// ---- segment-level recurrent graph (segment.cu) ------------------------------------------------
struct SegParams {
    int B, T, H, O, D;
    int hh;                       // humans->human messages on
    int mean_pool;                // message_aggregation 'mp': uniform weights over the valid senders instead of attention
    int att_noscale;              // attention_style 'v2': plain dot-product logits (no 1/sqrt(D))
    const float* dist[3];         // distance-based attention: hh (B,T,H,H), ho (B,T,H,O), oo (B,T,O,O); each may be null
    // hoisted frame-part pre-activations (incl. b_ih) and gates
    const float* gs_h;            // (B,T,H,2,3D)
    const float* gs_o;            // (B,T,O,2,3D)
    const float* u_h;             // (B,T,H) hard gates
    const float* u_o;             // (B,T,O)
    const float* om;              // (B,O)
    // cell weights per direction: segment-message columns of W_ih, W_hh, b_hh
    const float* wih_h[2]; int ldw_h; int col_h;    // human cell: W_ih (3D, ldw_h), segment part starts at col_h
    const float* wih_o[2]; int ldw_o; int col_o;
    const float* whh_h[2]; const float* bhh_h[2];
    const float* whh_o[2]; const float* bhh_o[2];
    // segment message MLPs (D,D) + bias; kinds: 0 hh, 1 oh (receiver human) ; 2 ho, 3 oo (receiver object)
    const float* wm[4]; const float* bm[4];
    // state / outputs
    float* hx_h;                  // (B,T,H,2D) [fwd | bwd]
    float* hx_o;                  // (B,T,O,2D)
    float* mg_h;                  // (2,B,H,nk_h*D) aggregated segment messages per direction
    float* mg_o;                  // (2,B,O,2D)
    float* att_f; float* att_b;   // (B,H,T,O) or null
    // saved for the backward (null / mg_T == 1 in inference): see tggcn_backward
    int mg_T;                     // T: mg_* hold every step [dir][b][t][e][..]; 1: a single step is kept
    float* sgates_h;              // (B,T,H,2,4D) r, z, n, hn of the human cells
    float* sgates_o;              // (B,T,O,2,4D)
    float* smsg[4];               // per kind (hh, oh, ho, oo): [dir][b][t][sender][D] post-ReLU messages
    float* salpha[4];             // per kind: [dir][b][t][receiver][sender] attention weights
    // tiling (filled by the launcher)
    int nk_h;                     // message kinds feeding the human cell (2 with hh, else 1)
    int bbv[4], n_vb[4];          // per message kind: videos per message tile, number of video blocks
    int msg_tiles_kind[4];        // unit blocks per kind
    int msg_tile_begin[5];        // prefix over kinds (per direction)
    int msg_tiles_dir;            // message tiles per direction
    int cfg_h, jeff_h, nrb_h, nub_h;
    int cfg_o, jeff_o, nrb_o, nub_o;
    int cell_tiles_h_dir, cell_tiles_dir;
    int tilesA, tilesB;
    int msg_g[4];                 // per message kind: video-block groups (= n_vb unless res_multi: a tile then walks over n_vb / msg_g blocks)
    int cell_g_h, cell_g_o;       // row-block groups of the cell tiles (= nrb unless res_multi: a tile then walks over nrb / cell_g row blocks)
    int res_multi;                // resident-weight variant, larger batches: tiles are weight slices, row / video blocks are walked inside the CTA
    int res_msg;                  // resident-weight variant: message-tile weights are kept in shared memory too
    int res_ring_floats;          // resident-weight variant: floats of the cp.async ring that precede the overflow fragments in shared memory
    int no_fp16_split;            // 1: never the on-chip resident fp16-split variant (dims.no_fp16_split)
    void* big_ws;                 // non-null: large-batch path (step_tc.cu) on this TGGCN_BUF_BIG region
    int precision;                // dims.precision (large-batch path)
    GridSync sync;
};

int launch_segment(SegParams& P, int persistent, cudaStream_t stream);

}  // namespace tg
