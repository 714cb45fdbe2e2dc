// Large-batch recurrent path: declarations shared by api.cu (workspace sizing, path selection) and step_tc.cu.
#pragma once
#include "common.cuh"
#include "bigru.h"

namespace tg {

// Carving of the TGGCN_BUF_BIG workspace region (all offsets in bytes, 1024-byte aligned).
struct BigLayout {
    // fp16 (or bf16) operand copies of the recurrent weights, planes [dir][hi, lo] of row-major [3D][K] matrices
    size_t whh_g[3];        // BiGRU W_hh per group (humans, objects, geometry): [2][2][3D][D]
    size_t wih_h, wih_o;    // segment cells, segment-message columns of W_ih: [2][2][3D][nk_h*D], [2][2][3D][2D]
    size_t whh_h, whh_o;    // segment cells W_hh: [2][2][3D][D]
    size_t wm;              // segment message MLPs: [4 kinds][2][D][D]
    // state rings: planes [slot][dir][hi, lo] of [rows][D]
    size_t ring_g[3];       // BiGRU groups
    size_t ring_h, ring_o;  // segment level
    // aggregated segment messages as GEMM operands: planes [dir][hi, lo] of [rows][nk*D]
    size_t mg_h, mg_o;
    // per-step post-ReLU sender messages (inference; training writes the SMSG_* save buffers instead): [dir][kind][rows_s][D] fp32
    size_t msg;
    size_t total;
};

void big_layout(int B, int H, int O, int D, int hh, BigLayout& L);

// true when the forward should run the recurrent stages on the large-batch path (dims.recurrent_mode, rows per step, shape limits)
bool use_big_path(const tggcn_dims& d, int stage);      // stage 0: BiGRUs, 1: segment level

int launch_bigru_big(BiGruParams& P, void* big_ws, int precision, cudaStream_t stream);
int launch_segment_big(SegParams& P, void* big_ws, int precision, int T_save, cudaStream_t stream);

}  // namespace tg
