// Shared between api.cu (forward) and api_bwd.cu (backward): workspace layout and argument checks.
#pragma once
#include "common.cuh"

namespace tg {

struct Layout {
    size_t off[TGGCN_BUF_COUNT];
    size_t bytes[TGGCN_BUF_COUNT];
    size_t total;
};

inline int nkh_of(const tggcn_dims& d) { return d.hh ? 2 : 1; }
// widths (in units of D) around the segment cells' input rows: [h, m.. (, time)] frame part, then the segment-message columns
inline int ts_of(const tggcn_dims& d) { return d.time_position == 1 ? 1 : 0; }       // time block in the segment-level inputs
inline int tu_of(const tggcn_dims& d) { return d.time_position == 2 ? 1 : 0; }       // time block in the gate MLP inputs
inline int gh_of(const tggcn_dims& d) { return d.geo_to_human ? 1 : 0; }             // geometry -> human message block
inline int tl_of(const tggcn_dims& d) { return d.segment_length ? 1 : 0; }           // segment-length block (last of the frame part)
inline int kh_of(const tggcn_dims& d) { return (1 + nkh_of(d) + gh_of(d) + ts_of(d) + tl_of(d)) * d.D; }   // xx_h row = frame-part columns of the human W_ih
inline int ldwh_of(const tggcn_dims& d) { return kh_of(d) + nkh_of(d) * d.D; }
inline int ko_of(const tggcn_dims& d) { return (4 + ts_of(d) + tl_of(d)) * d.D; }               // xx_o row
inline int ldwo_of(const tggcn_dims& d) { return ko_of(d) + 2 * d.D; }
// gate MLP input widths: [x, h, m_hh?, m_oh, m_gh?, time_u?] / [x, h, m_ho, m_oo, m_go, time_u?]
inline int ginh_of(const tggcn_dims& d) { return (2 + nkh_of(d) + gh_of(d) + tu_of(d)) * d.D; }
inline int gino_of(const tggcn_dims& d) { return (5 + tu_of(d)) * d.D; }
inline bool gate2_of(const tggcn_dims& d) { return d.gate_layers >= 2; }      // gate MLPs with hidden layers
inline bool gate3_of(const tggcn_dims& d) { return d.gate_layers == 3; }
void make_layout(const tggcn_dims& d, Layout& L);
int check_dims(const tggcn_dims& d);

}  // namespace tg
