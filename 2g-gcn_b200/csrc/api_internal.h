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
void make_layout(const tggcn_dims& d, Layout& L);
int check_dims(const tggcn_dims& d);

}  // namespace tg
