// tcgen05 (5th-gen tensor core) projection path — placeholder until the 3xTF32 kernel lands.
#include "common.cuh"
#include "gemm.h"

namespace tg {

int launch_gemm_tc(GemmGroup& grp, cudaStream_t stream) {
    (void)grp; (void)stream;
    set_error("gemm_path=1 (tcgen05) is not built into this library");
    return 3;
}

}  // namespace tg
