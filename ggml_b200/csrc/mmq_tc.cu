// mmq_tc.cu — batched quantized GEMM on the tcgen05 tensor cores (placeholder until the kernel lands:
// reports "not eligible" so that AUTO falls through to the generic kernel; never silently computes elsewhere).
#include "b200_internal.h"

namespace b200 {
bool   mmq_tc_eligible(const ggml_b200_mul_mat_args &) { return false; }
size_t mmq_tc_workspace(const ggml_b200_mul_mat_args &) { return 0; }
int    launch_mmq_tc(const ggml_b200_mul_mat_args &, cudaStream_t) { set_error("tcgen05 GEMM not built"); return GGML_B200_EUNSUPPORTED; }
} // namespace b200
