// tcgen05/TMEM implicit-GEMM path of the modulated conv (placeholder until the kernel lands: reports "unsupported" so
// every call takes the generic SIMT path).
#include "modconv_common.cuh"

namespace vfm {
namespace modconv {
bool tc_supported(const vfm_modconv_desc&) { return false; }
size_t tc_workspace_bytes(const vfm_modconv_desc&, int) { return 0; }
int tc_forward(const vfm_modconv_fwd_params&, const Coefs&, void*, size_t, cudaStream_t) { set_error("tcgen05 path not built"); return VFM_ERR_NO_KERNEL; }
int tc_backward(const vfm_modconv_bwd_params&, const Coefs&, float*, float*, void*, size_t, cudaStream_t) { set_error("tcgen05 path not built"); return VFM_ERR_NO_KERNEL; }
}  // namespace modconv
}  // namespace vfm
