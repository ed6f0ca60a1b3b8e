// Temporary: tcgen05 family not built yet.
#include "mlstm_common.cuh"
namespace mlstm {
bool tc_supported(const mlstm_params&) { return false; }
size_t tc_bwd_workspace(const mlstm_params&) { return 0; }
int tc_fwd(const mlstm_params&, cudaStream_t) { set_error("tcgen05 family not built"); return MLSTM_ERR_UNSUPPORTED; }
int tc_bwd(const mlstm_params&, cudaStream_t) { set_error("tcgen05 family not built"); return MLSTM_ERR_UNSUPPORTED; }
}
