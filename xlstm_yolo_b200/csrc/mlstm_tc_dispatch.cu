// Dispatch glue of the tcgen05 family.  Forward: mlstm_tc_fwd.cu.  Backward: until the
// tcgen05 backward lands, bf16 gradients run on the fp32 SIMT kernels (same saved rows).
#include "mlstm_common.cuh"
namespace mlstm {
bool tc_supported(const mlstm_params& p) {
  return p.dtype == MLSTM_BF16 && p.DHQK == p.DHV && (p.DHQK == 64 || p.DHQK == 128);
}
size_t tc_bwd_workspace(const mlstm_params& p) { return simt_bwd_workspace(p); }
int tc_bwd(const mlstm_params& p, cudaStream_t st) { return simt_bwd(p, st); }
}
