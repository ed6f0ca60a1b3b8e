// Shape/dtype gate of the tcgen05 family (forward: mlstm_tc_fwd.cu, backward: mlstm_tc_bwd.cu).
#include "mlstm_common.cuh"
namespace mlstm {
bool tc_supported(const mlstm_params& p) {
  return p.dtype == MLSTM_BF16 && p.DHQK == p.DHV && (p.DHQK == 64 || p.DHQK == 128);
}
}  // namespace mlstm
