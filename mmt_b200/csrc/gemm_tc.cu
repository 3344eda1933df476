// placeholder until the tcgen05 kernel lands (replaced below in this round)
#include "common.cuh"
namespace mmt {
int gemm_tc(const mmt_gemm_desc& d, cudaStream_t stream) {
  set_error("mmt_gemm: MMT_PREC_TF32 not built");
  return MMT_E_UNSUPPORTED;
}
}
