// query_tc.cu -- tensor-core (tcgen05) engine of vsm_query.  Placeholder until the kernel lands: fails loudly.
#include "state.cuh"

namespace vsm {
int query_tc(vsm_map*, const float*, int, int, int, int64_t*, float*, cudaStream_t) {
  set_error("vsm_query: engine 2 (tcgen05) is not built in this version");
  return VSM_E_STATE;
}
}  // namespace vsm
