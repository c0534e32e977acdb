// flat_tensor.cu — K2: tensor-core (tcgen05) candidate filter for large query batches.
#include "store.h"

namespace scn {

bool tensor_path_supported(const scn_store*, uint32_t) { return false; }

int32_t flat_search_tensor(scn_store*, const float*, uint64_t, uint32_t, uint64_t, uint64_t*, cudaStream_t, Profiler*) {
  return fail(SCN_ERR_INTERNAL, "tensor-core flat path not built");
}

}  // namespace scn
